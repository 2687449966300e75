// superinstr.cu -- the elementwise CC super-instructions that sit BETWEEN contractions in the reference's pardo
// bodies (rlccd_rhf.sialx:262,572,598; rccsdpt_aab.sialx:550-566,1829), on resident device blocks, so that a block
// does not bounce to the host between two device contractions.  HBM-bound (16 bytes per element, read + write).
//
// Reference behaviour followed (not code): src/sip/super_instructions/qm/qm-generic/energy_denominator_rhf.F:15-460,
// qm/utility/stripi.F, anti_symm_o.F, anti_symm_v.F, return_sval.F, invert_diagonal.F; calling convention
// special_instructions.h:27-97 (per argument: array slot, rank, index values, size, extents, data; then ierr).
// The predefined int array "moa_seg_ranges" that the Fortran routines fetch through the sip_interface upcall
// (sip_interface.h:17-35) is registered once with sipgpu_set_predefined_int_array.
#include <map>
#include <string>
#include <vector>

#include "elementwise.h"

namespace sipgpu {
namespace {

constexpr int kT = 256;

std::map<std::string, std::vector<int>>& int_arrays() {
    static std::map<std::string, std::vector<int>> m;
    return m;
}

// offset of segment `index_value` = sum of the extents of the segments before it (energy_denominator_rhf.F:66-90)
int seg_offset(const std::vector<int>& seg, int index_value, int* out) {
    if (index_value < 1 || index_value - 1 > (int)seg.size()) return SIPGPU_E_ARG;
    int off = 0;
    for (int i = 0; i < index_value - 1; ++i) off += seg[i];
    *out = off;
    return SIPGPU_OK;
}

struct DenArgs {
    int rank;
    int ext[6], off[6];
    int fock_ld;  // 0: rank-1 Fock array
    long long n;
    const double* shift;  // energy_ty_denominator_rhf: a device scalar added to eps (nullptr: none)
};

__global__ void __launch_bounds__(kT) energy_denominator_kernel(double* __restrict__ d, const double* __restrict__ fock,
                                                                const __grid_constant__ DenArgs a) {
    for (long long lin = (long long)blockIdx.x * kT + threadIdx.x; lin < a.n; lin += (long long)gridDim.x * kT) {
        long long r = lin;
        double e[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            e[k] = 0.0;
            if (k < a.rank) {
                const int p = (int)(r % a.ext[k]) + a.off[k];  // 0-based global orbital index
                r /= a.ext[k];
                e[k] = a.fock_ld ? __ldg(fock + (size_t)p * a.fock_ld + p) : __ldg(fock + p);
            }
        }
        // same association as the reference: epsb + epsd (+ epsf) - epsa - epsc (- epse)
        double eps;
        if (a.rank == 2) eps = e[1] - e[0];
        else if (a.rank == 4) eps = e[1] + e[3] - e[0] - e[2];
        else eps = e[1] + e[3] + e[5] - e[0] - e[2] - e[4];
        if (a.shift) eps = eps + __ldg(a.shift);   // energy_ty_denominator_rhf.F: epsb + epsd - epsa - epsc + shift
        d[lin] = d[lin] / eps;
    }
}

struct SymArgs {
    int ext[4], off[4];
    long long n;
    int special;
};

// one thread per element of the block; phase 0 mirrors, phase 1 zeroes (two launches keep the reference's order)
template <bool OCC>
__global__ void __launch_bounds__(kT) anti_symm_kernel(double* __restrict__ x, const __grid_constant__ SymArgs s, int phase) {
    const long long s1 = s.ext[0], s2 = s1 * s.ext[1], s3 = s2 * s.ext[2];
    for (long long lin = (long long)blockIdx.x * kT + threadIdx.x; lin < s.n; lin += (long long)gridDim.x * kT) {
        long long r = lin;
        const int a = (int)(r % s.ext[0]); r /= s.ext[0];
        const int i = (int)(r % s.ext[1]); r /= s.ext[1];
        const int b = (int)(r % s.ext[2]); r /= s.ext[2];
        const int j = (int)r;
        const int ga = a + s.off[0], gi = i + s.off[1], gb = b + s.off[2], gj = j + s.off[3];
        if (phase == 0) {
            if (OCC) {
                if (gi < gj) {
                    const int mi = gj - s.off[1], mj = gi - s.off[3];
                    if (mi >= 0 && mi < s.ext[1] && mj >= 0 && mj < s.ext[3]) x[a + mi * s1 + b * s2 + mj * s3] = x[lin] * (-1.0);
                }
            } else {
                const int ma = gb - s.off[0], mb = ga - s.off[2];
                const bool inside = ma >= 0 && ma < s.ext[0] && mb >= 0 && mb < s.ext[2];
                if (ga < gb && inside) x[ma + i * s1 + mb * s2 + j * s3] = x[lin] * (-1.0);
                if (s.special && ga == gb && inside) x[ma + i * s1 + mb * s2 + j * s3] = 0.0 * (-1.0);
            }
        } else if (gi == gj || ga == gb) {
            x[lin] = 0.0;
        }
    }
}

__global__ void __launch_bounds__(kT) invert_diagonal_kernel(double* __restrict__ a1, const double* __restrict__ a2, long long n) {
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n; i += (long long)gridDim.x * kT) {
        const double div = a2[i];
        if (div != 0.0) a1[i] = a1[i] / div;
    }
}

// invert_diagonal_asym.F do_return_inv5_as: rank 5 (a, b, c, d, e): where global b != d and c != e, a1 /= a2 (skipped where a2 == 0);
// everywhere else a1 = 0.  off[1..4] = segment offsets of b .. e.
struct Inv5Args {
    int ext[5], off[5];
    long long n;
};
__global__ void __launch_bounds__(kT) invert_diagonal_asym_kernel(double* __restrict__ a1, const double* __restrict__ a2, const __grid_constant__ Inv5Args s) {
    for (long long lin = (long long)blockIdx.x * kT + threadIdx.x; lin < s.n; lin += (long long)gridDim.x * kT) {
        long long r = lin / s.ext[0];
        const int b = (int)(r % s.ext[1]); r /= s.ext[1];
        const int c = (int)(r % s.ext[2]); r /= s.ext[2];
        const int d = (int)(r % s.ext[3]); r /= s.ext[3];
        const int e = (int)r;
        if (b + s.off[1] != d + s.off[3] && c + s.off[2] != e + s.off[4]) {
            const double div = a2[lin];
            if (div != 0.0) a1[lin] = a1[lin] / div;
        } else {
            a1[lin] = 0.0;
        }
    }
}

// return_diagonal_elements.F ret_diag_tensor2 / 4: keep x(p,p) / x(p,p,r,r), zero the rest (positions inside the block: the
// Fortran declares both dimensions of a pair over the range of the first one)
__global__ void __launch_bounds__(kT) return_diagonal_kernel(double* __restrict__ x, int e0, int e2, long long n) {
    for (long long lin = (long long)blockIdx.x * kT + threadIdx.x; lin < n; lin += (long long)gridDim.x * kT) {
        long long r = lin;
        const int p = (int)(r % e0); r /= e0;
        const int q = (int)(r % e0); r /= e0;
        bool keep = p == q;
        if (e2 > 0) {
            const int rr = (int)(r % e2); r /= e2;
            keep = keep && rr == (int)r;
        }
        if (!keep) x[lin] = 0.0;
    }
}

int grid_for(long long n) {
    long long b = (n + kT - 1) / kT;
    const long long cap = (long long)ctx().num_sms * 8;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

const std::vector<int>* moa_segs() {
    auto it = int_arrays().find("moa_seg_ranges");
    return it == int_arrays().end() ? nullptr : &it->second;
}

}  // namespace

int si_set_int_array(const char* name, int n, const int* values) {
    if (!name || n < 0 || (n > 0 && !values)) return SIPGPU_E_ARG;
    int_arrays()[name] = std::vector<int>(values, values + n);
    return SIPGPU_OK;
}

int si_energy_denominator_rhf(int rank, const int* index_values, const int* ext, double* data, int fock_rank,
                              const int* fock_ext, const double* fock, const double* d_shift) {
    SIP_TRY(ensure_init());
    if (d_shift && !(rank == 4 && fock_rank == 2)) return SIPGPU_E_ARG;   // the shifted form exists for rank-4 blocks only
    const std::vector<int>* seg = moa_segs();
    if (!seg) { set_error("energy_denominator_rhf: predefined int array moa_seg_ranges is not registered"); return SIPGPU_E_STATE; }
    if (!(rank == 2 || rank == 4 || rank == 6) || !(fock_rank == 1 || fock_rank == 2) || (rank == 6 && fock_rank != 2) ||
        !index_values || !ext || !data || !fock || !fock_ext)
        return SIPGPU_E_ARG;
    DenArgs a;
    memset(&a, 0, sizeof(a));
    a.rank = rank;
    a.n = 1;
    const int nfock = fock_ext[0];
    for (int d = 0; d < rank; ++d) {
        a.ext[d] = ext[d];
        if (rank == 6 && ext[d] == 1) a.off[d] = index_values[d] - 1;
        else SIP_TRY(seg_offset(*seg, index_values[d], &a.off[d]));
        if (ext[d] < 1 || a.off[d] + ext[d] > nfock) return SIPGPU_E_ARG;  // the Fock diagonal must cover the block's range
        a.n *= ext[d];
    }
    a.fock_ld = fock_rank == 2 ? fock_ext[0] : 0;
    a.shift = d_shift;
    energy_denominator_kernel<<<grid_for(a.n), kT, 0, ctx().stream>>>(data, fock, a);
    SIP_CUDA(cudaGetLastError());
    count_launch();
    return SIPGPU_OK;
}

int si_stripi(int rank, const int* iv0, const int* ext0, const double* x, const int* iv1, const int* ext1, double* y) {
    SIP_TRY(ensure_init());
    const std::vector<int>* seg = moa_segs();
    if (!seg) { set_error("stripi: predefined int array moa_seg_ranges is not registered"); return SIPGPU_E_STATE; }
    if (rank < 2 || rank > 4 || !iv0 || !ext0 || !x || !iv1 || !ext1 || !y) return SIPGPU_E_ARG;
    int beg[4], cnt[4];
    for (int d = 0; d < rank; ++d) {
        int o0;
        SIP_TRY(seg_offset(*seg, iv0[d], &o0));
        const int x1 = 1 + o0, x2 = ext0[d] + o0;
        if (ext1[d] == 1 || d == rank - 1) {
            int match = iv1[d];
            if (d != rank - 1) {
                int o1;
                SIP_TRY(seg_offset(*seg, iv1[d], &o1));
                match = 1 + o1;
            }
            if (match < x1 || match > x2) { set_error("stripi: mismatch trying to strip indices"); return SIPGPU_E_ARG; }
            beg[d] = match - x1;
            cnt[d] = 1;
        } else {
            if (ext1[d] != ext0[d]) return SIPGPU_E_ARG;
            beg[d] = 0;
            cnt[d] = ext0[d];
        }
    }
    return ew_slice(rank, x, ext0, y, cnt, beg);  // a strip is a one-index-thick slice of the block (F90:271-330)
}

template <bool OCC>
int anti_symm(int rank, const int* iv, const int* ext, double* x) {
    SIP_TRY(ensure_init());
    const std::vector<int>* seg = moa_segs();
    if (!seg) { set_error("anti_symm: predefined int array moa_seg_ranges is not registered"); return SIPGPU_E_STATE; }
    if (rank != 4 || !iv || !ext || !x) return SIPGPU_E_ARG;
    SymArgs s;
    s.n = 1;
    for (int d = 0; d < 4; ++d) {
        s.ext[d] = ext[d];
        SIP_TRY(seg_offset(*seg, iv[d], &s.off[d]));
        s.n *= ext[d];
    }
    s.special = (!OCC && s.off[1] + 1 == 1 && s.off[1] + ext[1] == 1) ? 1 : 0;
    anti_symm_kernel<OCC><<<grid_for(s.n), kT, 0, ctx().stream>>>(x, s, 0);
    count_launch();
    if (!s.special) {
        anti_symm_kernel<OCC><<<grid_for(s.n), kT, 0, ctx().stream>>>(x, s, 1);
        count_launch();
    }
    SIP_CUDA(cudaGetLastError());
    return SIPGPU_OK;
}
int si_anti_symm_o(int rank, const int* iv, const int* ext, double* x) { return anti_symm<true>(rank, iv, ext, x); }
int si_anti_symm_v(int rank, const int* iv, const int* ext, double* x) { return anti_symm<false>(rank, iv, ext, x); }

int si_return_sval(int rank, const int* ext, const double* data, double* d_scalar) {
    SIP_TRY(ensure_init());
    if (rank < 1 || rank > 2 || !ext || !data || !d_scalar) return SIPGPU_E_ARG;
    long long n = 1;
    for (int d = 0; d < rank; ++d) n *= ext[d];
    SIP_CUDA(cudaMemcpyAsync(d_scalar, data + (n - 1), sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
    return SIPGPU_OK;
}

int si_invert_diagonal(int rank0, int rank1, const int* ext, double* a1, const double* a2) {
    SIP_TRY(ensure_init());
    if (rank0 != rank1 || !(rank0 == 3 || rank0 == 5) || !ext || !a1 || !a2) return SIPGPU_E_ARG;
    long long n = 1;
    for (int d = 0; d < rank0; ++d) n *= ext[d];
    invert_diagonal_kernel<<<grid_for(n), kT, 0, ctx().stream>>>(a1, a2, n);
    SIP_CUDA(cudaGetLastError());
    count_launch();
    return SIPGPU_OK;
}

int si_invert_diagonal_asym(int rank0, int rank1, const int* iv, const int* ext, double* a1, const double* a2) {
    SIP_TRY(ensure_init());
    const std::vector<int>* seg = moa_segs();
    if (!seg) { set_error("invert_diagonal_asym: predefined int array moa_seg_ranges is not registered"); return SIPGPU_E_STATE; }
    if (rank0 != rank1 || rank0 != 5 || !iv || !ext || !a1 || !a2) return SIPGPU_E_ARG;
    Inv5Args s;
    s.n = 1;
    s.off[0] = 0;
    for (int d = 0; d < 5; ++d) {
        s.ext[d] = ext[d];
        if (d > 0) SIP_TRY(seg_offset(*seg, iv[d], &s.off[d]));
        s.n *= ext[d];
    }
    invert_diagonal_asym_kernel<<<grid_for(s.n), kT, 0, ctx().stream>>>(a1, a2, s);
    SIP_CUDA(cudaGetLastError());
    count_launch();
    return SIPGPU_OK;
}

int si_return_diagonal_elements(int rank, const int* ext, double* x) {
    SIP_TRY(ensure_init());
    if (!(rank == 2 || rank == 4) || !ext || !x) return SIPGPU_E_ARG;
    if (ext[0] != ext[1] || (rank == 4 && ext[2] != ext[3])) { set_error("return_diagonal_elements: the block is not square"); return SIPGPU_E_ARG; }
    long long n = 1;
    for (int d = 0; d < rank; ++d) n *= ext[d];
    return_diagonal_kernel<<<grid_for(n), kT, 0, ctx().stream>>>(x, ext[0], rank == 4 ? ext[2] : 0, n);
    SIP_CUDA(cudaGetLastError());
    count_launch();
    return SIPGPU_OK;
}

}  // namespace sipgpu

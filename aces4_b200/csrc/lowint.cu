// lowint.cu -- the bandwidth-shaped contraction kernel: low arithmetic intensity block contractions on sm_100a.
//
// contract.cu is built for the compute-bound contractions (rank-4 x rank-4 -> rank-4: 128-wide DMMA tiles, a
// warp-specialised cp.async ring, per-tile offset tables built with div/mod on the device).  A third of the reference's
// SIAL contraction patterns are not like that (tests/golden/sial_contraction_patterns.txt): their roofline is HBM, not the
// FP64 tensor pipe --
//     D[a,b]     = L[a,i,c,j] * R[b,i,c,j]     rank-2 result of rank-4 blocks (K = 8 000 .. 50 000, M, N <= 50)
//     D[a,i,b,j] = L[a,i,c,j] * R[c,b]         one segment-sized free and contracted index (68 of the 170 patterns)
//     D[a,i]     = L[a,i,b,j] * R[b,j]         matrix-vector shaped (N = 1)
//     D[a,b]     = L[a,c] * R[c,b]             tiny matrices (a few hundred elements per block)
//     D[k1,k2]   = E[a,i,b,j,k1] * E[a,i,b,j,k2]   DIIS matrix elements: M = N = 1, a dot product of 10^6 elements
// (tensor_dil_omp.F90:662-796 runs all of them through permute -> dgemm -> permute; its special cases :768-774 are the
// scalar operands only).  Run through 128-wide tiles these pay for padding (N = 50 -> 56/64/128), for a tile prologue
// that costs more than the tile, and for idle SMs when there are few destinations.  Here every operand element is read
// once and every destination element written once, and nothing else is on the critical path:
//
//   * one launch = one Shape (a pardo body repeats a handful): the three index permutes are offset tables built ONCE on
//     the host per shape (m: {offset in L, offset in D}, n: {offset in R, offset in D}, k: {offset in L, offset in R}),
//     cached on the device and shared by every CTA and every later launch -- no div/mod in the kernel;
//   * work item = (destination block, 64-row tile of M, slice of the contracted range); N <= 64 is one tile.  A launch with
//     few destinations is cut along K (and along the chain of operand pairs) until every SM streams; the partial sums
//     meet in D through red.global.add.f64;
//   * the operands of an item flow through a 3-stage cp.async ring as chunks of KC contracted elements laid out [k][m] /
//     [k][n]; the ring runs ACROSS item boundaries (the first chunks of the next destination are in flight while the
//     current one is finished), so thousands of tiny blocks per launch keep the memory system busy;
//   * math on the FP64 tensor pipe (DMMA m8n8k4) from 8x8 fragments -- M and N are padded to 8, not to a tile of 64/128,
//     and the four warps split the fragment grid 2x2, 1x4 or 4x1, whichever balances (50 x 50 -> 7 x 7 fragments -> 1x4);
//   * M = N = 1 (dot products) skip the tensor pipe: dotk_kernel streams both operands with grid-wide slices.
#include <string>
#include <unordered_map>
#include <vector>

#include "contract.h"
#include "elementwise.h"

namespace sipgpu {
namespace {

constexpr int kLT = 128;      // threads per CTA (4 warps)
constexpr int kLStages = 3;   // cp.async ring depth
constexpr int kLBM = 64;      // rows of an m tile (the whole M when M <= 64)

struct LowProb {
    double* D;
    int pair_begin, pair_len;
};

struct LowArgs {
    const LowProb* probs;  // device; nullptr: the inline p0 / pair0
    const Pair* pairs;
    const int2* mtab;  // [ntile_m * BM]  {offset in L, offset in D}; -1 beyond M
    const int2* ntab;  // [Np]            {offset in R, offset in D}; -1 beyond N
    const int2* ktab;  // [K]             {offset in L, offset in R}; nullptr: k * ksL1, k * ksR1
    int nprob, M, N, K;
    int ksL1, ksR1;
    int BM, Np;      // rows of an m tile / N, both rounded up to 8
    int lda, ldb;    // leading dimensions of the staged chunks ([k][m], [k][n]); = 4 (mod 8): conflict-free fragment loads
    int KC;          // contracted elements per chunk, multiple of 4
    int cpp;         // chunks per operand pair = ceil(K / KC)
    int ntile_m, nslice;
    int a_kfast, b_kfast;  // the operand's contiguous direction is contracted: lanes walk k (else m / n)
    int a_vec, b_vec;      // m-/n-contiguous operand fetched as 16-byte pairs
    int atomic;
    double alpha, beta;
    LowProb p0;
    Pair pair0;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpa8(double* s, const double* g, bool valid) {
    const int sz = valid ? 8 : 0;  // src-size 0: zero fill, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cpa16(double* s, const double* g, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// position of a CTA in its sequence of chunks: work item w, chunks [c, c_end) of it still to go
struct Cursor {
    int w, c, c_end, tile, pair_begin;
    double* D;
};
__device__ __forceinline__ void cursor_load(Cursor& q, const LowArgs& a, int w, int nwork) {
    for (;; w += gridDim.x) {  // skip empty slices (a chain shorter than the slice count)
        q.w = w;
        if (w >= nwork) { q.c = q.c_end = 0; return; }
        const int s = w % a.nslice, r = w / a.nslice;
        q.tile = r % a.ntile_m;
        const int p = r / a.ntile_m;
        LowProb pr;
        if (a.probs) {
            const int4 raw = __ldg(reinterpret_cast<const int4*>(a.probs + p));
            memcpy(&pr, &raw, sizeof(pr));
        } else {
            pr = a.p0;
        }
        const long long T = (long long)pr.pair_len * a.cpp;
        q.c = (int)(s * T / a.nslice);
        q.c_end = (int)((s + 1) * T / a.nslice);
        q.pair_begin = pr.pair_begin;
        q.D = pr.D;
        if (q.c < q.c_end) return;
    }
}

// One operand chunk global -> shared: element (k, x) of the chunk (x = row of the m tile / column n) goes to
// dst[k * ld + x].  rows = x extent that exists (others are never read into a stored result), kcnt = contracted elements
// that exist; the k tail up to the next multiple of 4 is zero-filled (a DMMA k-step reads 4).
__device__ __forceinline__ void issue_operand(double* dst, int ld, const double* __restrict__ base, const int2* __restrict__ xtab,
                                              int rows, const int2* __restrict__ ktab, int ksel, int kstride, int k0, int kcnt,
                                              bool kfast, bool vec, int tid) {
    const int kcnt4 = (kcnt + 3) & ~3;
    if (kfast) {  // consecutive lanes -> consecutive k
        int lanes = 4;
        while (lanes < kcnt4 && lanes < kLT) lanes <<= 1;
        const int f = tid & (lanes - 1), s0 = tid / lanes, step = kLT / lanes;
        for (int kk = f; kk < kcnt4; kk += lanes) {
            const bool kv = kk < kcnt;
            int ko = 0;
            if (kv) ko = ktab ? (ksel ? __ldg(&ktab[k0 + kk].y) : __ldg(&ktab[k0 + kk].x)) : (k0 + kk) * kstride;
            double* d = dst + kk * ld;
#pragma unroll 4
            for (int x = s0; x < rows; x += step) {
                const int xo = __ldg(&xtab[x].x);
                cpa8(d + x, base + (kv ? ko + xo : 0), kv);
            }
        }
    } else if (vec) {  // consecutive lanes -> consecutive PAIRS of x (16 bytes)
        const int pairs = (rows + 1) >> 1;
        int lanes = 4;
        while (lanes < pairs && lanes < kLT) lanes <<= 1;
        const int f = (tid & (lanes - 1)) * 2, s0 = tid / lanes, step = kLT / lanes;
        for (int x = f; x < rows; x += 2 * lanes) {
            const int xo = __ldg(&xtab[x].x);
#pragma unroll 4
            for (int kk = s0; kk < kcnt4; kk += step) {
                const bool kv = kk < kcnt;
                int ko = 0;
                if (kv) ko = ktab ? (ksel ? __ldg(&ktab[k0 + kk].y) : __ldg(&ktab[k0 + kk].x)) : (k0 + kk) * kstride;
                cpa16(dst + kk * ld + x, base + (kv ? ko + xo : 0), kv);
            }
        }
    } else {  // consecutive lanes -> consecutive x
        int lanes = 4;
        while (lanes < rows && lanes < kLT) lanes <<= 1;
        const int f = tid & (lanes - 1), s0 = tid / lanes, step = kLT / lanes;
        for (int x = f; x < rows; x += lanes) {
            const int xo = __ldg(&xtab[x].x);
#pragma unroll 4
            for (int kk = s0; kk < kcnt4; kk += step) {
                const bool kv = kk < kcnt;
                int ko = 0;
                if (kv) ko = ktab ? (ksel ? __ldg(&ktab[k0 + kk].y) : __ldg(&ktab[k0 + kk].x)) : (k0 + kk) * kstride;
                cpa8(dst + kk * ld + x, base + (kv ? ko + xo : 0), kv);
            }
        }
    }
}

// MF x NF fragments of 8 x 8 per warp, the 4 warps arranged WR x (4 / WR) over the fragment grid of the tile.
template <int MF, int NF, int WR>
__global__ void __launch_bounds__(kLT) lowint_kernel(const __grid_constant__ LowArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int mrow0 = (warp % WR) * MF * 8, ncol0 = (warp / WR) * NF * 8;
    const int stage_elems = a.KC * (a.lda + a.ldb);
    const int nwork = a.nprob * a.ntile_m * a.nslice;
    // fragments of this warp that exist in the (padded) tile; uniform per warp
    const int mf_n = max(0, min(MF, (a.BM - mrow0) >> 3)), nf_n = max(0, min(NF, (a.Np - ncol0) >> 3));

    Cursor pf, cp;  // prefetch cursor (kLStages - 1 chunks ahead) and compute cursor walk the same sequence
    cursor_load(pf, a, blockIdx.x, nwork);
    cp = pf;

    auto issue = [&](int stage) {
        if (pf.w < nwork) {
            const int pair = pf.c / a.cpp, kc = pf.c - pair * a.cpp;
            const int k0 = kc * a.KC, kcnt = min(a.KC, a.K - k0);
            Pair pq;
            if (a.probs) {
                const int4 raw = __ldg(reinterpret_cast<const int4*>(a.pairs + pf.pair_begin + pair));
                memcpy(&pq, &raw, sizeof(pq));
            } else {
                pq = a.pair0;
            }
            double* As = sm + (size_t)stage * stage_elems;
            double* Bs = As + a.KC * a.lda;
            const int rows = min(a.BM, a.M - pf.tile * kLBM);
            issue_operand(As, a.lda, pq.L, a.mtab + pf.tile * a.BM, rows, a.ktab, 0, a.ksL1, k0, kcnt, a.a_kfast, a.a_vec, tid);
            issue_operand(Bs, a.ldb, pq.R, a.ntab, a.N, a.ktab, 1, a.ksR1, k0, kcnt, a.b_kfast, a.b_vec, tid);
            if (++pf.c == pf.c_end) cursor_load(pf, a, pf.w + gridDim.x, nwork);
        }
        cp_commit();  // always: the group count stays in step with the iteration count
    };

#pragma unroll
    for (int s = 0; s < kLStages - 1; ++s) issue(s);

    double acc[MF][NF][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const bool unit_alpha = a.alpha == 1.0;
    for (int it = 0; cp.w < nwork; ++it) {
        cp_wait<kLStages - 2>();
        __syncthreads();  // chunk `it` has landed for every thread; everybody is done with the stage refilled next
        issue((it + kLStages - 1) % kLStages);

        const int kc = cp.c % a.cpp;
        const int kcnt = min(a.KC, a.K - kc * a.KC);
        const double* As = sm + (size_t)(it % kLStages) * stage_elems + t4 * a.lda + mrow0 + g;
        const double* Bs = sm + (size_t)(it % kLStages) * stage_elems + a.KC * a.lda + t4 * a.ldb + ncol0 + g;
        const int nks = (kcnt + 3) >> 2;
        if (mf_n > 0 && nf_n > 0) {
#pragma unroll 2
            for (int ks = 0; ks < nks; ++ks) {
                double af[MF], bf[NF];
#pragma unroll
                for (int i = 0; i < MF; ++i)
                    if (i < mf_n) af[i] = As[ks * 4 * a.lda + i * 8];
#pragma unroll
                for (int j = 0; j < NF; ++j)
                    if (j < nf_n) bf[j] = Bs[ks * 4 * a.ldb + j * 8];
#pragma unroll
                for (int i = 0; i < MF; ++i)
#pragma unroll
                    for (int j = 0; j < NF; ++j)
                        if (i < mf_n && j < nf_n) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        if (++cp.c == cp.c_end) {
            // ---- epilogue of the item: D[perm(m, n)] (+)= alpha * acc, the output permute as a scatter ----
            const int2* mt = a.mtab + cp.tile * a.BM;
            const int rows = min(a.BM, a.M - cp.tile * kLBM);
            double* __restrict__ Dp = cp.D;
#pragma unroll
            for (int i = 0; i < MF; ++i) {
                const int row = mrow0 + i * 8 + g;
                if (i < mf_n && row < rows) {
                    const int mo = __ldg(&mt[row].y);
#pragma unroll
                    for (int j = 0; j < NF; ++j) {
                        if (j < nf_n) {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const int col = ncol0 + j * 8 + 2 * t4 + c;
                                if (col < a.N) {
                                    double* dst = Dp + (size_t)(mo + __ldg(&a.ntab[col].y));
                                    const double v = unit_alpha ? acc[i][j][c] : a.alpha * acc[i][j][c];
                                    if (a.atomic) asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(dst), "d"(v) : "memory");
                                    else if (a.beta == 0.0) *dst = v;
                                    else *dst = v + a.beta * *dst;
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            cursor_load(cp, a, cp.w + gridDim.x, nwork);
        }
    }
    cp_wait<0>();
}

// M = N = 1: D (+)= alpha * sum_k L[kL(k)] * R[kR(k)] over a chain of pairs; work item = (destination, slice of chain x K)
struct DotArgs {
    const LowProb* probs;
    const Pair* pairs;
    const int2* ktab;
    int nprob, K, ksL1, ksR1, nslice, atomic;
    double alpha, beta;
    LowProb p0;
    Pair pair0;
};
constexpr int kDotT = 256;
__global__ void __launch_bounds__(kDotT) dotk_kernel(const __grid_constant__ DotArgs a) {
    __shared__ double red[kDotT / 32];
    const int tid = threadIdx.x;
    const int nwork = a.nprob * a.nslice;
    const bool contig = !a.ktab && a.ksL1 == 1 && a.ksR1 == 1;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int s = w % a.nslice, p = w / a.nslice;
        const LowProb pr = a.probs ? a.probs[p] : a.p0;
        const long long T = (long long)pr.pair_len * a.K;  // elements of the whole chain
        long long e0 = s * T / a.nslice, e1 = (s + 1) * T / a.nslice;
        e0 &= ~1LL;  // slices start on even elements (16-byte loads); the last slice ends at T
        if (s + 1 < a.nslice) e1 &= ~1LL;
        double sum = 0.0;
        while (e0 < e1) {
            const int pair = (int)(e0 / a.K);
            const int k0 = (int)(e0 - (long long)pair * a.K);
            const int k1 = (int)min((long long)a.K, k0 + (e1 - e0));
            const Pair pq = a.probs ? a.pairs[pr.pair_begin + pair] : a.pair0;
            const double* __restrict__ Lp = pq.L;
            const double* __restrict__ Rp = pq.R;
            if (contig && ((((uintptr_t)Lp | (uintptr_t)Rp) & 15) == 0) && !(k0 & 1)) {
                const int n2 = (k1 - k0) >> 1;
                const double2* L2 = reinterpret_cast<const double2*>(Lp + k0);
                const double2* R2 = reinterpret_cast<const double2*>(Rp + k0);
                double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                int i = tid;
                for (; i + 3 * kDotT < n2; i += 4 * kDotT) {
                    const double2 x0 = __ldg(L2 + i), x1 = __ldg(L2 + i + kDotT), x2 = __ldg(L2 + i + 2 * kDotT), x3 = __ldg(L2 + i + 3 * kDotT);
                    const double2 y0 = __ldg(R2 + i), y1 = __ldg(R2 + i + kDotT), y2 = __ldg(R2 + i + 2 * kDotT), y3 = __ldg(R2 + i + 3 * kDotT);
                    s0 += x0.x * y0.x + x0.y * y0.y;
                    s1 += x1.x * y1.x + x1.y * y1.y;
                    s2 += x2.x * y2.x + x2.y * y2.y;
                    s3 += x3.x * y3.x + x3.y * y3.y;
                }
                for (; i < n2; i += kDotT) {
                    const double2 x = __ldg(L2 + i), y = __ldg(R2 + i);
                    s0 += x.x * y.x + x.y * y.y;
                }
                if (tid == 0 && ((k1 - k0) & 1)) s1 += Lp[k1 - 1] * Rp[k1 - 1];
                sum += (s0 + s1) + (s2 + s3);
            } else {
                double s0 = 0, s1 = 0;
                int k = k0 + tid;
                for (; k + kDotT < k1; k += 2 * kDotT) {
                    const int2 o0 = a.ktab ? __ldg(a.ktab + k) : make_int2(k * a.ksL1, k * a.ksR1);
                    const int2 o1 = a.ktab ? __ldg(a.ktab + k + kDotT) : make_int2((k + kDotT) * a.ksL1, (k + kDotT) * a.ksR1);
                    s0 += __ldg(Lp + o0.x) * __ldg(Rp + o0.y);
                    s1 += __ldg(Lp + o1.x) * __ldg(Rp + o1.y);
                }
                if (k < k1) {
                    const int2 o = a.ktab ? __ldg(a.ktab + k) : make_int2(k * a.ksL1, k * a.ksR1);
                    s0 += __ldg(Lp + o.x) * __ldg(Rp + o.y);
                }
                sum += s0 + s1;
            }
            e0 += k1 - k0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if ((tid & 31) == 0) red[tid >> 5] = sum;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < kDotT / 32; ++i) t += red[i];
            t *= a.alpha;
            if (a.atomic) asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(pr.D), "d"(t) : "memory");
            else pr.D[0] = a.beta == 0.0 ? t : t + a.beta * pr.D[0];
        }
        __syncthreads();
    }
}

// ---- host: offset tables per shape, cached on the device ----
struct Tables {
    int2 *mtab = nullptr, *ntab = nullptr, *ktab = nullptr;
    int BM = 0, Np = 0, ntile_m = 0;
};
std::unordered_map<std::string, Tables>& table_cache() {
    static std::unordered_map<std::string, Tables> c;
    return c;
}

void offsets_of(int lin, int nd, const int* ext, const int* s0, const int* s1, int& o0, int& o1) {
    o0 = o1 = 0;
    for (int i = 0; i < nd; ++i) {
        const int r = lin % ext[i];
        lin /= ext[i];
        o0 += r * s0[i];
        o1 += r * s1[i];
    }
}

int get_tables(const Shape& s, Tables* out) {
    std::string key(reinterpret_cast<const char*>(&s), offsetof(Shape, a_kc));  // M, N, K and the three index groups
    auto it = table_cache().find(key);
    if (it != table_cache().end()) { *out = it->second; return SIPGPU_OK; }
    Tables t;
    t.BM = s.M >= kLBM ? kLBM : (s.M + 7) & ~7;
    t.Np = (s.N + 7) & ~7;
    t.ntile_m = (s.M + kLBM - 1) / kLBM;
    const bool need_k = !(s.nk <= 1);
    std::vector<int2> h((size_t)t.ntile_m * t.BM + t.Np + (need_k ? s.K : 0));
    int2* hm = h.data();
    int2* hn = hm + (size_t)t.ntile_m * t.BM;
    int2* hk = hn + t.Np;
    for (int tile = 0; tile < t.ntile_m; ++tile)
        for (int i = 0; i < t.BM; ++i) {
            const int m = tile * kLBM + i;
            int oL = -1, oD = -1;
            if (m < s.M) offsets_of(m, s.nm, s.mext, s.msL, s.msD, oL, oD);
            hm[(size_t)tile * t.BM + i] = make_int2(oL, oD);
        }
    for (int n = 0; n < t.Np; ++n) {
        int oR = -1, oD = -1;
        if (n < s.N) offsets_of(n, s.nn, s.next, s.nsR, s.nsD, oR, oD);
        hn[n] = make_int2(oR, oD);
    }
    if (need_k)
        for (int k = 0; k < s.K; ++k) {
            int oL, oR;
            offsets_of(k, s.nk, s.kext, s.ksL, s.ksR, oL, oR);
            hk[k] = make_int2(oL, oR);
        }
    int2* d = reinterpret_cast<int2*>(pool_alloc(sizeof(int2) * h.size()));
    if (!d) return SIPGPU_E_NOMEM;
    // built once per shape; a blocking upload keeps the host vector simple (as permute.cu's plan tables)
    SIP_CUDA(cudaMemcpyAsync(d, h.data(), sizeof(int2) * h.size(), cudaMemcpyHostToDevice, ctx().stream));
    SIP_CUDA(cudaStreamSynchronize(ctx().stream));
    t.mtab = d;
    t.ntab = d + (size_t)t.ntile_m * t.BM;
    t.ktab = need_k ? t.ntab + t.Np : nullptr;
    table_cache().emplace(key, t);
    *out = t;
    return SIPGPU_OK;
}

template <int MF, int NF, int WR>
int launch_low(const LowArgs& a, int grid, size_t smem) {
    static bool attr_set = false;
    if (!attr_set) {
        SIP_CUDA(cudaFuncSetAttribute(lowint_kernel<MF, NF, WR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    lowint_kernel<MF, NF, WR><<<grid, kLT, smem, ctx().stream>>>(a);
    SIP_CUDA(cudaGetLastError());
    count_launch();
    return SIPGPU_OK;
}

}  // namespace

void lowint_cache_clear() { table_cache().clear(); }

static double& lowint_max_intensity() {
    static double v = [] {
        const char* on = getenv("SIPGPU_LOWINT");
        if (on && atoi(on) == 0) return -1.0;
        const char* e = getenv("SIPGPU_LOWINT_MAXI");
        return e ? atof(e) : 7.0;
    }();
    return v;
}
void lowint_set_max_intensity(double v) { lowint_max_intensity() = v; }

// Is this (launcher-oriented: the small free dimension already on the n side) contraction one for the bandwidth-shaped
// kernel?  N fits one tile and the flops per algorithmic byte stay below the ridge of the roofline
// (37 TFLOP/s / 6.5 TB/s = 5.7); everything denser belongs to the 128-wide tiles of contract.cu.
bool lowint_eligible(const Shape& s) {
    const double maxi = lowint_max_intensity();
    if (maxi < 0 || s.N > 64 || s.M < 1 || s.N < 1 || s.K < 1) return false;
    const double flops = 2.0 * s.M * s.N * s.K, bytes = 8.0 * ((double)s.M * s.K + (double)s.N * s.K + (double)s.M * s.N);
    return flops / bytes <= maxi;
}

// n destinations of ONE shape: destination i = alpha * sum over pairs [chain[i], chain[i+1]) + beta * D_i.
// pairs are already in kernel orientation ({R, L} when the shape is swapped).  dense_d: every D_i is a dense M x N
// block (split-K pre-scales it as one contiguous run); otherwise no split.
int lowint_launch(const Shape& s, int n, const std::vector<Pair>& pairs, const std::vector<int>& chain, double* const* D,
                  double alpha, double beta, bool dense_d) {
    if (n <= 0) return SIPGPU_OK;
    SIP_TRY(ensure_init());
    Ctx& c = ctx();
    const bool is_dot = s.M == 1 && s.N == 1;
    Tables t;
    if (!is_dot || s.nk > 1) SIP_TRY(get_tables(s, &t));
    long long longest = 0, total_pairs = 0;
    for (int i = 0; i < n; ++i) {
        longest = std::max<long long>(longest, chain[i + 1] - chain[i]);
        total_pairs += chain[i + 1] - chain[i];
    }
    // descriptors
    std::vector<LowProb> probs((size_t)n);
    for (int i = 0; i < n; ++i) probs[i] = LowProb{D[i], chain[i], chain[i + 1] - chain[i]};
    const LowProb* d_probs = nullptr;
    const Pair* d_pairs = nullptr;
    if (n > 1 || pairs.size() > 1) {
        const size_t b_probs = (sizeof(LowProb) * probs.size() + 255) & ~(size_t)255, b_pairs = sizeof(Pair) * pairs.size();
        void *h, *d;
        SIP_TRY(scratch_reserve(b_probs + b_pairs, &h, &d));
        memcpy(h, probs.data(), sizeof(LowProb) * probs.size());
        memcpy((char*)h + b_probs, pairs.data(), b_pairs);
        SIP_CUDA(cudaMemcpyAsync(d, h, b_probs + b_pairs, cudaMemcpyHostToDevice, c.stream));
        d_probs = (const LowProb*)d;
        d_pairs = (const Pair*)((char*)d + b_probs);
    }
    auto prescale = [&]() -> int {  // split partial sums meet through red.add: beta is applied once, up front
        if (beta == 1.0) return SIPGPU_OK;
        std::vector<double*> dd(D, D + n);
        std::vector<long long> cnt((size_t)n, (long long)s.M * s.N);
        return ew_scale_many(n, dd.data(), cnt.data(), beta);
    };

    if (is_dot) {
        DotArgs a;
        memset(&a, 0, sizeof(a));
        a.probs = d_probs; a.pairs = d_pairs; a.ktab = s.nk > 1 ? t.ktab : nullptr;
        a.nprob = n; a.K = s.K;
        a.ksL1 = s.nk >= 1 ? s.ksL[0] : 0; a.ksR1 = s.nk >= 1 ? s.ksR[0] : 0;
        a.alpha = alpha; a.beta = beta;
        a.p0 = probs[0]; a.pair0 = pairs[0];
        const long long slots = (long long)c.num_sms * 8;
        long long ns = 1;
        if (n < slots && dense_d) {  // at least 32 K elements per slice
            ns = std::min<long long>((slots + n - 1) / n, std::max<long long>(1, longest * s.K / 32768));
        }
        a.nslice = (int)ns;
        a.atomic = ns > 1;
        if (a.atomic) SIP_TRY(prescale());
        const long long nwork = (long long)n * ns;
        dotk_kernel<<<(int)std::min<long long>(nwork, slots), kDotT, 0, c.stream>>>(a);
        SIP_CUDA(cudaGetLastError());
        count_launch();
        return SIPGPU_OK;
    }

    LowArgs a;
    memset(&a, 0, sizeof(a));
    a.probs = d_probs; a.pairs = d_pairs;
    a.mtab = t.mtab; a.ntab = t.ntab; a.ktab = t.ktab;
    a.nprob = n; a.M = s.M; a.N = s.N; a.K = s.K;
    a.ksL1 = s.nk >= 1 ? s.ksL[0] : 0; a.ksR1 = s.nk >= 1 ? s.ksR[0] : 0;
    a.BM = t.BM; a.Np = t.Np; a.ntile_m = t.ntile_m;
    a.lda = a.BM + 4; a.ldb = a.Np + 4;
    const int perk = 8 * (a.lda + a.ldb);       // staged bytes per contracted element
    const int k4 = (s.K + 3) & ~3;
    int KC = (24 * 1024 / perk) & ~3;           // <= 24 KB per stage ...
    if ((long long)k4 * perk <= 32 * 1024) KC = k4;  // ... unless the whole contracted range fits 32 KB: one chunk per pair
    KC = std::max(4, std::min(KC, std::min(k4, 512)));
    a.KC = KC;
    a.cpp = (s.K + KC - 1) / KC;
    a.a_kfast = s.nk >= 1 && s.ksL[0] == 1 && s.kext[0] > 1;
    a.b_kfast = s.nk >= 1 && s.ksR[0] == 1 && s.kext[0] > 1;
    // 16-byte pairs along m / n: unit stride, even extent (pairs never straddle a dimension), every other stride even
    auto even_strides = [](int nd, const int* st) { for (int i = 1; i < nd; ++i) if (st[i] & 1) return false; return true; };
    bool ka_even = true, kb_even = true;
    for (int i = 0; i < s.nk; ++i) { ka_even = ka_even && !(s.ksL[i] & 1); kb_even = kb_even && !(s.ksR[i] & 1); }
    a.a_vec = !a.a_kfast && s.nm >= 1 && s.msL[0] == 1 && s.mext[0] % 2 == 0 && even_strides(s.nm, s.msL) && ka_even;
    a.b_vec = !a.b_kfast && s.nn >= 1 && s.nsR[0] == 1 && s.next[0] % 2 == 0 && even_strides(s.nn, s.nsR) && kb_even;
    if (a.a_vec || a.b_vec)
        for (const Pair& p : pairs) {
            if (a.a_vec && ((uintptr_t)p.L & 15)) a.a_vec = 0;
            if (a.b_vec && ((uintptr_t)p.R & 15)) a.b_vec = 0;
        }
    a.alpha = alpha; a.beta = beta;
    a.p0 = probs[0]; a.pair0 = pairs[0];

    const size_t smem = (size_t)kLStages * KC * perk;
    // resident CTAs per SM: 3 by registers (~155 x 128 threads), fewer when the ring is large
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(3, (200 * 1024) / (smem + 1024)));
    const long long slots = (long long)c.num_sms * per_sm;
    const long long nwork0 = (long long)n * t.ntile_m;
    long long ns = 1;
    if (nwork0 < slots && dense_d) {
        // every slice streams at least ~256 KB of operands (a split costs a pre-scale launch and atomics)
        const long long chunk_bytes = (long long)KC * 8 * (std::min(s.M, kLBM) + s.N);
        const long long min_chunks = std::max<long long>(1, (256 * 1024) / std::max<long long>(1, chunk_bytes));
        const long long avg_chunks = std::max<long long>(1, total_pairs * a.cpp / n);
        ns = std::min<long long>((2 * slots + nwork0 - 1) / nwork0, std::max<long long>(1, avg_chunks / min_chunks));
    }
    a.nslice = (int)ns;
    a.atomic = ns > 1;
    if (a.atomic) SIP_TRY(prescale());
    const long long nwork = nwork0 * ns;
    if (nwork >= (1LL << 31)) return SIPGPU_E_ARG;
    const int grid = (int)std::min<long long>(nwork, slots);
    // warp arrangement over the fragment grid: the one with the fewest fragments on its busiest warp
    const int fm = a.BM / 8, fn = a.Np / 8;
    auto busiest = [&](int wr, int mf, int nf) {
        const int wc = 4 / wr;
        if (wr * mf < fm || wc * nf < fn) return 1 << 20;  // does not cover the tile
        int worst = 0;
        for (int w = 0; w < 4; ++w) {
            const int r = std::max(0, std::min(mf, fm - (w % wr) * mf)), q = std::max(0, std::min(nf, fn - (w / wr) * nf));
            worst = std::max(worst, r * q);
        }
        return worst;
    };
    const int c22 = busiest(2, 4, 4), c14 = busiest(1, 8, 2), c41 = busiest(4, 2, 8);
    if (c14 < c22 && c14 <= c41) return launch_low<8, 2, 1>(a, grid, smem);
    if (c41 < c22) return launch_low<2, 8, 4>(a, grid, smem);
    return launch_low<4, 4, 2>(a, grid, smem);
}

}  // namespace sipgpu

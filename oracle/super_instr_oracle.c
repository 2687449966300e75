/*
 * super_instr_oracle.c -- CPU ORACLE (test infrastructure, NOT product code) for the elementwise CC
 * super-instructions that sit between contractions in the reference's pardo bodies.
 *
 * Plain-C restatement of
 *   src/sip/super_instructions/qm/qm-generic/energy_denominator_rhf.F:15-460
 *   src/sip/super_instructions/qm/utility/stripi.F, anti_symm_o.F, anti_symm_v.F, return_sval.F,
 *   invert_diagonal.F, invert_diagonal_asym.F, return_diagonal_elements.F, qm-generic/energy_ty_denominator_rhf.F
 * with the reference's super-instruction calling convention reduced to what the arithmetic reads:
 * (rank, index_values, extents, data) per argument (special_instructions.h:27-97), the predefined int
 * array "moa_seg_ranges" passed explicitly instead of through the sip_interface upcall.
 * Index values are 1-based segment numbers; "global" orbital indices are 1-based.
 *
 * PARITY PINNING: the reference has no known-answer test for these routines other than
 * return_sval_test.sialx (test_basic_sial.cpp); the restatement is checked against closed-form numpy
 * statements in tests/test_super_instr_cpu.py -> "parity unpinned" at the reference level for this file,
 * except energy_denominator_rhf, which is pinned through the reference's energy goldens
 * (tests/test_lccd_water_energy_cpu.py: rank 4 by LCCD, rank 2 by the LCCSD / CCSD singles update, rank 6 by (T)).
 */
#include <string.h>

/* offset of segment `index_value` = sum of the extents of the segments before it (energy_denominator_rhf.F:66-90) */
static int seg_offset(const int* seg, int index_value) {
    int off = 0;
    for (int i = 1; i <= index_value - 1; ++i) off += seg[i - 1];
    return off;
}

/* energy_denominator_rhf.F:15-247 + do_rhf_den2/_1d/4/_1d/6: array(a,b,..) /= eps, eps = sum of the Fock diagonal at
 * the EVEN positions (b, d, f) minus the sum at the ODD positions (a, c, e).  fock_rank 2: fock(p,p), leading
 * dimension fock_ld; fock_rank 1: fock(p).  Rank 6 only exists with a rank-2 Fock array and treats extent-1
 * dimensions as simple indices (offset = index_value - 1, F:166-219).  Returns ierr (0 ok, 1 unsupported). */
int oracle_si_energy_denominator_rhf(int rank, const int* index_values, const int* ext, double* data, int fock_rank,
                                     int fock_ld, const double* fock, const int* moa_seg_ranges) {
    if (!(rank == 2 || rank == 4 || rank == 6) || !(fock_rank == 1 || fock_rank == 2) || (rank == 6 && fock_rank != 2))
        return 1;
    int off[6];
    for (int d = 0; d < rank; ++d) {
        if (rank == 6 && ext[d] == 1) off[d] = index_values[d] - 1;
        else off[d] = seg_offset(moa_seg_ranges, index_values[d]);
    }
    long long n = 1;
    for (int d = 0; d < rank; ++d) n *= ext[d];
    for (long long lin = 0; lin < n; ++lin) {
        long long r = lin;
        double eps = 0.0;
        /* the reference adds in the order epsb + epsd (+ epsf) - epsa - epsc (- epse) */
        double e[6];
        for (int d = 0; d < rank; ++d) {
            const int p = (int)(r % ext[d]) + 1 + off[d];  /* global 1-based orbital index */
            r /= ext[d];
            e[d] = fock_rank == 2 ? fock[(size_t)(p - 1) * fock_ld + (p - 1)] : fock[p - 1];
        }
        if (rank == 2) eps = e[1] - e[0];
        else if (rank == 4) eps = e[1] + e[3] - e[0] - e[2];
        else eps = e[1] + e[3] + e[5] - e[0] - e[2] - e[4];
        data[lin] = data[lin] / eps;
    }
    return 0;
}

/* stripi.F: y(..., match, ...) = x(..., match, ...) in GLOBAL indices: every dimension of y with extent 1 is a
 * "stripped" dimension whose single global index is  index_values_1(rank)  for the LAST dimension (the simple index)
 * and  1 + offset  for the others; it must lie inside x's range (else the reference aborts -> ierr 2).  Ranks 2-4. */
int oracle_si_stripi(int rank, const int* iv0, const int* ext0, const double* x, const int* iv1, const int* ext1,
                     double* y, const int* moa_seg_ranges) {
    if (rank < 2 || rank > 4) return 1;
    int beg[4], cnt[4];
    for (int d = 0; d < rank; ++d) {
        const int x1 = 1 + seg_offset(moa_seg_ranges, iv0[d]);
        const int x2 = ext0[d] + x1 - 1;
        if (ext1[d] == 1 || d == rank - 1) {
            const int match = d == rank - 1 ? iv1[d] : 1 + seg_offset(moa_seg_ranges, iv1[d]);
            if (match < x1 || match > x2) return 2;
            beg[d] = match - x1;
            cnt[d] = 1;
        } else {
            if (ext1[d] != ext0[d]) return 3;
            beg[d] = 0;
            cnt[d] = ext0[d];
        }
    }
    long long n = 1;
    for (int d = 0; d < rank; ++d) n *= cnt[d];
    for (long long lin = 0; lin < n; ++lin) {
        long long r = lin, xo = 0, xs = 1;
        for (int d = 0; d < rank; ++d) {
            const int i = (int)(r % cnt[d]);
            r /= cnt[d];
            xo += (long long)(i + beg[d]) * xs;
            xs *= ext0[d];
        }
        y[lin] = x[xo];
    }
    return 0;
}

/* anti_symm_o.F dosymmforce4o: x(a,j,b,i) = -x(a,i,b,j) for global i < j, then x = 0 where i == j or a == b.
 * (Pairs whose mirror element lies outside the block are skipped; the reference would read/write out of bounds.) */
int oracle_si_anti_symm_o(int rank, const int* iv, const int* ext, double* x, const int* moa_seg_ranges) {
    if (rank != 4) return 1;
    int o[4];
    for (int d = 0; d < 4; ++d) o[d] = seg_offset(moa_seg_ranges, iv[d]);
    const long long s1 = ext[0], s2 = s1 * ext[1], s3 = s2 * ext[2];
    for (int b = 0; b < ext[2]; ++b)
        for (int a = 0; a < ext[0]; ++a)
            for (int j = 0; j < ext[3]; ++j)
                for (int i = 0; i < ext[1]; ++i) {
                    const int gi = i + 1 + o[1], gj = j + 1 + o[3];
                    if (gi < gj) {
                        const int mi = gj - 1 - o[1], mj = gi - 1 - o[3];  /* local position of (j, i) */
                        if (mi >= 0 && mi < ext[1] && mj >= 0 && mj < ext[3]) x[a + mi * s1 + b * s2 + mj * s3] = x[a + i * s1 + b * s2 + j * s3] * (-1.0);
                    }
                }
    for (int b = 0; b < ext[2]; ++b)
        for (int a = 0; a < ext[0]; ++a)
            for (int j = 0; j < ext[3]; ++j)
                for (int i = 0; i < ext[1]; ++i)
                    if (i + o[1] == j + o[3] || a + o[0] == b + o[2]) x[a + i * s1 + b * s2 + j * s3] = 0.0;
    return 0;
}

/* anti_symm_v.F dosymmforce4v: x(b,i,a,j) = -x(a,i,b,j) for global a < b; general case then zeroes i == j or a == b;
 * the simple-index special case (i range == 1:1) writes -0.0 on a == b and does not zero i == j. */
int oracle_si_anti_symm_v(int rank, const int* iv, const int* ext, double* x, const int* moa_seg_ranges) {
    if (rank != 4) return 1;
    int o[4];
    for (int d = 0; d < 4; ++d) o[d] = seg_offset(moa_seg_ranges, iv[d]);
    const long long s1 = ext[0], s2 = s1 * ext[1], s3 = s2 * ext[2];
    const int special = (o[1] + 1 == 1 && o[1] + ext[1] == 1);
    for (int b = 0; b < ext[2]; ++b)
        for (int a = 0; a < ext[0]; ++a)
            for (int j = 0; j < ext[3]; ++j)
                for (int i = 0; i < ext[1]; ++i) {
                    const int ga = a + 1 + o[0], gb = b + 1 + o[2];
                    const int ma = gb - 1 - o[0], mb = ga - 1 - o[2];  /* local position of (b, a) */
                    const int inside = ma >= 0 && ma < ext[0] && mb >= 0 && mb < ext[2];
                    if (ga < gb && inside) x[ma + i * s1 + mb * s2 + j * s3] = x[a + i * s1 + b * s2 + j * s3] * (-1.0);
                    if (special && ga == gb && inside) x[ma + i * s1 + mb * s2 + j * s3] = 0.0 * (-1.0);
                }
    if (special) return 0;
    for (int b = 0; b < ext[2]; ++b)
        for (int a = 0; a < ext[0]; ++a)
            for (int j = 0; j < ext[3]; ++j)
                for (int i = 0; i < ext[1]; ++i)
                    if (i + o[1] == j + o[3] || a + o[0] == b + o[2]) x[a + i * s1 + b * s2 + j * s3] = 0.0;
    return 0;
}

/* return_sval.F: the LAST element of a rank-1/2 block -> scalar */
int oracle_si_return_sval(int rank, const int* ext, const double* data, double* sval) {
    if (rank < 1 || rank > 2) return 1;
    long long n = 1;
    for (int d = 0; d < rank; ++d) n *= ext[d];
    *sval = data[n - 1];
    return 0;
}

/* invert_diagonal.F: array1 /= array2 where array2 != 0; ranks 3 and 5 only */
int oracle_si_invert_diagonal(int rank0, int rank1, const int* ext, double* a1, const double* a2) {
    if (rank0 != rank1 || !(rank0 == 3 || rank0 == 5)) return 1;
    long long n = 1;
    for (int d = 0; d < rank0; ++d) n *= ext[d];
    for (long long i = 0; i < n; ++i)
        if (a2[i] != 0.0) a1[i] = a1[i] / a2[i];
    return 0;
}

/* return_diagonal_elements.F ret_diag_tensor2 / ret_diag_tensor4: keep x(p,p) (rank 2) or x(p,p,r,r) (rank 4), zero the rest.
 * The Fortran declares BOTH dimensions of a pair over the range of the FIRST one (tensor(n1:n2,n1:n2),
 * tensor(a1:a2,a1:a2,b1:b2,b1:b2)), so "diagonal" compares positions inside the block, whatever segments the pair of
 * indices carries; the block must be square per pair (abort otherwise: ierr 1). */
int oracle_si_return_diagonal_elements(int rank, const int* iv, const int* ext, double* x, const int* moa_seg_ranges) {
    (void)iv; (void)moa_seg_ranges;
    if (rank == 2) {
        if (ext[0] != ext[1]) return 1;
        for (int q = 0; q < ext[1]; ++q)
            for (int p = 0; p < ext[0]; ++p)
                if (p != q) x[p + (long long)q * ext[0]] = 0.0;
        return 0;
    }
    if (rank == 4) {
        if (ext[0] != ext[1] || ext[2] != ext[3]) return 1;
        const long long s1 = ext[0], s2 = s1 * ext[1], s3 = s2 * ext[2];
        for (int s = 0; s < ext[3]; ++s)
            for (int r = 0; r < ext[2]; ++r)
                for (int q = 0; q < ext[1]; ++q)
                    for (int p = 0; p < ext[0]; ++p)
                        if (!(p == q && r == s)) x[p + q * s1 + r * s2 + s * s3] = 0.0;
        return 0;
    }
    return 1;
}

/* invert_diagonal_asym.F do_return_inv5_as: rank 5 (a, b, c, d, e), b..e carry segment offsets: where global b != d and
 * c != e, array1 /= array2 (skipped where array2 == 0); everywhere else array1 = 0. */
int oracle_si_invert_diagonal_asym(int rank0, int rank1, const int* iv, const int* ext, double* a1, const double* a2,
                                   const int* moa_seg_ranges) {
    if (rank0 != rank1 || rank0 != 5) return 1;
    int o[5] = {0, 0, 0, 0, 0};
    for (int d = 1; d < 5; ++d) o[d] = seg_offset(moa_seg_ranges, iv[d]);
    long long i = 0;
    for (int e = 0; e < ext[4]; ++e)
        for (int d = 0; d < ext[3]; ++d)
            for (int c = 0; c < ext[2]; ++c)
                for (int b = 0; b < ext[1]; ++b)
                    for (int a = 0; a < ext[0]; ++a, ++i) {
                        if (b + o[1] != d + o[3] && c + o[2] != e + o[4]) {
                            if (a2[i] != 0.0) a1[i] = a1[i] / a2[i];
                        } else {
                            a1[i] = 0.0;
                        }
                    }
    return 0;
}

/* energy_ty_denominator_rhf.F do_rhfty_den4: rank 4 only, rank-2 Fock array: array(a,b,c,d) /= epsb + epsd - epsa - epsc + shift
 * (the shifted denominator of the CIS(D) program, rcis_d_rhf.sialx:1175) */
int oracle_si_energy_ty_denominator_rhf(int rank, const int* index_values, const int* ext, double* data, int fock_ld,
                                        const double* fock, double shift, const int* moa_seg_ranges) {
    if (rank != 4) return 1;
    int off[4];
    for (int d = 0; d < 4; ++d) off[d] = seg_offset(moa_seg_ranges, index_values[d]);
    long long lin = 0;
    for (int d = 0; d < ext[3]; ++d)
        for (int c = 0; c < ext[2]; ++c)
            for (int b = 0; b < ext[1]; ++b)
                for (int a = 0; a < ext[0]; ++a, ++lin) {
                    const double epsa = fock[(size_t)(a + off[0]) * fock_ld + a + off[0]], epsb = fock[(size_t)(b + off[1]) * fock_ld + b + off[1]];
                    const double epsc = fock[(size_t)(c + off[2]) * fock_ld + c + off[2]], epsd = fock[(size_t)(d + off[3]) * fock_ld + d + off[3]];
                    const double eps = epsb + epsd - epsa - epsc + shift;
                    data[lin] = data[lin] / eps;
                }
    return 0;
}

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
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "contract.h"
#include "elementwise.h"

namespace sipgpu {
namespace {

constexpr int kLT = 128;      // threads per CTA (4 warps)
constexpr int kLStages = 4;   // cp.async ring depth
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
    int BM;          // table rows per m tile (rows of a full tile rounded up to 8)
    // staged chunk of an operand: element (k, x) at k * sk + x * sx.  An operand whose contiguous direction is a free
    // index is laid out [k][x] (sx = 1, sk = ld), one whose contiguous direction is contracted [x][k] (sk = 1, sx = ld):
    // consecutive lanes of a cp.async instruction then write consecutive shared-memory words, and with ld = 4 (mod 8)
    // doubles the DMMA fragment loads are conflict-free in both layouts
    int a_sk, a_sx, b_sk, b_sx;
    int a_elems, stage_elems;
    int KC;          // contracted elements per chunk, multiple of 4
    int cpp;         // chunks per operand pair = ceil(K / KC)
    int ntile_m, nslice;
    int a_kfast, b_kfast;  // lanes walk k (else m / n)
    int a_vec, b_vec;      // 16-byte items along the walking direction
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

// position of a CTA in its sequence of chunks: work item w, chunks [c, c_end) of it still to go; (pair, kc) = c split
// into operand pair and chunk inside the pair, kept incrementally (no division per chunk)
struct Cursor {
    int w, c, c_end, tile, pair, kc;
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
        const int pr0 = q.c / a.cpp;
        q.kc = q.c - pr0 * a.cpp;
        q.pair = pr.pair_begin + pr0;
        q.D = pr.D;
        if (q.c < q.c_end) return;
    }
}
__device__ __forceinline__ void cursor_next_chunk(Cursor& q, const LowArgs& a) {
    ++q.c;
    if (++q.kc == a.cpp) { q.kc = 0; ++q.pair; }
}

// how the 128 threads of a CTA cover one operand chunk: the walking direction (k or x) across `lanes` consecutive
// threads, the other direction across the 128 / lanes groups; fixed for the whole kernel
struct LaneMap {
    int f, s0, lanes, step;
};
__device__ __forceinline__ LaneMap lane_map(int walk_extent, bool vec, int tid) {
    const int items = vec ? (walk_extent + 1) >> 1 : walk_extent;
    int sh = 2;
    while ((1 << sh) < items && sh < 7) ++sh;
    LaneMap m;
    m.lanes = 1 << sh;
    m.f = (tid & (m.lanes - 1)) * (vec ? 2 : 1);
    m.s0 = tid >> sh;
    m.step = kLT >> sh;
    return m;
}

// One operand chunk global -> shared: element (k, x) of the chunk (x = row of the m tile / column n) goes to
// dst[k * sk + x * sx].  rows = x extent that exists (other rows are never read into a stored result), kcnt = contracted
// elements that exist; the k tail up to the next multiple of 4 is zero-filled (a DMMA k-step reads 4).
// KFAST: consecutive lanes walk k (else x); VEC: 16-byte items (pairs along the walking direction); KTAB: k offsets from
// the table (several contracted dimensions) instead of k * kstride.  The inner loops carry no branch on these.
template <bool VEC>
__device__ __forceinline__ void cpa(double* s, const double* g, bool valid) {
    if constexpr (VEC) cpa16(s, g, valid); else cpa8(s, g, valid);
}
template <bool KFAST, bool VEC, bool KTAB>
__device__ __forceinline__ void issue_loop(double* dst, int sk, int sx, const double* __restrict__ base, const int2* __restrict__ xtab,
                                           int rows, const int* __restrict__ ktab, int kstride, int k0, int kcnt, const LaneMap& lm) {
    const int kcnt4 = (kcnt + 3) & ~3;
    const int adv = lm.lanes * (VEC ? 2 : 1);
    if constexpr (KFAST) {
        for (int kk = lm.f; kk < kcnt4; kk += adv) {
            const bool kv = kk < kcnt;
            const int ko = !kv ? 0 : (KTAB ? __ldg(ktab + 2 * (k0 + kk)) : (k0 + kk) * kstride);
            const double* g = base + ko;
            double* d = dst + kk * sk + lm.s0 * sx;
            const int2* xt = xtab + lm.s0;
            const int dsx = lm.step * sx;
#pragma unroll 4
            for (int x = lm.s0; x < rows; x += lm.step, d += dsx, xt += lm.step) cpa<VEC>(d, g + (kv ? __ldg(&xt->x) : 0), kv);
        }
    } else {
        for (int x = lm.f; x < rows; x += adv) {
            const double* g = base + __ldg(&xtab[x].x);
            double* d = dst + x * sx + lm.s0 * sk;
            const int dsk = lm.step * sk;
            int kk = lm.s0;
            if constexpr (KTAB) {
                const int* kt = ktab + 2 * (k0 + kk);
#pragma unroll 4
                for (; kk < kcnt; kk += lm.step, d += dsk, kt += 2 * lm.step) cpa<VEC>(d, g + __ldg(kt), true);
            } else {
                int ko = (k0 + kk) * kstride;
                const int dko = lm.step * kstride;
#pragma unroll 4
                for (; kk < kcnt; kk += lm.step, d += dsk, ko += dko) cpa<VEC>(d, g + ko, true);
            }
            for (; kk < kcnt4; kk += lm.step, d += dsk) cpa<VEC>(d, g, false);  // zero fill of the k tail
        }
    }
}
// ktab points at the .x (L) or .y (R) component of the {offset in L, offset in R} table, stride 2 ints
__device__ __forceinline__ void issue_operand(double* dst, int sk, int sx, const double* __restrict__ base,
                                              const int2* __restrict__ xtab, int rows, const int* __restrict__ ktab, int kstride,
                                              int k0, int kcnt, int mode, const LaneMap& lm) {
    switch (mode) {  // (kfast << 2) | (vec << 1) | ktab -- uniform for the whole kernel
        case 0: issue_loop<false, false, false>(dst, sk, sx, base, xtab, rows, ktab, kstride, k0, kcnt, lm); break;
        case 1: issue_loop<false, false, true>(dst, sk, sx, base, xtab, rows, ktab, kstride, k0, kcnt, lm); break;
        case 2: issue_loop<false, true, false>(dst, sk, sx, base, xtab, rows, ktab, kstride, k0, kcnt, lm); break;
        case 3: issue_loop<false, true, true>(dst, sk, sx, base, xtab, rows, ktab, kstride, k0, kcnt, lm); break;
        case 4: issue_loop<true, false, false>(dst, sk, sx, base, xtab, rows, ktab, kstride, k0, kcnt, lm); break;
        case 5: issue_loop<true, false, true>(dst, sk, sx, base, xtab, rows, ktab, kstride, k0, kcnt, lm); break;
        case 6: issue_loop<true, true, false>(dst, sk, sx, base, xtab, rows, ktab, kstride, k0, kcnt, lm); break;
        default: issue_loop<true, true, true>(dst, sk, sx, base, xtab, rows, ktab, kstride, k0, kcnt, lm); break;
    }
}

// Every warp owns exactly MF x NF fragments of 8 x 8 (compile time: the DMMA loop carries no guards); the 4 warps are
// arranged WR x (4 / WR) over the tile.  A tile smaller than the warps' footprint leaves rows / columns of the staged chunk
// unwritten: their products land in accumulators that are never stored.
template <int MF, int NF, int WR>
__global__ void __launch_bounds__(kLT, (MF * NF <= 8) ? 4 : 3) lowint_kernel(const __grid_constant__ LowArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int mrow0 = (warp % WR) * MF * 8, ncol0 = (warp / WR) * NF * 8;
    const int nwork = a.nprob * a.ntile_m * a.nslice;
    const LaneMap lma = lane_map(a.a_kfast ? a.KC : a.BM, a.a_vec, tid);
    const LaneMap lmb = lane_map(a.b_kfast ? a.KC : a.N, a.b_vec, tid);
    // fragment addressing: A(k, m) of the tile at As[k * a_sk + m * a_sx]; a lane reads (k = 4 ks + t4, m = mrow0 + 8 i + g)
    const int a_thr = t4 * a.a_sk + (mrow0 + g) * a.a_sx, b_thr = a.a_elems + t4 * a.b_sk + (ncol0 + g) * a.b_sx;
    const int a_i = 8 * a.a_sx, b_j = 8 * a.b_sx, a_ks = 4 * a.a_sk, b_ks = 4 * a.b_sk;

    const int* ktab_i = reinterpret_cast<const int*>(a.ktab);
    const int a_mode = (a.a_kfast << 2) | (a.a_vec << 1) | (a.ktab ? 1 : 0), b_mode = (a.b_kfast << 2) | (a.b_vec << 1) | (a.ktab ? 1 : 0);

    Cursor pf, cp;  // prefetch cursor (kLStages - 1 chunks ahead) and compute cursor walk the same sequence
    cursor_load(pf, a, blockIdx.x, nwork);
    cp = pf;

    double acc[MF][NF][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const bool unit_alpha = a.alpha == 1.0;
    for (int it = -(kLStages - 1); cp.w < nwork; ++it) {
        if (it >= 0) {
            cp_wait<kLStages - 2>();
            __syncthreads();  // chunk `it` has landed for every thread; everybody is done with the stage refilled next
        }
        {   // ---- prefetch: the chunk kLStages - 1 ahead (the ring runs across item boundaries) ----
            if (pf.w < nwork) {
                const int stage = (it + kLStages - 1) % kLStages;
                const int k0 = pf.kc * a.KC, kcnt = min(a.KC, a.K - k0);
                Pair pq;
                if (a.probs) {
                    const int4 raw = __ldg(reinterpret_cast<const int4*>(a.pairs + pf.pair));
                    memcpy(&pq, &raw, sizeof(pq));
                } else {
                    pq = a.pair0;
                }
                double* As = sm + (size_t)stage * a.stage_elems;
                const int rows = min(a.BM, a.M - pf.tile * kLBM);
                issue_operand(As, a.a_sk, a.a_sx, pq.L, a.mtab + pf.tile * a.BM, rows, ktab_i, a.ksL1, k0, kcnt, a_mode, lma);
                issue_operand(As + a.a_elems, a.b_sk, a.b_sx, pq.R, a.ntab, a.N, ktab_i ? ktab_i + 1 : nullptr, a.ksR1, k0, kcnt, b_mode, lmb);
                cursor_next_chunk(pf, a);
                if (pf.c == pf.c_end) cursor_load(pf, a, pf.w + gridDim.x, nwork);
            }
            cp_commit();  // always: the group count stays in step with the iteration count
        }
        if (it < 0) continue;

        const int kcnt = min(a.KC, a.K - cp.kc * a.KC);
        const double* st = sm + (size_t)(it % kLStages) * a.stage_elems;
        const double* Ap = st + a_thr;
        const double* Bp = st + b_thr;
        const int nks = (kcnt + 3) >> 2;
#pragma unroll 2
        for (int ks = 0; ks < nks; ++ks) {
            double af[MF], bf[NF];
#pragma unroll
            for (int i = 0; i < MF; ++i) af[i] = Ap[i * a_i];
#pragma unroll
            for (int j = 0; j < NF; ++j) bf[j] = Bp[j * b_j];
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int j = 0; j < NF; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            Ap += a_ks;
            Bp += b_ks;
        }
        cursor_next_chunk(cp, a);
        if (cp.c == cp.c_end) {
            // ---- epilogue of the item: D[perm(m, n)] (+)= alpha * acc, the output permute as a scatter ----
            const int2* mt = a.mtab + cp.tile * a.BM;
            const int rows = min(a.BM, a.M - cp.tile * kLBM);
            double* __restrict__ Dp = cp.D;
            int no[NF][2];
#pragma unroll
            for (int j = 0; j < NF; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = ncol0 + j * 8 + 2 * t4 + c;
                    no[j][c] = col < a.N ? __ldg(&a.ntab[col].y) : -1;
                }
#pragma unroll
            for (int i = 0; i < MF; ++i) {
                const int row = mrow0 + i * 8 + g;
                const int mo = row < rows ? __ldg(&mt[row].y) : -1;
#pragma unroll
                for (int j = 0; j < NF; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        if ((mo | no[j][c]) >= 0) {
                            double* dst = Dp + (size_t)(mo + no[j][c]);
                            const double v = unit_alpha ? acc[i][j][c] : a.alpha * acc[i][j][c];
                            if (a.atomic) asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(dst), "d"(v) : "memory");
                            else if (a.beta == 0.0) *dst = v;
                            else *dst = v + a.beta * *dst;
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

// ---- slab path: small result, long contracted range, operands that arrive as CONTIGUOUS SLABS ----------------------------
// D[a,b] = L[a,c,d,e] * R[b,c,d,e] and its relatives (rank-2 results of rank-4 blocks, M, N <= 64, K = 10^4..10^5) sit on the
// ridge of the roofline: the kernel above spends ~90 instructions per cp.async on table-driven gathers (ncu: 20 instructions
// per DMMA, `wait` + `selected` stalls, DMMA pipe 36 %).  When a chunk of KC values of ONE contracted dimension q, taken
// together with all the free dimensions, is a handful of contiguous runs in BOTH operands (L[a,c | d,e]: 50 x 20 doubles in
// one run), the chunk is moved by TMA: one cp.async.bulk per run (<= 32 runs, one per lane of warp 0) into a ring of stages,
// completion on the stage's mbarrier -- the staged layout IS the memory layout, the fragment loads address it through a
// per-row offset kept in registers and a constant k stride.  No table look-ups and no address arithmetic in the loop:
// 4 LDS-sized loads + the DMMAs.  Small tiles (<= 3 x 7 fragments) give every warp the whole tile and a quarter of the k
// steps (reduced through shared memory in a fixed order at the end of an item); larger ones split N over the warps.
constexpr int kSlabMaxRuns = 32, kSlabMaxStages = 6;
struct SlabArgs {
    const LowProb* probs;
    const Pair* pairs;
    const int2* mtab;   // .y = offset in D of row m
    const int2* ntab;   // .y = offset in D of column n
    const int2* ctab;   // [cpp] {offset in L, offset in R} of a chunk
    int nprob, M, N, KC, cpp, nslice, atomic, stages;
    int a_sk, b_sk;            // stride of k inside a staged operand
    int a_elems, stage_elems;  // doubles (both even)
    int a_nruns, b_nruns, a_runlen, b_runlen;
    int a_run[kSlabMaxRuns], b_run[kSlabMaxRuns];  // offset of run r inside the chunk of the operand (elements)
    short a_moff[64], b_noff[64];                  // staged offset of row m / column n (0 beyond M / N)
    // hybrid: an operand whose chunk is NOT a few long runs (R[b,e,d,c] next to L[a,c,d,e]: 400-byte pieces) is gathered by
    // the 128 threads instead -- one cp.async per table item {offset in the chunk, staged offset}, the table built once per
    // shape in the operand's memory order; its completion arrives on the same mbarrier (cp.async.mbarrier.arrive.noinc)
    const int2* a_items;   // nullptr: operand A travels as TMA runs
    const int2* b_items;
    int a_nitems, b_nitems, a_vec, b_vec;          // vec: 16-byte items
    double alpha, beta;
    LowProb p0;
    Pair pair0;
};

struct SlabCursor {
    int w, c, c_end, pair, kc;
    double* D;
};
__device__ __forceinline__ void slab_cursor_load(SlabCursor& q, const SlabArgs& a, int w, int nwork) {
    for (;; w += gridDim.x) {
        q.w = w;
        if (w >= nwork) { q.c = q.c_end = 0; return; }
        const int s = w % a.nslice, p = w / a.nslice;
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
        const int pr0 = q.c / a.cpp;
        q.kc = q.c - pr0 * a.cpp;
        q.pair = pr.pair_begin + pr0;
        q.D = pr.D;
        if (q.c < q.c_end) return;
    }
}
__device__ __forceinline__ void slab_cursor_next(SlabCursor& q, const SlabArgs& a) {
    ++q.c;
    if (++q.kc == a.cpp) { q.kc = 0; ++q.pair; }
}

// GATHER: one operand is gathered by the threads (hybrid).  A separate instantiation: with the gather code compiled into the
// all-TMA kernel the K-split variants lost 15-20 % (0.63 -> 0.77 ms on 50 x 20 results), although none of it executes there.
// NW: warps per CTA.  4, or 1 for tiny matrices (K <= 64: every item is ONE chunk, so a CTA-wide barrier and the cross-warp sum
// per item would cost more than the item; a single warp owns the whole tile and only ever waits for its own copies)
template <int MF, int NF, bool KSPLIT, bool GATHER, int NW>
__global__ void __launch_bounds__(32 * NW) slab_kernel(const __grid_constant__ SlabArgs a) {
    static_assert(NW == 4 || (NW == 1 && !KSPLIT), "one warp holds the whole tile");
    extern __shared__ __align__(16) double sm[];
    __shared__ unsigned long long full[kSlabMaxStages];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int ncol0 = KSPLIT ? 0 : warp * NF * 8;
    const int nwork = a.nprob * a.nslice;
    double* red = sm + (size_t)a.stages * a.stage_elems;   // KSPLIT: MF * NF * 64 doubles for the reduction over the warps
    int moff[MF], noff[NF];
#pragma unroll
    for (int i = 0; i < MF; ++i) moff[i] = a.a_moff[8 * i + g];
#pragma unroll
    for (int j = 0; j < NF; ++j) noff[j] = a.a_elems + a.b_noff[ncol0 + 8 * j + g];
    constexpr bool gather = GATHER;
    if (tid == 0) {
        const int count = gather ? 1 + 32 * NW : 1;   // the expect_tx arrival (+ one asynchronous arrival per gathering thread)
        for (int s = 0; s < a.stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(full + s)), "r"(count) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    SlabCursor pf, cp;
    slab_cursor_load(cp, a, blockIdx.x, nwork);
    pf = cp;
    const unsigned stage_bytes = 8u * (unsigned)((a.a_items ? 0 : a.a_nruns * a.a_runlen) + (a.b_items ? 0 : a.b_nruns * a.b_runlen));
    auto issue = [&](int stage) {   // the chunk at the prefetch cursor -> stage (warp 0: TMA runs; everybody: gathers); cursor moves on
        if (pf.w >= nwork) return;
        Pair pq;
        if (a.probs) {
            const int4 raw = __ldg(reinterpret_cast<const int4*>(a.pairs + pf.pair));
            memcpy(&pq, &raw, sizeof(pq));
        } else {
            pq = a.pair0;
        }
        const int2 co = __ldg(a.ctab + pf.kc);
        const unsigned bar = (unsigned)__cvta_generic_to_shared(full + stage);
        double* st = sm + (size_t)stage * a.stage_elems;
        if constexpr (GATHER) {
            if (a.a_items) {
                const double* base = pq.L + co.x;
                for (int i = tid; i < a.a_nitems; i += 32 * NW) {
                    const int2 e = __ldg(a.a_items + i);
                    if (a.a_vec) cpa16(st + e.y, base + e.x, true); else cpa8(st + e.y, base + e.x, true);
                }
            }
            if (a.b_items) {
                const double* base = pq.R + co.y;
                for (int i = tid; i < a.b_nitems; i += 32 * NW) {
                    const int2 e = __ldg(a.b_items + i);
                    if (a.b_vec) cpa16(st + a.a_elems + e.y, base + e.x, true); else cpa8(st + a.a_elems + e.y, base + e.x, true);
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
        }
        if (!GATHER || warp == 0) {
        if (lane == 0) asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(bar), "r"(stage_bytes) : "memory");
        __syncwarp();
        if (!a.a_items && lane < a.a_nruns)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                             (unsigned)__cvta_generic_to_shared(st + lane * a.a_runlen)),
                         "l"(pq.L + co.x + a.a_run[lane]), "r"(8u * (unsigned)a.a_runlen), "r"(bar)
                         : "memory");
        if (!a.b_items && lane < a.b_nruns)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                             (unsigned)__cvta_generic_to_shared(st + a.a_elems + lane * a.b_runlen)),
                         "l"(pq.R + co.y + a.b_run[lane]), "r"(8u * (unsigned)a.b_runlen), "r"(bar)
                         : "memory");
        }
        slab_cursor_next(pf, a);
        if (pf.c == pf.c_end) slab_cursor_load(pf, a, pf.w + gridDim.x, nwork);
    };
    if (GATHER || warp == 0)
        for (int s = 0; s < a.stages; ++s) issue(s);

    double acc[MF][NF][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nks = (a.KC + 3) >> 2;
    const bool unit_alpha = a.alpha == 1.0;
    int stage = 0;
    unsigned parity = 0;
    int rot = warp;   // KSPLIT: the k steps are dealt round-robin over the warps, continuing across chunks
    while (cp.w < nwork) {
        {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(full + stage);
            asm volatile(
                "{\n .reg .pred p;\n"
                "WS: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                " @!p bra WS;\n}\n" ::"r"(bar), "r"(parity)
                : "memory");
        }
        const double* st = sm + (size_t)stage * a.stage_elems;
        for (int ks = KSPLIT ? rot : 0; ks < nks; ks += KSPLIT ? 4 : 1) {
            const int k = 4 * ks + t4;
            const bool kv = k < a.KC;
            const double* Ap = st + k * a.a_sk;
            const double* Bp = st + k * a.b_sk;
            double af[MF], bf[NF];
#pragma unroll
            for (int i = 0; i < MF; ++i) af[i] = kv ? Ap[moff[i]] : 0.0;
#pragma unroll
            for (int j = 0; j < NF; ++j) bf[j] = kv ? Bp[noff[j]] : 0.0;
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int j = 0; j < NF; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        if (KSPLIT) rot = (rot + 4 - (nks & 3)) & 3;
        slab_cursor_next(cp, a);
        const bool item_done = cp.c == cp.c_end;
        __syncthreads();   // every warp is done with the stage
        if (GATHER || warp == 0) issue(stage);
        if (++stage == a.stages) { stage = 0; parity ^= 1; }
        if (!item_done) continue;
        // ---- end of the item: (KSPLIT: sum the four warps' partial tiles in a fixed order) and write D ----
        if (KSPLIT) {
            for (int w = 1; w < 4; ++w) {
                if (warp == w) {
#pragma unroll
                    for (int i = 0; i < MF; ++i)
#pragma unroll
                        for (int j = 0; j < NF; ++j)
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                double* r = red + ((i * NF + j) * 2 + c) * 32 + lane;
                                *r = w == 1 ? acc[i][j][c] : *r + acc[i][j][c];
                            }
                }
                __syncthreads();
            }
        }
        if (!KSPLIT || warp == 0) {
            double* __restrict__ Dp = cp.D;
#pragma unroll
            for (int i = 0; i < MF; ++i) {
                const int row = 8 * i + g;
                const int mo = row < a.M ? __ldg(&a.mtab[row].y) : -1;
#pragma unroll
                for (int j = 0; j < NF; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int col = ncol0 + 8 * j + 2 * t4 + c;
                        const int no = col < a.N ? __ldg(&a.ntab[col].y) : -1;
                        if ((mo | no) < 0) continue;
                        double v = acc[i][j][c];
                        if (KSPLIT) v += red[((i * NF + j) * 2 + c) * 32 + lane];
                        if (!unit_alpha) v *= a.alpha;
                        double* dst = Dp + (size_t)(mo + no);
                        if (a.atomic) asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(dst), "d"(v) : "memory");
                        else if (a.beta == 0.0) *dst = v;
                        else *dst = v + a.beta * *dst;
                    }
            }
        }
        if (KSPLIT) __syncthreads();   // the reduction buffer is free again
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
            for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        slab_cursor_load(cp, a, cp.w + gridDim.x, nwork);
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
    int2* d = reinterpret_cast<int2*>(pool_alloc(sizeof(int2) * h.size(), true));
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
int launch_low(LowArgs& a, size_t smem, long long nwork0, bool may_split, long long total_chunks, long long chunk_bytes,
               const std::function<int()>& prescale) {
    static int per_sm = 0;
    auto* kern = lowint_kernel<MF, NF, WR>;
    if (!per_sm) {
        SIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        // several CTAs per SM: ask for the full shared-memory carve-out
        SIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        per_sm = -1;
    }
    int occ = 1;
    SIP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kLT, smem));
    if (occ < 1) occ = 1;
    const long long slots = (long long)ctx().num_sms * occ;
    long long ns = 1;
    if (nwork0 < slots && may_split) {
        // every slice streams at least ~256 KB of operands (a split costs a pre-scale launch and atomics)
        const long long min_chunks = std::max<long long>(1, (256 * 1024) / std::max<long long>(1, chunk_bytes));
        const long long avg_chunks = std::max<long long>(1, total_chunks / std::max<long long>(1, a.nprob));
        // the slice count whose items fill whole waves of CTA slots best (items are dealt round-robin)
        const long long cap = std::min<long long>(std::max<long long>(1, avg_chunks / min_chunks), (4 * slots + nwork0 - 1) / nwork0);
        double best = -1.0;
        for (long long k = 1; k <= cap; ++k) {
            const long long items = nwork0 * k, waves = (items + slots - 1) / slots;
            const double fill = (double)items / (double)(waves * slots) - 0.002 * k;
            if (fill > best) { best = fill; ns = k; }
        }
    }
    a.nslice = (int)ns;
    a.atomic = ns > 1;
    if (a.atomic) SIP_TRY(prescale());
    const long long nwork = nwork0 * ns;
    if (nwork >= (1LL << 31)) return SIPGPU_E_ARG;
    const int grid = (int)std::min<long long>(nwork, slots);
    auto go = [a, grid, smem, kern]() -> int {
        kern<<<grid, kLT, smem, ctx().stream>>>(a);
        SIP_CUDA(cudaGetLastError());
        count_launch();
        return SIPGPU_OK;
    };
    if (Capture* cap = capture()) { cap->steps.push_back(go); return SIPGPU_OK; }
    return go();
}


// ---- slab path, host side ----
struct SlabPlan {
    bool ok = false;
    int q = -1, KC = 0, cpp = 0;
    bool tiny = false;   // K <= 64: one chunk per item, single-warp CTAs
    int a_sk = 0, b_sk = 0, a_elems = 0, b_elems = 0;
    int a_nruns = 0, b_nruns = 0, a_runlen = 0, b_runlen = 0;
    int a_run[kSlabMaxRuns], b_run[kSlabMaxRuns];
    short a_moff[64], b_noff[64];
    int2* ctab = nullptr;   // device
    // hybrid: the operand that is gathered by the threads (host copy of its item table until the plan is chosen)
    std::vector<int2> a_items_h, b_items_h;
    int2 *a_items = nullptr, *b_items = nullptr;   // device
    int a_vec = 0, b_vec = 0;
};
struct SlabDim {
    int ext, gs;   // extent inside a chunk, stride in the operand
    int kind;      // 0: free dimension i of the group, 1: the contracted dimension q
    int idx;
    int cs;        // compact (staged) stride
};
// one operand's view of a chunk: `nf` free dimensions (ext, stride) + KC values of a contracted dimension of stride qs and
// full extent qext.  Runs = the leading dimensions (in memory order) that are complete; everything else enumerates runs.
bool slab_operand(int nf, const int* fext, const int* fs, int qs, int qext, int KC, int X, int* runlen, int* nruns, int* run,
                  int* sk, short* xoff, int* elems) {
    SlabDim d[kMaxRank + 1];
    int n = 0;
    for (int i = 0; i < nf; ++i) d[n++] = SlabDim{fext[i], fs[i], 0, i, 0};
    d[n++] = SlabDim{KC, qs, 1, 0, 0};
    for (int i = 1; i < n; ++i)   // by memory stride (a handful of entries)
        for (int j = i; j > 0 && d[j].gs < d[j - 1].gs; --j) std::swap(d[j], d[j - 1]);
    long long cs = 1;
    for (int i = 0; i < n; ++i) { d[i].cs = (int)cs; cs *= d[i].ext; }
    if (cs > 32767) return false;   // staged offsets are shorts
    *elems = (int)cs;
    // the run: leading dimensions whose memory stride equals the staged stride (nothing of the block lies between)
    long long rl = 1;
    int lead = 0;
    for (; lead < n; ++lead) {
        if (d[lead].gs != d[lead].cs) break;
        rl *= d[lead].ext;
        if (d[lead].kind == 1 && KC != qext) { ++lead; break; }   // a partial contracted dimension ends the run
    }
    if (lead == 0 || (rl & 1) || rl < 128) return false;           // 16-byte granules; runs of at least 1 KB (a 400-byte bulk
                                                                   // copy costs the TMA unit as much as an 8 KB one: measured 2x slower)
    long long nr = 1;
    for (int i = lead; i < n; ++i) nr *= d[i].ext;
    if (nr > kSlabMaxRuns) return false;
    *runlen = (int)rl;
    *nruns = (int)nr;
    for (int r = 0; r < (int)nr; ++r) {   // run r: the non-leading dimensions counted in staged order
        int lin = r, off = 0;
        for (int i = lead; i < n; ++i) { off += (lin % d[i].ext) * d[i].gs; lin /= d[i].ext; }
        run[r] = off;
        if (off & 1) return false;        // 16-byte aligned sources
    }
    for (int i = 0; i < n; ++i)
        if (d[i].kind == 1) *sk = d[i].cs;
    for (int x = 0; x < 64; ++x) {
        int off = 0;
        if (x < X) {
            int lin = x;
            for (int i = 0; i < nf; ++i) {   // x enumerates the free group in the Shape's order
                const int r = lin % fext[i];
                lin /= fext[i];
                for (int j = 0; j < n; ++j)
                    if (d[j].kind == 0 && d[j].idx == i) off += r * d[j].cs;
            }
        }
        xoff[x] = (short)off;
    }
    return true;
}

// the other way to bring an operand's chunk in: the 128 threads gather it with cp.async, one table item {offset inside the
// chunk, staged offset} per 16-byte (or 8-byte) piece, walking the operand's memory order.  Staged [k][x] when a free index is
// the operand's fastest (x contiguous), [x][k] when the chunked contracted index is; the leading dimension is 4 (mod 8)
// doubles, so the fragment loads of a half-warp fall on 16 different banks.  false: neither is the fastest (8-byte pieces
// scattered over sectors -- leave the shape to the gather kernel).
bool slab_gather(int nf, const int* fext, const int* fs, int qs, int KC, int X, std::vector<int2>* items, int* vec, int* sk,
                 short* xoff, int* elems) {
    auto pad4 = [](int n) { int v = n; while (v % 8 != 4) ++v; return v; };
    std::vector<int> xg((size_t)X);
    for (int x = 0; x < X; ++x) {
        int lin = x, off = 0;
        for (int i = 0; i < nf; ++i) { off += (lin % fext[i]) * fs[i]; lin /= fext[i]; }
        xg[x] = off;
    }
    auto all_even = [&](int from) { for (int i = from; i < nf; ++i) if (fs[i] & 1) return false; return true; };
    items->clear();
    if (nf >= 1 && fs[0] == 1) {           // a free index is fastest: [k][x]
        const int ld = pad4((X + 7) & ~7);
        *vec = fext[0] % 2 == 0 && all_even(1) && qs % 2 == 0;
        for (int k = 0; k < KC; ++k)
            for (int x = 0; x < X; x += *vec ? 2 : 1) items->push_back(make_int2(k * qs + xg[x], k * ld + x));
        *sk = ld;
        for (int x = 0; x < 64; ++x) xoff[x] = (short)(x < X ? x : 0);
        *elems = KC * ld;
    } else if (qs == 1) {                  // the chunked contracted index is fastest: [x][k]
        const int ld = pad4(KC);
        *vec = KC % 2 == 0 && all_even(0);
        for (int x = 0; x < X; ++x)
            for (int k = 0; k < KC; k += *vec ? 2 : 1) items->push_back(make_int2(xg[x] + k, x * ld + k));
        *sk = 1;
        for (int x = 0; x < 64; ++x) xoff[x] = (short)(x < X ? x * ld : 0);
        *elems = X * ld;
    } else {
        return false;
    }
    return *elems <= 32767;
}

std::unordered_map<std::string, SlabPlan>& slab_cache() {
    static std::unordered_map<std::string, SlabPlan> c;
    return c;
}

int& slab_mode() {
    static int v = [] { const char* e = getenv("SIPGPU_LOWINT_SLAB"); return e ? atoi(e) : 1; }();
    return v;
}

// menu of warp arrangements: KSPLIT (every warp the whole tile, a quarter of the k steps) and N split over the 4 warps.
// (KSPLIT 7 x 7 for 50 x 50 was tried: 254 registers, one CTA per SM, every warp reading the whole staged tile -- 0.88 ms
// against 0.58 ms for the N split with its 8 padded columns.)
struct SlabVariant { int mf, nf; bool ksplit; int nw; };
const SlabVariant kSlabMenu[] = {{2, 2, true, 4}, {3, 3, true, 4}, {4, 4, true, 4}, {3, 7, true, 4}, {7, 3, true, 4},
                                 {5, 2, false, 4}, {6, 2, false, 4}, {7, 2, false, 4}, {8, 2, false, 4},
                                 {3, 3, false, 1}, {3, 7, false, 1}, {7, 3, false, 1}};   // tiny matrices: one warp per CTA
// (7 x 7 fragments in one warp: 1.68 ms against 0.51 ms for 16 384 blocks of 50 x 50 x 50 on the gather kernel -- not on the menu)
const SlabVariant* slab_variant(int M, int N, bool tiny = false) {
    const int mf = (M + 7) / 8, nf = (N + 7) / 8;
    const SlabVariant* best = nullptr;
    double best_cost = 1e30;
    for (const SlabVariant& v : kSlabMenu) {
        if ((v.nw == 1) != tiny) continue;
        const bool covers = (v.ksplit || v.nw == 1) ? (v.mf >= mf && v.nf >= nf) : (v.mf >= mf && 4 * v.nf >= nf);
        if (!covers) continue;
        const double cost = v.ksplit ? v.mf * v.nf / 4.0 : v.mf * v.nf;
        if (cost < best_cost) { best_cost = cost; best = &v; }
    }
    return best;
}

// the plan of a shape (cached): which contracted dimension is chunked and how; ok = false: not a slab shape
const SlabPlan& slab_plan(const Shape& s) {
    std::string key(reinterpret_cast<const char*>(&s), offsetof(Shape, a_kc));
    auto it = slab_cache().find(key);
    if (it != slab_cache().end()) return it->second;
    SlabPlan best;
    long long best_score = -1;
    static const int min_k = [] { const char* e = getenv("SIPGPU_SLAB_MINK"); return e ? atoi(e) : 1024; }();
    static const int hybrid = [] { const char* e = getenv("SIPGPU_SLAB_HYBRID"); return e ? atoi(e) : 1; }();
    static const long long chunk_cap = [] { const char* e = getenv("SIPGPU_SLAB_CHUNK_KB"); return (e ? atoll(e) : 24LL) * 1024; }();
    static const long long stage_cap = [] { const char* e = getenv("SIPGPU_SLAB_STAGE_KB"); return (e ? atoll(e) : 40LL) * 1024; }();
    // tiny matrices (D[a,b] = L[a,c]*R[c,b], a few hundred elements): measured against the gather kernel at 16 384 blocks per launch
    // 20 x 20 x 50: 118 vs 201 us, 20 x 50 x 20: 155-163 vs 204-287 us, 20 x 50 x 50: 304-310 vs 287-370 us (SIPGPU_SLAB_TINY=0: off)
    static const int tiny_on = [] { const char* e = getenv("SIPGPU_SLAB_TINY"); return e ? atoi(e) : 1; }();
    const bool tiny = tiny_on > 0 && s.K < min_k && s.nk == 1 && s.K >= 8 && s.K <= 64;   // one chunk per item, one warp per CTA
    if (s.M <= 64 && s.N <= 64 && s.M >= 2 && s.N >= 2 && (s.K >= min_k || tiny) && slab_variant(s.M, s.N, tiny)) {
        const long long perk = 8LL * (s.M + s.N);
        for (int q = 0; q < s.nk; ++q) {
            for (int KC = std::min(s.kext[q], 128); KC >= 8; --KC) {
                if (s.kext[q] % KC) continue;
                if (tiny && KC != s.K) continue;
                if (perk * KC > stage_cap) continue;
                SlabPlan p;
                p.q = q; p.KC = KC; p.tiny = tiny;
                const bool a_tma = slab_operand(s.nm, s.mext, s.msL, s.ksL[q], s.kext[q], KC, s.M, &p.a_runlen, &p.a_nruns, p.a_run, &p.a_sk, p.a_moff, &p.a_elems);
                const bool b_tma = slab_operand(s.nn, s.next, s.nsR, s.ksR[q], s.kext[q], KC, s.N, &p.b_runlen, &p.b_nruns, p.b_run, &p.b_sk, p.b_noff, &p.b_elems);
                if (!a_tma && !b_tma && hybrid < 2) continue;   // (2: both operands gathered through item tables -- measured
                                                               //  SLOWER than the gather kernel: 2.16 vs 1.31 ms, 1.95 vs 1.21 ms)
                if (!a_tma || !b_tma) {   // hybrid: the other operand is gathered by the threads
                    if (hybrid <= 0) continue;
                    if (!a_tma && !slab_gather(s.nm, s.mext, s.msL, s.ksL[q], KC, s.M, &p.a_items_h, &p.a_vec, &p.a_sk, p.a_moff, &p.a_elems)) continue;
                    if (!b_tma && !slab_gather(s.nn, s.next, s.nsR, s.ksR[q], KC, s.N, &p.b_items_h, &p.b_vec, &p.b_sk, p.b_noff, &p.b_elems)) continue;
                    if (!a_tma) { p.a_nruns = 0; p.a_runlen = 1 << 20; }
                    if (!b_tma) { p.b_nruns = 0; p.b_runlen = 1 << 20; }
                    if (8LL * (p.a_elems + p.b_elems) > 48 * 1024) continue;   // the padded staging of the gathered operand counts
                }
                // prefer both operands by TMA, long runs, then long chunks (fewer barriers per byte), then few wasted k steps
                const long long score = (a_tma && b_tma ? 8000000000LL : a_tma || b_tma ? 4000000000LL : 0) + (long long)std::min(std::min(p.a_runlen, p.b_runlen), 512) * 1000000 +
                                        (long long)std::min(perk * KC, chunk_cap) * 10 + (KC % 4 == 0 ? 1 : 0);
                if (score > best_score) { best_score = score; best = p; best.ok = true; }
            }
        }
    }
    if (best.ok) {
        best.cpp = s.K / best.KC;
        std::vector<int2> h((size_t)best.cpp);
        // chunk c: sub-range j of dimension q fastest, then the other contracted dimensions in the Shape's order
        const int nj = s.kext[best.q] / best.KC;
        for (int c = 0; c < best.cpp; ++c) {
            int lin = c;
            const int j = lin % nj;
            lin /= nj;
            int oL = j * best.KC * s.ksL[best.q], oR = j * best.KC * s.ksR[best.q];
            for (int i = 0; i < s.nk; ++i) {
                if (i == best.q) continue;
                const int r = lin % s.kext[i];
                lin /= s.kext[i];
                oL += r * s.ksL[i];
                oR += r * s.ksR[i];
            }
            if ((oL | oR) & 1) { best.ok = false; break; }
            h[c] = make_int2(oL, oR);
        }
        if (best.ok) {
            auto upload = [&](const std::vector<int2>& v) -> int2* {
                int2* d = reinterpret_cast<int2*>(pool_alloc(sizeof(int2) * v.size(), true));
                if (!d || cudaMemcpyAsync(d, v.data(), sizeof(int2) * v.size(), cudaMemcpyHostToDevice, ctx().stream) != cudaSuccess ||
                    cudaStreamSynchronize(ctx().stream) != cudaSuccess) { best.ok = false; return nullptr; }
                return d;
            };
            best.ctab = upload(h);
            if (best.ok && !best.a_items_h.empty()) best.a_items = upload(best.a_items_h);
            if (best.ok && !best.b_items_h.empty()) best.b_items = upload(best.b_items_h);
        }
    }
    return slab_cache().emplace(key, best).first->second;
}

template <int MF, int NF, bool KSPLIT, bool GATHER, int NW>
int launch_slab(SlabArgs& a, size_t stage_bytes, bool may_split, long long total_chunks, const std::function<int()>& prescale) {
    auto* kern = slab_kernel<MF, NF, KSPLIT, GATHER, NW>;
    static bool attr = false;
    if (!attr) {
        SIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr = true;
    }
    const size_t red_bytes = KSPLIT ? (size_t)MF * NF * 64 * 8 : 0;
    // ring depth: two CTAs per SM when the stages are small (<= 96 KB per CTA), else one
    static const int ring_kb = [] { const char* e = getenv("SIPGPU_SLAB_RING_KB"); return e ? atoi(e) : 96; }();
    // three stages, not as many as fit: occupancy beats ring depth (20 x 20 results, 16 KB stages: 6 stages = 2 CTAs per SM
    // 0.85 ms, 3 stages = 4 CTAs per SM 0.63 ms = 6.3 TB/s; the 28-32 KB stages of the larger tiles get 3 either way)
    static const int max_st = [] { const char* e = getenv("SIPGPU_SLAB_MAX_STAGES"); return e ? std::min(atoi(e), kSlabMaxStages) : 3; }();
    int stages = (int)std::min<size_t>(max_st, ((size_t)ring_kb * 1024 - red_bytes) / stage_bytes);
    if (NW == 1) stages = 2;   // one chunk per item: the next item's chunk in flight is all there is to overlap
    if (stages < 3 && NW != 1) stages = (int)std::min<size_t>(kSlabMaxStages, (200 * 1024 - red_bytes) / stage_bytes);
    if (stages < 2) return SIPGPU_E_ARG;
    a.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + red_bytes;
    int occ = 1;
    SIP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * NW, smem));
    if (occ < 1) occ = 1;
    const long long slots = (long long)ctx().num_sms * occ;
    long long ns = 1;
    if (a.nprob < slots && may_split) {   // (a launch that fills the slots is never cut: its sums stay bit-reproducible)
        // cut along K so that the items fill whole waves of CTA slots (items are dealt round-robin: 600 items on 296 slots
        // take as long as 888): the slice count with the best fill, every slice streaming at least ~256 KB
        const long long min_chunks = std::max<long long>(1, (256 * 1024) / (long long)stage_bytes);
        const long long avg_chunks = std::max<long long>(1, total_chunks / std::max(1, a.nprob));
        const long long cap = std::min<long long>(std::max<long long>(1, avg_chunks / min_chunks), (6 * slots + a.nprob - 1) / a.nprob);
        double best = -1.0;
        for (long long k = 1; k <= cap; ++k) {
            const long long items = (long long)a.nprob * k, waves = (items + slots - 1) / slots;
            const double fill = (double)items / (double)(waves * slots) - 0.002 * k;   // ties: fewer slices
            if (fill > best) { best = fill; ns = k; }
        }
    }
    a.nslice = (int)ns;
    a.atomic = ns > 1;
    if (a.atomic) SIP_TRY(prescale());
    const long long nwork = (long long)a.nprob * ns;
    if (nwork >= (1LL << 31)) return SIPGPU_E_ARG;
    const int grid = (int)std::min<long long>(nwork, slots);
    auto go = [a, grid, smem, kern]() -> int {
        kern<<<grid, 32 * NW, smem, ctx().stream>>>(a);
        SIP_CUDA(cudaGetLastError());
        count_launch();
        return SIPGPU_OK;
    };
    if (Capture* cap = capture()) { cap->steps.push_back(go); return SIPGPU_OK; }
    return go();
}

// true + launched, or false: not a slab shape / operands not 16-byte aligned (the caller goes on with the gather kernel)
int slab_try(const Shape& s, const Tables& t, int n, const std::vector<Pair>& pairs, const LowProb* d_probs, const Pair* d_pairs,
             const std::vector<LowProb>& probs, long long total_pairs, double alpha, double beta, bool dense_d,
             const std::function<int()>& run_prescale, bool* done) {
    *done = false;
    if (slab_mode() <= 0) return SIPGPU_OK;
    const SlabPlan& p = slab_plan(s);
    if (!p.ok) return SIPGPU_OK;
    for (const Pair& pr : pairs)
        if (((uintptr_t)pr.L | (uintptr_t)pr.R) & 15) return SIPGPU_OK;
    const SlabVariant* v = slab_variant(s.M, s.N, p.tiny);
    SlabArgs a;
    memset(&a, 0, sizeof(a));
    a.probs = d_probs; a.pairs = d_pairs;
    a.mtab = t.mtab; a.ntab = t.ntab; a.ctab = p.ctab;
    a.nprob = n; a.M = s.M; a.N = s.N; a.KC = p.KC; a.cpp = p.cpp;
    a.a_sk = p.a_sk; a.b_sk = p.b_sk;
    a.a_elems = (p.a_elems + 1) & ~1;
    a.stage_elems = (a.a_elems + p.b_elems + 15) & ~15;
    a.a_nruns = p.a_nruns; a.b_nruns = p.b_nruns; a.a_runlen = p.a_runlen; a.b_runlen = p.b_runlen;
    memcpy(a.a_run, p.a_run, sizeof(a.a_run));
    memcpy(a.b_run, p.b_run, sizeof(a.b_run));
    memcpy(a.a_moff, p.a_moff, sizeof(a.a_moff));
    memcpy(a.b_noff, p.b_noff, sizeof(a.b_noff));
    a.a_items = p.a_items; a.b_items = p.b_items;
    a.a_nitems = (int)p.a_items_h.size(); a.b_nitems = (int)p.b_items_h.size();
    a.a_vec = p.a_vec; a.b_vec = p.b_vec;
    a.alpha = alpha; a.beta = beta;
    a.p0 = probs[0]; a.pair0 = pairs[0];
    const size_t stage_bytes = (size_t)a.stage_elems * 8;
    const long long total_chunks = total_pairs * p.cpp;
    *done = true;
#define SLAB_CASE(MF_, NF_, KS_, NW_) \
    if (v->mf == MF_ && v->nf == NF_ && v->ksplit == KS_ && v->nw == NW_)                                                            \
        return (a.a_items || a.b_items) ? launch_slab<MF_, NF_, KS_, true, NW_>(a, stage_bytes, dense_d, total_chunks, run_prescale) \
                                        : launch_slab<MF_, NF_, KS_, false, NW_>(a, stage_bytes, dense_d, total_chunks, run_prescale)
    SLAB_CASE(2, 2, true, 4); SLAB_CASE(3, 3, true, 4); SLAB_CASE(4, 4, true, 4); SLAB_CASE(3, 7, true, 4); SLAB_CASE(7, 3, true, 4);
    SLAB_CASE(5, 2, false, 4); SLAB_CASE(6, 2, false, 4); SLAB_CASE(7, 2, false, 4); SLAB_CASE(8, 2, false, 4);
    SLAB_CASE(3, 3, false, 1); SLAB_CASE(3, 7, false, 1); SLAB_CASE(7, 3, false, 1);
#undef SLAB_CASE
    *done = false;
    return SIPGPU_OK;
}

}  // namespace

void lowint_cache_clear() { table_cache().clear(); slab_cache().clear(); }
void lowint_set_slab(int on) { slab_mode() = on; }

static double& lowint_max_intensity() {
    static double v = [] {
        const char* on = getenv("SIPGPU_LOWINT");
        if (on && atoi(on) == 0) return -1.0;
        const char* e = getenv("SIPGPU_LOWINT_MAXI");
        return e ? atof(e) : 7.0;
    }();
    return v;
}
static int& lowint_scope() {
    static int v = [] { const char* e = getenv("SIPGPU_LOWINT_SCOPE"); return e ? atoi(e) : 1; }();
    return v;
}
void lowint_set_max_intensity(double v) { lowint_max_intensity() = v; }
void lowint_set_scope(int v) { lowint_scope() = v; }

// Is this (launcher-oriented: the small free dimension already on the n side) contraction one for the bandwidth-shaped
// kernel?  Scope 1 (default) is where it beats the 128-wide tiles on the B200 (profiles/r02_lowint_vs_tiles.txt, all 170
// SIAL patterns at o = 20, v = 50): dot products (M = N = 1: 5-7x), destinations that fit ONE tile with a short contracted
// range (tiny matrices: 1.2-1.4x) or with at most 512 elements (rank-2 results of rank-4 blocks, 20 x 20 / 50 x 1:
// 1.3-1.9x).  Matrix-vector shapes run level with the tiles and the M-tiled skinny shapes slower (one warp role issues the
// loads AND the DMMAs), so they stay with contract.cu.  Scope 2 (tests, A/B runs) takes everything with N <= 64 below
// `lowint_max_intensity` flops per algorithmic byte.
bool lowint_eligible(const Shape& s) {
    const double maxi = lowint_max_intensity();
    const int scope = lowint_scope();
    if (maxi < 0 || scope <= 0 || s.N > 64 || s.M < 1 || s.N < 1 || s.K < 1) return false;
    const double flops = 2.0 * s.M * s.N * s.K, bytes = 8.0 * ((double)s.M * s.K + (double)s.N * s.K + (double)s.M * s.N);
    if (flops / bytes > maxi) return false;
    if (scope >= 2) return true;
    if (slab_mode() > 0 && slab_plan(s).ok) return true;   // contiguous-slab shapes: TMA-fed, see slab_kernel
    return s.M <= 64 && ((long long)s.M * s.N <= 512 || s.K <= 64);
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
        SIP_TRY(desc_alloc(b_probs + b_pairs, &h, &d));
        memcpy(h, probs.data(), sizeof(LowProb) * probs.size());
        memcpy((char*)h + b_probs, pairs.data(), b_pairs);
        SIP_TRY(desc_commit(h, d, b_probs + b_pairs));
        d_probs = (const LowProb*)d;
        d_pairs = (const Pair*)((char*)d + b_probs);
    }
    // split partial sums meet through red.add: beta is applied once, up front (a closure that owns its pointer list, so
    // that a prepared launch can replay it)
    std::function<int()> prescale = [dd = std::vector<double*>(D, D + n), mn = (long long)s.M * s.N, beta]() -> int {
        if (beta == 1.0) return SIPGPU_OK;
        std::vector<long long> cnt(dd.size(), mn);
        return ew_scale_many((int)dd.size(), dd.data(), cnt.data(), beta);
    };
    auto run_prescale = [&]() -> int {
        if (Capture* cap = capture()) { cap->steps.push_back(prescale); return SIPGPU_OK; }
        return prescale();
    };

    if (!is_dot) {
        bool done = false;
        SIP_TRY(slab_try(s, t, n, pairs, d_probs, d_pairs, probs, total_pairs, alpha, beta, dense_d, run_prescale, &done));
        if (done) return SIPGPU_OK;
    }
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
        if (a.atomic) SIP_TRY(run_prescale());
        const int grid = (int)std::min<long long>((long long)n * ns, slots);
        auto go = [a, grid]() -> int {
            dotk_kernel<<<grid, kDotT, 0, ctx().stream>>>(a);
            SIP_CUDA(cudaGetLastError());
            count_launch();
            return SIPGPU_OK;
        };
        if (Capture* cap = capture()) { cap->steps.push_back(go); return SIPGPU_OK; }
        return go();
    }

    LowArgs a;
    memset(&a, 0, sizeof(a));
    a.probs = d_probs; a.pairs = d_pairs;
    a.mtab = t.mtab; a.ntab = t.ntab; a.ktab = t.ktab;
    a.nprob = n; a.M = s.M; a.N = s.N; a.K = s.K;
    a.ksL1 = s.nk >= 1 ? s.ksL[0] : 0; a.ksR1 = s.nk >= 1 ? s.ksR[0] : 0;
    a.BM = t.BM; a.ntile_m = t.ntile_m;
    a.alpha = alpha; a.beta = beta;
    a.p0 = probs[0]; a.pair0 = pairs[0];
    // ---- warp arrangement: fragments (8 x 8) the tile needs, rounded up to the menu {1, 2, 4, 8}; WR x WC warps with
    // exactly MF x NF fragments each, the arrangement with the fewest DMMAs per warp ----
    auto pow2 = [](int x) { int p = 1; while (p < x) p <<= 1; return p; };
    const int R = pow2(t.BM / 8), Cc = pow2(t.Np / 8);
    int best = 1 << 20, WR = 2, MF = 4, NF = 4;
    for (int wr : {2, 4, 1}) {  // ties: the square arrangement first
        const int wc = 4 / wr;
        const int mf = std::max(1, R / wr), nf = std::max(1, Cc / wc);
        if (mf * nf < best) { best = mf * nf; WR = wr; MF = mf; NF = nf; }
    }
    const int rows_p = WR * MF * 8, cols_p = (4 / WR) * NF * 8;  // footprint of the warps (>= BM, Np)
    a.a_kfast = s.nk >= 1 && s.ksL[0] == 1 && s.kext[0] > 1;
    a.b_kfast = s.nk >= 1 && s.ksR[0] == 1 && s.kext[0] > 1;
    // bytes staged per contracted element decide the chunk length: <= 14 KB per stage (4 stages, 3-4 CTAs per SM), unless
    // the whole contracted range fits 20 KB (one chunk per operand pair)
    const int perk = 8 * (rows_p + cols_p + 8);
    const int k4 = (s.K + 3) & ~3;
    int KC = (14 * 1024 / perk) & ~3;
    if ((long long)k4 * perk <= 20 * 1024) KC = k4;
    KC = std::max(4, std::min(KC, std::min(k4, 512)));
    for (;; KC -= 4) {  // exact footprint of the ring (the [x][k] layout pads every row): at most 80 KB per CTA
        a.KC = KC;
        const int ldk = KC + (KC % 8 == 0 ? 4 : 8);  // = 4 (mod 8)
        if (a.a_kfast) { a.a_sk = 1; a.a_sx = ldk; a.a_elems = rows_p * ldk; }
        else { a.a_sx = 1; a.a_sk = rows_p + 4; a.a_elems = KC * (rows_p + 4); }
        int b_elems;
        if (a.b_kfast) { a.b_sk = 1; a.b_sx = ldk; b_elems = cols_p * ldk; }
        else { a.b_sx = 1; a.b_sk = cols_p + 4; b_elems = KC * (cols_p + 4); }
        a.stage_elems = a.a_elems + b_elems;
        if ((size_t)kLStages * a.stage_elems * 8 <= 80 * 1024 || KC <= 4) break;
    }
    a.cpp = (s.K + KC - 1) / KC;
    // 16-byte items along the walking direction: unit stride there, even extent of the leading dimension (pairs never
    // straddle it), every other stride of the operand even, 16-byte aligned blocks; the staged layout keeps pairs adjacent
    auto all_even = [](int nd, const int* st, int from) { for (int i = from; i < nd; ++i) if (st[i] & 1) return false; return true; };
    if (a.a_kfast) a.a_vec = s.kext[0] % 2 == 0 && all_even(s.nk, s.ksL, 1) && all_even(s.nm, s.msL, 0) && KC % 2 == 0;
    else a.a_vec = s.nm >= 1 && s.msL[0] == 1 && s.mext[0] % 2 == 0 && all_even(s.nm, s.msL, 1) && all_even(s.nk, s.ksL, 0);
    if (a.b_kfast) a.b_vec = s.kext[0] % 2 == 0 && all_even(s.nk, s.ksR, 1) && all_even(s.nn, s.nsR, 0) && KC % 2 == 0;
    else a.b_vec = s.nn >= 1 && s.nsR[0] == 1 && s.next[0] % 2 == 0 && all_even(s.nn, s.nsR, 1) && all_even(s.nk, s.ksR, 0);
    if (a.a_vec || a.b_vec)
        for (const Pair& p : pairs) {
            if (a.a_vec && ((uintptr_t)p.L & 15)) a.a_vec = 0;
            if (a.b_vec && ((uintptr_t)p.R & 15)) a.b_vec = 0;
        }
    const size_t smem = (size_t)kLStages * a.stage_elems * 8;
    const long long nwork0 = (long long)n * t.ntile_m;
    const long long chunk_bytes = (long long)KC * 8 * (std::min(s.M, kLBM) + s.N);
    const long long total_chunks = total_pairs * a.cpp;
    const std::function<int()> pre = run_prescale;
#define LOW_CASE(mf, nf, wr) \
    if (MF == mf && NF == nf && WR == wr) return launch_low<mf, nf, wr>(a, smem, nwork0, dense_d, total_chunks, chunk_bytes, pre)
    LOW_CASE(4, 4, 2); LOW_CASE(4, 2, 2); LOW_CASE(2, 4, 2); LOW_CASE(2, 2, 2); LOW_CASE(4, 1, 2); LOW_CASE(1, 4, 2);
    LOW_CASE(2, 1, 2); LOW_CASE(1, 2, 2); LOW_CASE(1, 1, 2); LOW_CASE(2, 1, 4); LOW_CASE(1, 1, 4); LOW_CASE(1, 2, 1);
    LOW_CASE(1, 1, 1);
#undef LOW_CASE
    set_error("lowint: no kernel for the warp arrangement %d x %d fragments on %d x %d warps", MF, NF, WR, 4 / WR);
    return SIPGPU_E_ARG;
}

}  // namespace sipgpu

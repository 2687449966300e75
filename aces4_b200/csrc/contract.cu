// contract.cu -- the contraction core of the SIAL `contract` super-instruction on sm_100a.
//
// Replaces tensor_block_contract_ (tensor_dil_omp.F90:662-796): the reference permutes L and R into
// temporaries, zero-fills a third, calls dgemm('T','N') and permutes the result (4 extra passes over
// memory + 3 allocations per block).  Here ONE kernel does it: the operand permutes are folded into the
// global->shared gathers (cp.async with per-row / per-k offset tables built from the Shape's strides), the
// product runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64 = DMMA.8x8x4, the only FP64 MMA on this part;
// tcgen05 has no f64 kind), and the output permute (+ optional accumulate) is folded into the epilogue
// scatter.  A launch processes a whole work-list of blocks (block-sparse batching): persistent CTAs walk a
// prefix-summed tile list, so thousands of small heterogeneous blocks cost one launch.
#include "contract.h"

#include <stdlib.h>

#ifndef SIP_LOAD_DIV
#define SIP_LOAD_DIV 2  // loads of stage j+1 are spread over the first 1/SIP_LOAD_DIV of stage j's DMMA groups
#endif

namespace sipgpu {

namespace {

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 8-byte cp.async with zero-fill when !valid (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
template <bool VEC>
__device__ __forceinline__ void cp_async_item(double* smem_dst, const double* gsrc, bool valid) {
    if constexpr (VEC) cp_async16(smem_dst, gsrc, valid); else cp_async8(smem_dst, gsrc, valid);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// linear index -> element offset in up to two tensors, over the (collapsed) dims of one index group
__device__ __forceinline__ void decompose2(int lin, int nd, const int* ext, const int* s0, const int* s1, int& o0,
                                           int& o1) {
    o0 = 0;
    o1 = 0;
#pragma unroll 1
    for (int i = 0; i < nd; ++i) {
        const int e = ext[i];
        const int q = lin / e, r = lin - q * e;
        o0 += r * s0[i];
        o1 += r * s1[i];
        lin = q;
    }
}

template <int WARPS_M_, int WARPS_N_, int MF_, int NF_, int BK_, int STAGES_, bool A_KC_, bool B_KC_, bool VEC_>
struct Cfg {
    static constexpr int WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, MF = MF_, NF = NF_, STAGES = STAGES_;
    static constexpr bool A_KC = A_KC_, B_KC = B_KC_, VEC = VEC_;
    static constexpr int BM = WARPS_M * MF * 8, BN = WARPS_N * NF * 8, BK = BK_;
    static constexpr int NT = 32 * WARPS_M * WARPS_N;
    static constexpr int KWIN = 2048;    // k offsets are tabulated for a window of this many contracted elements
    static constexpr int LDK = BK + 4;   // K-contiguous tile: [row][k]; stride = 4 mod 16 doubles -> conflict-free frags
    static constexpr int LDAM = BM + 4;  // M-contiguous tile: [k][m]
    static constexpr int LDBN = BN + 4;
    static constexpr int A_ELEMS = A_KC ? BM * LDK : BK * LDAM;
    static constexpr int B_ELEMS = B_KC ? BN * LDK : BK * LDBN;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    // global->shared mapping: each cp.async item moves V consecutive doubles (16 bytes when VEC) along the
    // operand's contiguous direction; a thread keeps that lane fixed and walks the other direction in passes
    static constexpr int V = VEC ? 2 : 1;
    // LANES threads span the contiguous direction, PER = NT / LANES rows (or k's) are covered per pass; threads
    // beyond PER*LANES idle (only when BN is not a divisor of NT, e.g. the 128x80 tile)
    static constexpr int A_LANES = (A_KC ? BK : BM) / V, B_LANES = (B_KC ? BK : BN) / V;
    static constexpr int A_PER = NT / A_LANES, B_PER = NT / B_LANES;
    static constexpr int A_ITEMS = ((A_KC ? BM : BK) + A_PER - 1) / A_PER;
    static constexpr int B_ITEMS = ((B_KC ? BN : BK) + B_PER - 1) / B_PER;
    static constexpr int ITEMS = A_ITEMS + B_ITEMS;
    static constexpr int MMA_GROUPS = (BK / 4) * MF;  // groups of NF DMMAs per stage
    // the next stage's cp.async items are issued during the first LOAD_GROUPS groups only: the tail of the stage
    // is latency slack, so the barrier that opens the next stage does not wait for loads still in flight
    static constexpr int LOAD_GROUPS = MMA_GROUPS / SIP_LOAD_DIV > 0 ? MMA_GROUPS / SIP_LOAD_DIV : 1;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * 8 + (size_t)(2 * BM + 2 * BN) * 4 + (size_t)4 * KWIN * 4;
    static_assert(A_PER >= 1 && B_PER >= 1 && LOAD_GROUPS + 1 <= MMA_GROUPS, "load mapping");
    static_assert(KWIN % BK == 0 && LDK % 16 == 4 && LDAM % 16 == 4 && LDBN % 16 == 4, "window / bank layout");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <class C>
__global__ void __launch_bounds__(C::NT, 1)
contract_kernel(const __grid_constant__ ContractArgs args) {
    constexpr int BM = C::BM, BN = C::BN, BK = C::BK, NT = C::NT, STAGES = C::STAGES, MF = C::MF, NF = C::NF;
    constexpr int KWIN = C::KWIN;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    int* mOffL = reinterpret_cast<int*>(tiles + (size_t)STAGES * C::STAGE_ELEMS);
    int* mOffD = mOffL + BM;
    int* nOffR = mOffD + BM;
    int* nOffD = nOffR + BN;
    int* kOffL = nOffD + BN;  // [2][KWIN]: double-buffered by segment parity when K spans several windows
    int* kOffR = kOffL + 2 * KWIN;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int wm = (warp % C::WARPS_M) * (MF * 8), wn = (warp / C::WARPS_M) * (NF * 8);
    const double alpha = args.alpha, beta = args.beta;

    // per-thread constants of the global->shared mapping
    const int a_fix = (tid % C::A_LANES) * C::V;  // the k lane (K-contiguous) or the m lane (M-contiguous)
    const int a_var = tid / C::A_LANES;           // first row / first k of the passes
    const int b_fix = (tid % C::B_LANES) * C::V;
    const int b_var = tid / C::B_LANES;

    for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
        // ---- which block of the work-list does this tile belong to? (uniform binary search) ----
        int p = 0;
        if (args.nprob > 1) {
            int lo = 0, hi = args.nprob;  // tile_prefix[lo] <= tile < tile_prefix[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(args.tile_prefix + mid) <= tile) lo = mid; else hi = mid;
            }
            p = lo;
        }
        const Problem* pr = args.probs ? args.probs + p : &args.p0;
        const Shape* sh = args.probs ? args.shapes + pr->shape : &args.s0;
        const Pair* chain = args.probs ? args.pairs + pr->chain_begin : &args.pair0;
        const int chain_len = pr->chain_len;
        const int local = tile - (args.nprob > 1 ? __ldg(args.tile_prefix + p) : 0);
        const int M = sh->M, N = sh->N, K = sh->K;
        const int tiles_m = (M + BM - 1) / BM;
        const int m0 = (local % tiles_m) * BM, n0 = (local / tiles_m) * BN;
        double* __restrict__ Dp = pr->D;
        const int nk = sh->nk;
        const int nwin = (K + KWIN - 1) / KWIN;

        __syncthreads();  // previous tile fully consumed (tables + stages)
        // ---- offset tables: the input/output permutes of F90:731-743,782-785 as address arithmetic ----
        for (int i = tid; i < BM + BN; i += NT) {
            if (i < BM) {
                const int m = m0 + i;
                int oL = -1, oD = -1;
                if (m < M) decompose2(m, sh->nm, sh->mext, sh->msL, sh->msD, oL, oD);
                mOffL[i] = oL;
                mOffD[i] = oD;
            } else {
                const int j = i - BM, n = n0 + j;
                int oR = -1, oD = -1;
                if (n < N) decompose2(n, sh->nn, sh->next, sh->nsR, sh->nsD, oR, oD);
                nOffR[j] = oR;
                nOffD[j] = oD;
            }
        }

        double acc[MF][NF][2];
#pragma unroll
        for (int i = 0; i < MF; ++i)
#pragma unroll
            for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // ---- the K loop runs over SEGMENTS = (operand pair of the chain) x (k-offset window) as ONE software
        // pipeline: stages keep flowing across pair boundaries, so a chain costs no drain/refill per block ----
        const int nseg = chain_len * nwin;
        const int last_steps = (K - (nwin - 1) * KWIN + BK - 1) / BK;
        auto seg_steps = [&](int seg) { return (nwin == 1 || (seg % nwin) == nwin - 1) ? last_steps : KWIN / BK; };
        const int total_steps = chain_len * ((nwin - 1) * (KWIN / BK) + last_steps);
        // k offsets of one window (contracted-index permute as address arithmetic); -1 pads the ragged tail
        auto fill_ktable = [&](int seg, int buf) {
            const int kbase = (seg % nwin) * KWIN;
            const int kcount = min(KWIN, K - kbase);
            const int nfill = seg_steps(seg) * BK;
            for (int i = tid; i < nfill; i += NT) {
                int oL = -1, oR = -1;
                if (i < kcount) decompose2(kbase + i, nk, sh->kext, sh->ksL, sh->ksR, oL, oR);
                kOffL[buf * KWIN + i] = oL;
                kOffR[buf * KWIN + i] = oR;
            }
        };
        fill_ktable(0, 0);
        if (nwin > 1) fill_ktable(1, 1);  // nwin > 1 implies nseg > 1
        __syncthreads();

        const int mo_fix = (C::A_KC || a_var >= C::A_PER) ? 0 : mOffL[a_fix];
        const int no_fix = (C::B_KC || b_var >= C::B_PER) ? 0 : nOffR[b_fix];

        // load cursor: the (segment, step) whose tiles are fetched next
        int l_seg = 0, l_ks = 0, l_steps = seg_steps(0);
        const double* __restrict__ Lp = chain[0].L;
        const double* __restrict__ Rp = chain[0].R;
        const int* kl_tab = kOffL;
        const int* kr_tab = kOffR;
        auto advance_load_cursor = [&]() {
            if (++l_ks == l_steps) {
                l_ks = 0;
                ++l_seg;
                if (l_seg < nseg) {
                    l_steps = seg_steps(l_seg);
                    const Pair pq = chain[l_seg / nwin];
                    Lp = pq.L;
                    Rp = pq.R;
                    if (nwin > 1) {
                        kl_tab = kOffL + (l_seg & 1) * KWIN;
                        kr_tab = kOffR + (l_seg & 1) * KWIN;
                    }
                }
            }
        };
        // One cp.async item (8 bytes, or 16 when VEC) of the load cursor's step = item `it` of C::ITEMS; out-of-range
        // elements are zero-filled.  With VEC the host guarantees even extents/strides along the contiguous direction,
        // so both doubles of an item are valid or invalid together and 16-byte aligned.  The element offset of an
        // item is (stage-constant part) + (per-item part), both read from the offset tables: item_guard() says
        // whether this thread has the item, item_lookup() reads its per-item table entry, item_issue() consumes it.
        // In the main loop the lookup of item it+1 is issued before the cp.async of item it, so the dependent
        // address arithmetic never waits on shared memory in front of the DMMAs of an in-order warp.
        auto item_guard = [&](int it) -> bool {
            if (it < C::A_ITEMS) {
                const int x = a_var + it * C::A_PER;
                return (((C::A_KC ? BM : BK) % C::A_PER == 0) && NT % C::A_LANES == 0) || (x < (C::A_KC ? BM : BK) && a_var < C::A_PER);
            }
            const int x = b_var + (it - C::A_ITEMS) * C::B_PER;
            return (((C::B_KC ? BN : BK) % C::B_PER == 0) && NT % C::B_LANES == 0) || (x < (C::B_KC ? BN : BK) && b_var < C::B_PER);
        };
        auto item_lookup = [&](int it) -> int {
            if (!item_guard(it)) return -1;
            if (args.dbg & 2) return 0;
            const int kb = l_ks * BK;
            if (it < C::A_ITEMS) {
                const int x = a_var + it * C::A_PER;
                return C::A_KC ? mOffL[x] : kl_tab[kb + x];
            }
            const int x = b_var + (it - C::A_ITEMS) * C::B_PER;
            return C::B_KC ? nOffR[x] : kr_tab[kb + x];
        };
        // stage-constant part: the k offset of this thread's lane (K-contiguous) or its m/n offset (M/N-contiguous)
        auto stage_fix_a = [&]() -> int { return C::A_KC ? (a_var < C::A_PER ? kl_tab[l_ks * BK + a_fix] : -1) : mo_fix; };
        auto stage_fix_b = [&]() -> int { return C::B_KC ? (b_var < C::B_PER ? kr_tab[l_ks * BK + b_fix] : -1) : no_fix; };
        auto item_issue = [&](int it, int per_item, int fix_a, int fix_b, double* as, double* bs) {
            if (!item_guard(it)) return;
            if (args.dbg & 1) return;
            if (args.dbg & 2) { per_item = (it * 256 + tid) * 2; fix_a = fix_b = 0; }
            if (it < C::A_ITEMS) {
                const int x = a_var + it * C::A_PER;
                const bool v = (per_item | fix_a) >= 0;
                double* dst = C::A_KC ? as + x * C::LDK + a_fix : as + x * C::LDAM + a_fix;
                cp_async_item<C::VEC>(dst, Lp + (v ? per_item + fix_a : 0), v);
            } else {
                const int x = b_var + (it - C::A_ITEMS) * C::B_PER;
                const bool v = (per_item | fix_b) >= 0;
                double* dst = C::B_KC ? bs + x * C::LDK + b_fix : bs + x * C::LDBN + b_fix;
                cp_async_item<C::VEC>(dst, Rp + (v ? per_item + fix_b : 0), v);
            }
        };

#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            if (s < total_steps) {
                double* as = tiles + (size_t)s * C::STAGE_ELEMS;
                {
                    const int fa = stage_fix_a(), fb = stage_fix_b();
#pragma unroll
                    for (int it = 0; it < C::ITEMS; ++it) item_issue(it, item_lookup(it), fa, fb, as, as + C::A_ELEMS);
                }
                advance_load_cursor();
            }
            cp_async_commit();
        }

        // compute cursor: which steps are the ragged LAST step of an operand pair (only ceil(rem/4) of its BK/4
        // DMMA k-steps hold contracted elements; the rest is zero padding and is skipped)
        const int last_kk = (K - (nwin - 1) * KWIN - (last_steps - 1) * BK + 3) / 4;
        int c_win = 0, c_ks = 0, c_steps = seg_steps(0);

        for (int j = 0; j < total_steps; ++j) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            const int jn = j + STAGES - 1;
            const bool do_load = jn < total_steps;
            // entering a new segment: the table of the segment after it can be built now -- the last loads that read
            // that buffer (segment l_seg-1) were issued before the barrier above
            if (nwin > 1 && do_load && l_ks == 0 && l_seg + 1 < nseg && j > 0) fill_ktable(l_seg + 1, (l_seg + 1) & 1);
            double* ls = tiles + (size_t)(jn % STAGES) * C::STAGE_ELEMS;
            const double* as = tiles + (size_t)(j % STAGES) * C::STAGE_ELEMS;
            const double* bs = as + C::A_ELEMS;
            const int kk_lim = (c_ks == c_steps - 1 && c_win == nwin - 1) ? last_kk : BK / 4;
            double a[2][MF], b[2][NF];
            auto load_frags = [&](int kk, int buf) {
#pragma unroll
                for (int mi = 0; mi < MF; ++mi)
                    a[buf][mi] = C::A_KC ? as[(wm + mi * 8 + g) * C::LDK + kk * 4 + t4]
                                         : as[(kk * 4 + t4) * C::LDAM + wm + mi * 8 + g];
#pragma unroll
                for (int ni = 0; ni < NF; ++ni)
                    b[buf][ni] = C::B_KC ? bs[(wn + ni * 8 + g) * C::LDK + kk * 4 + t4]
                                         : bs[(kk * 4 + t4) * C::LDBN + wn + ni * 8 + g];
            };
            if (kk_lim == BK / 4) {
                // ---- full step: BK/4 x MF groups of NF DMMAs, fully unrolled ----
                int fix_a = -1, fix_b = -1, pre = -1;
                if (do_load) {
                    fix_a = stage_fix_a();
                    fix_b = stage_fix_b();
                    pre = item_lookup(0);
                }
                load_frags(0, 0);
#pragma unroll
                for (int kk = 0; kk < BK / 4; ++kk) {
                    if (kk + 1 < BK / 4) load_frags(kk + 1, (kk + 1) & 1);
#pragma unroll
                    for (int mi = 0; mi < MF; ++mi) {
#pragma unroll
                        for (int ni = 0; ni < NF; ++ni)
                            dmma8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[kk & 1][mi], b[kk & 1][ni]);
                        // the next stage's loads ride in the shadow of the DMMA pipe: groups 1..LOAD_GROUPS carry
                        // the cp.async items (the table entry of each item was fetched one group earlier)
                        const int grp = kk * MF + mi;
                        if (do_load && grp >= 1 && grp <= C::LOAD_GROUPS) {
#pragma unroll
                            for (int it = (grp - 1) * C::ITEMS / C::LOAD_GROUPS; it < grp * C::ITEMS / C::LOAD_GROUPS; ++it) {
                                const int cur = pre;
                                if (it + 1 < C::ITEMS) pre = item_lookup(it + 1);
                                item_issue(it, cur, fix_a, fix_b, ls, ls + C::A_ELEMS);
                            }
                        }
                    }
                }
            } else {
                // ---- ragged last step of an operand pair: loads first, then only the k-steps that hold data ----
                if (do_load) {
                    const int fa = stage_fix_a(), fb = stage_fix_b();
#pragma unroll
                    for (int it = 0; it < C::ITEMS; ++it) item_issue(it, item_lookup(it), fa, fb, ls, ls + C::A_ELEMS);
                }
#pragma unroll 1
                for (int kk = 0; kk < kk_lim; ++kk) {
                    load_frags(kk, 0);
#pragma unroll
                    for (int mi = 0; mi < MF; ++mi)
#pragma unroll
                        for (int ni = 0; ni < NF; ++ni)
                            dmma8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[0][mi], b[0][ni]);
                }
            }
            cp_async_commit();
            if (do_load) advance_load_cursor();
            if (++c_ks == c_steps) {
                c_ks = 0;
                if (++c_win == nwin) c_win = 0;
                c_steps = (c_win == nwin - 1) ? last_steps : KWIN / BK;
            }
        }
        cp_async_wait<0>();

        // ---- epilogue: D[perm(m,n)] = alpha*acc (+ beta*D): the output permute of F90:782-785 as a scatter ----
        if (!(args.dbg & 4) || acc[0][0][0] == 123.456)
#pragma unroll
        for (int mi = 0; mi < MF; ++mi) {
            const int mo = mOffD[wm + mi * 8 + g];
#pragma unroll
            for (int ni = 0; ni < NF; ++ni) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int no = nOffD[wn + ni * 8 + 2 * t4 + c];
                    if ((mo | no) >= 0) {
                        double* dst = Dp + (size_t)(mo + no);
                        double v = alpha * acc[mi][ni][c];
                        if (beta != 0.0) v += beta * *dst;
                        *dst = v;
                    }
                }
            }
        }
    }
}

template <class C>
int launch_cfg(const ContractArgs& a, int max_ctas) {
    static bool attr_set = false;
    if (!attr_set) {
        SIP_CUDA(cudaFuncSetAttribute(contract_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        attr_set = true;
    }
    int grid = a.total_tiles < max_ctas ? a.total_tiles : max_ctas;
    if (grid < 1) return SIPGPU_OK;
    static const int dbg = getenv("SIPGPU_DBG") ? atoi(getenv("SIPGPU_DBG")) : 0;
    if (dbg) {
        ContractArgs b = a;
        b.dbg = dbg;
        contract_kernel<C><<<grid, C::NT, C::SMEM, ctx().stream>>>(b);
    } else
    contract_kernel<C><<<grid, C::NT, C::SMEM, ctx().stream>>>(a);
    SIP_CUDA(cudaGetLastError());
    count_launch();
    return SIPGPU_OK;
}

}  // namespace

// Tile menu: 0 = 128x128 (8 warps of 64x32), 1 = 128x80 (8 warps of 32x40: N = 400 = o*o segments tile exactly)
void contract_tile_dims(int tile, int* bm, int* bn) {
    *bm = 128;
    *bn = tile == 1 ? 80 : 128;
}
int contract_pick_tile(int M, int N) {
    long long best = -1;
    int pick = 0;
    for (int t = 0; t < 2; ++t) {
        int bm, bn;
        contract_tile_dims(t, &bm, &bn);
        const long long padded = (long long)((M + bm - 1) / bm) * bm * (long long)((N + bn - 1) / bn) * bn;
        if (best < 0 || padded < best) { best = padded; pick = t; }  // ties keep the larger tile (listed first)
    }
    return pick;
}

template <int WM, int WN, int MF, int NF>
int launch_tile(const ContractArgs& a, bool a_kc, bool b_kc, bool vec, int ctas) {
    if (vec) {
        if (a_kc && b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, 32, 2, true, true, true>>(a, ctas);
        if (a_kc && !b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, 32, 2, true, false, true>>(a, ctas);
        if (!a_kc && b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, 32, 2, false, true, true>>(a, ctas);
        return launch_cfg<Cfg<WM, WN, MF, NF, 32, 2, false, false, true>>(a, ctas);
    }
    if (a_kc && b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, 32, 2, true, true, false>>(a, ctas);
    if (a_kc && !b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, 32, 2, true, false, false>>(a, ctas);
    if (!a_kc && b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, 32, 2, false, true, false>>(a, ctas);
    return launch_cfg<Cfg<WM, WN, MF, NF, 32, 2, false, false, false>>(a, ctas);
}

int launch_contract(const ContractArgs& a, bool a_kc, bool b_kc, bool vec, int tile) {
    const int ctas = ctx().num_sms;
    // BK = 32 double-buffered (one barrier per 32 contracted elements)
    static const int warps16 = getenv("SIPGPU_WARPS16") ? atoi(getenv("SIPGPU_WARPS16")) : 0;  // tuning experiment
    if (warps16) {
        if (tile == 1) return launch_tile<8, 2, 2, 5>(a, a_kc, b_kc, vec, ctas);
        return launch_tile<4, 4, 4, 4>(a, a_kc, b_kc, vec, ctas);
    }
    if (tile == 1) return launch_tile<4, 2, 4, 5>(a, a_kc, b_kc, vec, ctas);
    return launch_tile<2, 4, 8, 4>(a, a_kc, b_kc, vec, ctas);
}

// ---------------- register-resident DMMA issue-rate probe (roofline denominator for the FP64 tensor pipe) ------
__global__ void __launch_bounds__(256) dmma_probe_kernel(int iters, double* sink) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma8x8x4(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

int dmma_probe(int iters, double* tflops) {
    Ctx& c = ctx();
    cudaEvent_t e0, e1;
    SIP_CUDA(cudaEventCreate(&e0));
    SIP_CUDA(cudaEventCreate(&e1));
    const int ctas = c.num_sms * 2;
    dmma_probe_kernel<<<ctas, 256, 0, c.stream>>>(iters / 8 + 1, c.d_reduce);  // warm-up
    SIP_CUDA(cudaEventRecord(e0, c.stream));
    dmma_probe_kernel<<<ctas, 256, 0, c.stream>>>(iters, c.d_reduce);
    SIP_CUDA(cudaEventRecord(e1, c.stream));
    SIP_CUDA(cudaEventSynchronize(e1));
    count_launch(2);
    float ms = 0;
    SIP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (double)ctas * 8 /*warps*/ * (double)iters * 16 * (2.0 * 8 * 8 * 4);
    *tflops = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SIPGPU_OK;
}

}  // namespace sipgpu

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
//
// Structure (round-1 ablation, scripts/micro/dmma_ablate.cu: a loop that only does LDS.64 + DMMA reaches 99 % of
// the DMMA issue-rate peak, the same loop with the gather addressing and cp.async issue woven into the DMMA
// warps stalls at 90 %): the CTA is WARP-SPECIALISED.
//   * warpgroup 0 (4 warps, 56 registers after setmaxnreg.dec) is the PRODUCER: it resolves the work-list,
//     builds the offset tables, and streams operand tiles into a 4-stage shared-memory ring with cp.async
//     (zero-filled at ragged edges); each stage is published through an mbarrier that the cp.async engine
//     itself arrives on (cp.async.mbarrier.arrive.noinc);
//   * warpgroups 1-2 (8 warps, 224 registers after setmaxnreg.inc) are the CONSUMERS: wait(full) -> LDS.64
//     fragments -> DMMA -> arrive(empty), nothing else in the loop, then the epilogue scatter.  There is no
//     CTA-wide barrier in steady state, and the producer refills the ring for the next tile while the
//     consumers are still storing the previous one.
#include "contract.h"

#include <stdlib.h>

#include <type_traits>
#include <utility>

namespace sipgpu {

namespace {

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// fragment load with a compile-time byte offset; volatile so that its place between the DMMAs is kept
template <int OFF>
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];\n" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

// 8-byte cp.async with zero-fill when !valid (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
template <bool VEC>
__device__ __forceinline__ void cp_async_item(double* smem_dst, const double* gsrc, bool valid) {
    if constexpr (VEC) cp_async16(smem_dst, gsrc, valid); else cp_async8(smem_dst, gsrc, valid);
}

// ---- mbarrier helpers (shared::cta, 64-bit objects) ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// arrive-on triggered when all cp.async issued so far by this thread have landed in shared memory
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "WAIT_%=:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_%=;\n"
        " bra WAIT_%=;\n"
        "DONE_%=:\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void producer_bar_sync() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }

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

constexpr int kProducerThreads = 128;
constexpr int kProducerRegs = 56, kConsumerRegs = 224;  // 128*56 + 256*224 = 384*168 registers

// what the consumers need to know about a tile (written by the producer, double-buffered by tile parity)
struct TileInfo {
    double* D;
    int total_steps;     // ring stages this tile consumes (0 = no more tiles)
    int steps_per_pair;  // stages per operand pair of the chain
    int last_kk;         // DMMA k-steps (of 4) that hold data in the LAST stage of a pair
    int pad;
};

template <int WARPS_M_, int WARPS_N_, int MF_, int NF_, bool A_KC_, bool B_KC_, bool VEC_>
struct Cfg {
    static constexpr int WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, MF = MF_, NF = NF_;
    static constexpr bool A_KC = A_KC_, B_KC = B_KC_, VEC = VEC_;
    static constexpr int BM = WARPS_M * MF * 8, BN = WARPS_N * NF * 8, BK = 16;
    // 4-stage ring; the narrow two-CTAs-per-SM tiles (128x56, 128x24) take 3 so that two CTAs fit the 227 KB of an SM
    static constexpr int STAGES = (WARPS_M * WARPS_N == 4 && BN < 64) ? 3 : 4;
    static constexpr int NC = 32 * WARPS_M * WARPS_N;  // consumer threads
    static constexpr int NP = kProducerThreads;
    static constexpr int NT = NP + NC;
    static constexpr int KWIN = 2048;    // k offsets are tabulated for a window of this many contracted elements
    static constexpr int LDK = BK + 4;   // K-contiguous tile: [row][k]; stride = 4 mod 16 doubles -> conflict-free frags
    static constexpr int LDAM = BM + (20 - BM % 16) % 16;  // M-contiguous tile: [k][m]; stride = 4 mod 16 as well
    static constexpr int LDBN = BN + (20 - BN % 16) % 16;
    static constexpr int A_ELEMS = A_KC ? BM * LDK : BK * LDAM;
    static constexpr int B_ELEMS = B_KC ? BN * LDK : BK * LDBN;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    // global->shared mapping of the producer: each cp.async item moves V consecutive doubles (16 bytes when VEC)
    // along the operand's contiguous direction; a thread keeps that lane fixed and walks the other direction
    static constexpr int V = VEC ? 2 : 1;
    static constexpr int A_LANES = (A_KC ? BK : BM) / V, B_LANES = (B_KC ? BK : BN) / V;
    static constexpr int A_PER = NP / A_LANES, B_PER = NP / B_LANES;
    static constexpr int A_ITEMS = ((A_KC ? BM : BK) + A_PER - 1) / A_PER;
    static constexpr int B_ITEMS = ((B_KC ? BN : BK) + B_PER - 1) / B_PER;
    // shared memory: [stages][tables x2 parities][k tables][tile info x2][mbarriers]
    static constexpr size_t OFF_TABLES = (size_t)STAGES * STAGE_ELEMS * 8;
    static constexpr size_t OFF_KTAB = OFF_TABLES + (size_t)2 * (2 * BM + 2 * BN) * 4;
    static constexpr size_t OFF_INFO = OFF_KTAB + (size_t)2 * KWIN * 4;
    static constexpr size_t OFF_BARS = OFF_INFO + 2 * sizeof(TileInfo);
    static constexpr size_t SMEM = OFF_BARS + (size_t)(2 * STAGES + 4) * 8;
    // 8 consumer warps: one CTA per SM, registers rebalanced with setmaxnreg (56 / 224).  4 consumer warps (the 64x64
    // tile for launches with fewer tiles than SMs): two CTAs per SM, 56 / 200.
    static constexpr int MIN_CTAS = NC == 256 ? 1 : 2;
    static constexpr int CONSUMER_REGS = NC == 256 ? kConsumerRegs : 200;  // 128*(56+200) = 256*128: half an SM
    static_assert((NC == 256 || NC == 128) && A_PER >= 1 && B_PER >= 1, "thread mapping");
    static_assert(KWIN % BK == 0 && LDK % 16 == 4 && LDAM % 16 == 4 && LDBN % 16 == 4, "window / bank layout");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <class C>
__global__ void __launch_bounds__(C::NT, C::MIN_CTAS)
contract_kernel(const __grid_constant__ ContractArgs args) {
    constexpr int BM = C::BM, BN = C::BN, BK = C::BK, STAGES = C::STAGES, MF = C::MF, NF = C::NF, KWIN = C::KWIN;
    constexpr int NP = C::NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    int* tabs = reinterpret_cast<int*>(smem_raw + C::OFF_TABLES);   // per parity: mOffL[BM] mOffD[BM] nOffR[BN] nOffD[BN]
    int* kOffL = reinterpret_cast<int*>(smem_raw + C::OFF_KTAB);    // [KWIN]
    int* kOffR = kOffL + KWIN;
    TileInfo* info = reinterpret_cast<TileInfo*>(smem_raw + C::OFF_INFO);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + C::OFF_BARS);
    unsigned long long* full = bars;                 // [STAGES]  producer -> consumers (cp.async completion)
    unsigned long long* empty = bars + STAGES;       // [STAGES]  consumers -> producer
    unsigned long long* tab_full = bars + 2 * STAGES;  // [2]     tables + TileInfo of a tile are written
    unsigned long long* tab_empty = tab_full + 2;      // [2]     consumers are done with them (epilogue finished)

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, NP);
            mbar_init(empty + s, C::NC / 32);
        }
        for (int p = 0; p < 2; ++p) {
            mbar_init(tab_full + p, NP);
            mbar_init(tab_empty + p, C::NC / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (tid < NP) {
        // =====================================  PRODUCER  =====================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kProducerRegs));
        const int a_fix = (tid % C::A_LANES) * C::V;  // the k lane (K-contiguous) or the m lane (M-contiguous)
        const int a_var = tid / C::A_LANES;           // first row / first k of the passes
        const int b_fix = (tid % C::B_LANES) * C::V;
        const int b_var = tid / C::B_LANES;
        int ring = 0;  // stages produced so far (all tiles)
        int tcount = 0;
        bool lockstep = true;  // thread 0: still honouring the lockstep throttle
        for (int tile = blockIdx.x;; tile += gridDim.x, ++tcount) {
            const int par = tcount & 1;
            int* mOffL = tabs + par * (2 * BM + 2 * BN);
            int* mOffD = mOffL + BM;
            int* nOffR = mOffD + BM;
            int* nOffD = nOffR + BN;
            // the consumers must have finished the epilogue of the tile that used this parity (2 tiles ago)
            if (tcount >= 2) mbar_wait(tab_empty + par, ((tcount >> 1) - 1) & 1);
            if (tile >= args.total_tiles) {
                if (tid == 0) info[par].total_steps = 0;
                mbar_arrive(tab_full + par);
                break;
            }
            // ---- which block of the work-list does this tile belong to?  32-way search: every lane probes one split
            // point of the prefix-sum array, so a work-list of 32K blocks costs 3 dependent loads instead of 15 ----
            int p = 0;
            if (args.nprob > 1) {
                const int lane = tid & 31;
                int lo = 0, hi = args.nprob;  // tile_prefix[lo] <= tile < tile_prefix[hi]
                while (hi - lo > 1) {
                    const int step = (hi - lo + 31) / 32;
                    const int idx = lo + (lane + 1) * step;
                    const bool le = idx < hi && __ldg(args.tile_prefix + idx) <= tile;
                    const int cnt = __popc(__ballot_sync(0xffffffffu, le));  // monotone: the lanes with `le` are a prefix
                    lo += cnt * step;
                    hi = min(hi, lo + step);
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
            const int nk = sh->nk;
            const int nwin = (K + KWIN - 1) / KWIN;
            const int last_steps = (K - (nwin - 1) * KWIN + BK - 1) / BK;
            // split-K: this problem covers the k windows [w0, w1) only (Problem::kwin, 0 = all of them)
            const int w0 = pr->kwin ? (pr->kwin & 0xffff) : 0, w1 = pr->kwin ? (pr->kwin >> 16) : nwin;
            const bool has_tail = w1 == nwin;
            const int steps_per_pair = (w1 - w0 - (has_tail ? 1 : 0)) * (KWIN / BK) + (has_tail ? last_steps : 0);

            // ---- offset tables: the input/output permutes of F90:731-743,782-785 as address arithmetic ----
            for (int i = tid; i < BM + BN; i += NP) {
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
            if (tid == 0) {
                TileInfo ti;
                ti.D = pr->D;
                ti.total_steps = chain_len * steps_per_pair;
                ti.steps_per_pair = steps_per_pair;
                ti.last_kk = has_tail ? (K - (nwin - 1) * KWIN - (last_steps - 1) * BK + 3) / 4 : BK / 4;
                ti.pad = 0;
                info[par] = ti;
            }
            mbar_arrive(tab_full + par);  // release: tables + info are visible to whoever waits on this phase

            // k offsets of one window (contracted-index permute as address arithmetic); -1 pads the ragged tail.
            // Single buffer: only producer threads read it, between two producer-only barriers.
            auto fill_ktable = [&](int win) {
                producer_bar_sync();  // every producer thread is done reading the previous window
                const int kbase = win * KWIN;
                const int kcount = min(KWIN, K - kbase);
                const int nfill = (win == nwin - 1 ? last_steps : KWIN / BK) * BK;
                for (int i = tid; i < nfill; i += NP) {
                    int oL = -1, oR = -1;
                    if (i < kcount) decompose2(kbase + i, nk, sh->kext, sh->ksL, sh->ksR, oL, oR);
                    kOffL[i] = oL;
                    kOffR[i] = oR;
                }
                producer_bar_sync();  // (also orders the m/n tables written above before their first use below)
            };
            fill_ktable(w0);
            const int mo_fix = (C::A_KC || a_var >= C::A_PER) ? 0 : mOffL[a_fix];
            const int no_fix = (C::B_KC || b_var >= C::B_PER) ? 0 : nOffR[b_fix];

            int jstage = 0;  // stages of THIS tile produced so far
            for (int c = 0; c < chain_len; ++c) {
                const Pair pq = chain[c];
                const double* __restrict__ Lp = pq.L;
                const double* __restrict__ Rp = pq.R;
                for (int win = w0; win < w1; ++win) {
                    if (w1 - w0 > 1) fill_ktable(win);
                    const int steps = win == nwin - 1 ? last_steps : KWIN / BK;
                    for (int ks = 0; ks < steps; ++ks, ++ring, ++jstage) {
                        if (args.sync_q && jstage % kSyncEvery == 0) {
                            if (tid == 0) {
                                const int q = jstage / kSyncEvery;
                                int* ctr = args.sync_ctr + (size_t)tcount * args.sync_q;
                                atomicAdd(ctr + q, 1);
                                if (q >= kSyncWindow && lockstep) {
                                    const int target = min((int)gridDim.x, args.total_tiles - tcount * (int)gridDim.x);
                                    int budget = 1 << 16;   // never block progress: a stuck peer only costs the throttle
                                    while (*(volatile int*)(ctr + q - kSyncWindow) < target && --budget) __nanosleep(100);
                                    if (!budget) lockstep = false;
                                }
                            }
                            producer_bar_sync();
                        }
                        const int stage = ring % STAGES;
                        mbar_wait(empty + stage, ((ring / STAGES) & 1) ^ 1);  // slot free (first lap passes at once)
                        double* as = tiles + (size_t)stage * C::STAGE_ELEMS;
                        double* bs = as + C::A_ELEMS;
                        const int kb = ks * BK;
                        // One cp.async item = V consecutive doubles of the operand's contiguous direction; elements
                        // outside the block (ragged tile edges, K tail) are zero-filled.  With VEC the planner
                        // guarantees even extents/strides along that direction, so both doubles of an item are valid
                        // or invalid together and 16-byte aligned.
                        if (a_var < C::A_PER) {
                            const int fix = C::A_KC ? kOffL[kb + a_fix] : mo_fix;
#pragma unroll
                            for (int it = 0; it < C::A_ITEMS; ++it) {
                                const int x = a_var + it * C::A_PER;  // row (K-contiguous) or k (M-contiguous)
                                if ((C::A_KC ? BM : BK) % C::A_PER == 0 || x < (C::A_KC ? BM : BK)) {
                                    const int per = C::A_KC ? mOffL[x] : kOffL[kb + x];
                                    const bool v = (per | fix) >= 0;
                                    double* dst = C::A_KC ? as + x * C::LDK + a_fix : as + x * C::LDAM + a_fix;
                                    cp_async_item<C::VEC>(dst, Lp + (v ? per + fix : 0), v);
                                }
                            }
                        }
                        if (b_var < C::B_PER) {
                            const int fix = C::B_KC ? kOffR[kb + b_fix] : no_fix;
#pragma unroll
                            for (int it = 0; it < C::B_ITEMS; ++it) {
                                const int x = b_var + it * C::B_PER;
                                if ((C::B_KC ? BN : BK) % C::B_PER == 0 || x < (C::B_KC ? BN : BK)) {
                                    const int per = C::B_KC ? nOffR[x] : kOffR[kb + x];
                                    const bool v = (per | fix) >= 0;
                                    double* dst = C::B_KC ? bs + x * C::LDK + b_fix : bs + x * C::LDBN + b_fix;
                                    cp_async_item<C::VEC>(dst, Rp + (v ? per + fix : 0), v);
                                }
                            }
                        }
                        mbar_arrive_cp_async(full + stage);
                    }
                }
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else {
        // =====================================  CONSUMERS  =====================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(C::CONSUMER_REGS));
        const int ctid = tid - NP, lane = ctid & 31, warp = ctid >> 5;
        const int g = lane >> 2, t4 = lane & 3;
        const int wm = (warp % C::WARPS_M) * (MF * 8), wn = (warp / C::WARPS_M) * (NF * 8);
        const double alpha = args.alpha, beta = args.beta;
        int ring = 0;  // stages consumed so far (all tiles)
        for (int tcount = 0;; ++tcount) {
            const int par = tcount & 1;
            mbar_wait(tab_full + par, (tcount >> 1) & 1);
            const int total_steps = info[par].total_steps;
            if (total_steps == 0) break;
            const int steps_per_pair = info[par].steps_per_pair;
            const int last_kk = info[par].last_kk;

            double acc[MF][NF][2];
#pragma unroll
            for (int i = 0; i < MF; ++i)
#pragma unroll
                for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

            // Fragment addressing: per thread one base address for A and one for B inside a stage; the (mi, kk) /
            // (ni, kk) part is a compile-time immediate.  The LDS.64 of k-step kk+1 are hand-interleaved with the
            // DMMAs of k-step kk (one load after every second DMMA, volatile asm keeps the order), so a fragment is
            // always loaded a full k-step before the DMMA that consumes it.
            const unsigned tiles_s = (unsigned)__cvta_generic_to_shared(tiles);
            const unsigned a_thr = 8u * (C::A_KC ? (wm + g) * C::LDK + t4 : t4 * C::LDAM + wm + g);
            const unsigned b_thr = 8u * (C::A_ELEMS + (C::B_KC ? (wn + g) * C::LDK + t4 : t4 * C::LDBN + wn + g));
            constexpr int A_MI = 8 * (C::A_KC ? 8 * C::LDK : 8), A_KK = 8 * (C::A_KC ? 4 : 4 * C::LDAM);
            constexpr int B_NI = 8 * (C::B_KC ? 8 * C::LDK : 8), B_KK = 8 * (C::B_KC ? 4 : 4 * C::LDBN);
            auto stage_addr = [&](int r) { return tiles_s + (unsigned)(r % STAGES) * (unsigned)(C::STAGE_ELEMS * 8); };
            double a[2][MF], b[2][NF];
            // load fragment number f (0..MF+NF-1) of k-step KK into buffer BUF
            auto load_frag = [&](unsigned st, auto kk_c, auto buf_c, auto f_c) {
                constexpr int KK = decltype(kk_c)::value, BUF = decltype(buf_c)::value, F = decltype(f_c)::value;
                if constexpr (F < MF) a[BUF][F] = lds_f64<F * A_MI + KK * A_KK>(st + a_thr);
                else b[BUF][F - MF] = lds_f64<(F - MF) * B_NI + KK * B_KK>(st + b_thr);
            };
            auto load_frags_all = [&](unsigned st, auto kk_c, auto buf_c) {
                [&]<int... F>(std::integer_sequence<int, F...>) {
                    (load_frag(st, kk_c, buf_c, std::integral_constant<int, F>{}), ...);
                }(std::make_integer_sequence<int, MF + NF>{});
            };
            // the DMMAs of k-step buffer BUF; if PREF, the fragments of (st_next, KKN) are loaded into the other buffer
            // on the way (one LDS.64 after every second DMMA)
            auto mma_step = [&](auto buf_c, auto pref_c, unsigned st_next, auto kkn_c) {
                constexpr int BUF = decltype(buf_c)::value;
                constexpr bool PREF = decltype(pref_c)::value;
                [&]<int... Q>(std::integer_sequence<int, Q...>) {
                    (([&] {
                         constexpr int mi = Q / NF, ni = Q % NF;
                         dmma8x8x4(acc[mi][ni][0], acc[mi][ni][1], a[BUF][mi], b[BUF][ni]);
                         // one prefetch LDS after every second DMMA where the warp tile has enough DMMAs to carry
                         // them (MF*NF/2 >= MF+NF), after every DMMA for the narrow tiles
                         constexpr int EVERY = (MF * NF / 2 >= MF + NF) ? 2 : 1;
                         if constexpr (PREF && (Q % EVERY == EVERY - 1) && (Q / EVERY < MF + NF))
                             load_frag(st_next, kkn_c, std::integral_constant<int, 1 - BUF>{}, std::integral_constant<int, Q / EVERY>{});
                     }()),
                     ...);
                }(std::make_integer_sequence<int, MF * NF>{});
                static_assert(MF * NF >= MF + NF, "not enough DMMA slots to carry the next fragments");
            };
            using I0 = std::integral_constant<int, 0>;
            using I1 = std::integral_constant<int, 1>;
            using I2 = std::integral_constant<int, 2>;
            using I3 = std::integral_constant<int, 3>;
            static_assert(BK == 16, "the k-step schedule below is written for 4 DMMA k-steps per stage");

            // invariant at the top of every step: the stage is full and its first fragments are in buffer 0
            mbar_wait(full + ring % STAGES, (ring / STAGES) & 1);
            load_frags_all(stage_addr(ring), I0{}, I0{});
            int c_ks = 0;
            for (int j = 0; j < total_steps; ++j, ++ring) {
                const unsigned st = stage_addr(ring);
                const bool more = j + 1 < total_steps;
                const int kk_lim = (c_ks == steps_per_pair - 1) ? last_kk : BK / 4;
                if (++c_ks == steps_per_pair) c_ks = 0;
                if (kk_lim == BK / 4) {
                    mma_step(I0{}, std::true_type{}, st, I1{});
                    mma_step(I1{}, std::true_type{}, st, I2{});
                    mma_step(I0{}, std::true_type{}, st, I3{});
                    if (more) {
                        // the first fragments of the NEXT stage ride under the last k-step of this one
                        mbar_wait(full + (ring + 1) % STAGES, ((ring + 1) / STAGES) & 1);
                        mma_step(I1{}, std::true_type{}, stage_addr(ring + 1), I0{});
                    } else {
                        mma_step(I1{}, std::false_type{}, st, I0{});
                    }
                } else {
                    // ragged last stage of an operand pair: only the k-steps that hold contracted elements
                    mma_step(I0{}, std::false_type{}, st, I0{});
                    if (kk_lim > 1) {
                        load_frags_all(st, I1{}, I1{});
                        mma_step(I1{}, std::false_type{}, st, I0{});
                    }
                    if (kk_lim > 2) {
                        load_frags_all(st, I2{}, I0{});
                        mma_step(I0{}, std::false_type{}, st, I0{});
                    }
                    if (more) {
                        mbar_wait(full + (ring + 1) % STAGES, ((ring + 1) / STAGES) & 1);
                        load_frags_all(stage_addr(ring + 1), I0{}, I0{});
                    }
                }
                // every LDS of this stage has been consumed by an issued DMMA: hand the slot back
                __syncwarp();
                if (elect_one()) mbar_arrive(empty + ring % STAGES);
            }

            // ---- epilogue: D[perm(m,n)] = alpha*acc (+ beta*D): the output permute of F90:782-785 as a scatter ----
            const int* mOffD = tabs + par * (2 * BM + 2 * BN) + BM;
            const int* nOffD = mOffD + BM + BN;
            double* __restrict__ Dp = info[par].D;
            // all destination offsets first (independent LDS), then the stores: the address of a store never waits on a
            // table load; alpha == 1 (every reference op) skips the FP64 multiply, which shares its pipe with DMMA
            int mo[MF], no[NF][2];
#pragma unroll
            for (int mi = 0; mi < MF; ++mi) mo[mi] = mOffD[wm + mi * 8 + g];
#pragma unroll
            for (int ni = 0; ni < NF; ++ni) {
                no[ni][0] = nOffD[wn + ni * 8 + 2 * t4];
                no[ni][1] = nOffD[wn + ni * 8 + 2 * t4 + 1];
            }
            const bool unit_alpha = alpha == 1.0;
            if (args.atomic) {  // split-K / split-chain partial sums: the launcher has applied beta
#pragma unroll
                for (int mi = 0; mi < MF; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NF; ++ni)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if ((mo[mi] | no[ni][c]) >= 0) {
                                const double v = unit_alpha ? acc[mi][ni][c] : alpha * acc[mi][ni][c];
                                asm volatile("red.global.add.f64 [%0], %1;\n" ::"l"(Dp + (size_t)(mo[mi] + no[ni][c])), "d"(v) : "memory");
                            }
            } else if (beta == 0.0) {
#pragma unroll
                for (int mi = 0; mi < MF; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NF; ++ni)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if ((mo[mi] | no[ni][c]) >= 0)
                                Dp[(size_t)(mo[mi] + no[ni][c])] = unit_alpha ? acc[mi][ni][c] : alpha * acc[mi][ni][c];
            } else {
#pragma unroll
                for (int mi = 0; mi < MF; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NF; ++ni)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if ((mo[mi] | no[ni][c]) >= 0) {
                                double* dst = Dp + (size_t)(mo[mi] + no[ni][c]);
                                *dst = alpha * acc[mi][ni][c] + beta * *dst;
                            }
            }
            __syncwarp();
            if (elect_one()) mbar_arrive(tab_empty + par);
        }
    }
}

template <class C>
int launch_cfg_now(const ContractArgs& a, int max_ctas) {
    static bool attr_set = false;
    if (!attr_set) {
        SIP_CUDA(cudaFuncSetAttribute(contract_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        attr_set = true;
    }
    max_ctas *= C::MIN_CTAS;
    int grid = a.total_tiles < max_ctas ? a.total_tiles : max_ctas;
    if (grid < 1) return SIPGPU_OK;
    ContractArgs args = a;
    int* ctr = nullptr;
    static const bool lockstep_enabled = [] { const char* e = getenv("SIPGPU_LOCKSTEP"); return !e || atoi(e) != 0; }();
    if (args.sync_q > 0 && lockstep_enabled && grid > 1) {
        const size_t rounds = ((size_t)a.total_tiles + grid - 1) / grid;
        const size_t bytes = rounds * (size_t)args.sync_q * sizeof(int);
        ctr = (int*)pool_alloc(bytes);
        if (ctr && cudaMemsetAsync(ctr, 0, bytes, ctx().stream) == cudaSuccess) args.sync_ctr = ctr;
        else args.sync_q = 0;
    } else {
        args.sync_q = 0;
    }
    contract_kernel<C><<<grid, C::NT, C::SMEM, ctx().stream>>>(args);
    if (ctr) pool_free(ctr);  // recycled in stream order
    SIP_CUDA(cudaGetLastError());
    count_launch();
    return SIPGPU_OK;
}
template <class C>
int launch_cfg(const ContractArgs& a, int max_ctas) {
    if (Capture* cap = capture()) {  // prepared launch: replayed by sipgpu_plan_launch
        cap->steps.push_back([a, max_ctas] { return launch_cfg_now<C>(a, max_ctas); });
        return SIPGPU_OK;
    }
    return launch_cfg_now<C>(a, max_ctas);
}

}  // namespace

// Tile menu: 0 = 128x128 (8 consumer warps of 64x32), 1 = 128x80 (8 warps of 32x40: N = 400 = o*o segments tile
// exactly), 2 = 64x64 (4 consumer warps of 32x32, two CTAs per SM): launches with fewer large tiles than SMs (single
// small blocks -- the one-block-per-opcode calls of an unmodified interpreter) and blocks with both free dimensions
// <= 64, 3 = 128x56 and 4 = 128x24 (4 consumer warps of 32x56 / 32x24, two CTAs per SM so that the epilogue of one
// overlaps the DMMAs of the other): the skinny contractions D[a,i,b,j] = L[a,i,c,j]*R[c,b]
// with one segment-sized free dimension (v = 50, o = 20) -- 68 of the 170 SIAL patterns, HBM-bound, where padding
// N = 50 to 80 or 128 would turn a bandwidth-bound op into a (wasted-)compute-bound one.
void contract_tile_dims(int tile, int* bm, int* bn) {
    *bm = tile == 2 ? 64 : 128;
    *bn = tile == 1 ? 80 : tile == 2 ? 64 : tile == 3 ? 56 : tile == 4 ? 24 : 128;
}
long long contract_tile_count(int M, int N, int tile) {
    int bm, bn;
    contract_tile_dims(tile, &bm, &bn);
    return (long long)((M + bm - 1) / bm) * ((N + bn - 1) / bn);
}
int contract_pick_tile(int M, int N) {
    // padded work x a per-tile efficiency penalty (shared-memory fragment loads per DMMA, operand re-reads)
    static const double penalty[5] = {1.0, 1.0, 1.15, 1.10, 1.25};
    double best = -1.0;
    int pick = 0;
    for (int t = 0; t < 5; ++t) {
        int bm, bn;
        contract_tile_dims(t, &bm, &bn);
        if (t >= 3 && N > 64) continue;  // the narrow tiles are for ONE segment-sized n dimension
        const double cost = (double)((M + bm - 1) / bm) * bm * (double)((N + bn - 1) / bn) * bn * penalty[t];
        if (best < 0 || cost < best) { best = cost; pick = t; }  // ties keep the tile listed first
    }
    return pick;
}

template <int WM, int WN, int MF, int NF>
int launch_tile(const ContractArgs& a, bool a_kc, bool b_kc, bool vec, int ctas) {
    if (vec) {
        if (a_kc && b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, true, true, true>>(a, ctas);
        if (a_kc && !b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, true, false, true>>(a, ctas);
        if (!a_kc && b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, false, true, true>>(a, ctas);
        return launch_cfg<Cfg<WM, WN, MF, NF, false, false, true>>(a, ctas);
    }
    if (a_kc && b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, true, true, false>>(a, ctas);
    if (a_kc && !b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, true, false, false>>(a, ctas);
    if (!a_kc && b_kc) return launch_cfg<Cfg<WM, WN, MF, NF, false, true, false>>(a, ctas);
    return launch_cfg<Cfg<WM, WN, MF, NF, false, false, false>>(a, ctas);
}

int launch_contract(const ContractArgs& a, bool a_kc, bool b_kc, bool vec, int tile) {
    const int ctas = ctx().num_sms;
    if (tile == 4) return launch_tile<4, 1, 4, 3>(a, a_kc, b_kc, vec, ctas);
    if (tile == 3) return launch_tile<4, 1, 4, 7>(a, a_kc, b_kc, vec, ctas);
    if (tile == 2) return launch_tile<2, 2, 4, 4>(a, a_kc, b_kc, vec, ctas);
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

// permute.cu -- the SIAL `transpose` super-instruction (block_permute_op, interpreter.cpp:656-679 ->
// Block::transpose_copy block.cpp:216-255 -> tensor_block_copy_ tensor_dil_omp.F90:438-660) on sm_100a.
//
// out[new position of idx] = in[idx].  Pure data movement: 8 B read + 8 B written per element, so the
// roofline is HBM bandwidth.  Design: dims that stay adjacent are collapsed (plan.cpp); the block is cut
// into tiles that contain a contiguous INPUT run (leading input dims) and a contiguous OUTPUT run (leading
// output dims) of >= 32 elements each; a CTA reads a tile in input order (coalesced), parks it in shared
// memory, and writes it in output order (coalesced).  The per-element index arithmetic (the div/mod chains
// of the reference's index walk, F90:531-654) is hoisted into two small offset tables built once per
// (shape, permutation) on the host, cached on the device and shared by every tile and every later call.
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "elementwise.h"
#include "plan.h"

namespace sipgpu {
bool permute_bulk_enabled();
int permute_variant();
int permute_vec_mode();
namespace {

constexpr int kPT = 256;        // threads per CTA
constexpr int kMaxTile = 4096;   // hard cap: input run (< 64) x output run (< 64) always fits
constexpr int kWidenTo = 2048;   // tiles are widened towards this many elements (16.5 KB of shared memory)

struct PermArgs {
    int rank;
    int ntile[kMaxRank];
    int tstep_in[kMaxRank], tstep_out[kMaxRank];
    int V;
    int rag_dim[2], rag_te[2], rag_ext[2];
    long long ntiles;
    const int2* rtab;  // input order:  {in offset, ragged coords}
    const int4* wtab;  // output order: {out offset, position in the tile (input order), ragged coords, -}
};

// the blocks of one launch: n (in, out) pairs that share the plan; n == 1 travels inline (no upload)
struct PermBatch {
    const double* const* in;  // device arrays of n pointers, or nullptr: use in0/out0
    double* const* out;
    const double* in0;
    double* out0;
    int n;
    double alpha, beta;  // ACC variants: out = alpha * permuted(in) + beta * out
};

__device__ __forceinline__ int skew(int e) { return e + (e >> 5); }

constexpr int kNoLimit = 1 << 15;   // ragged tile extents are < 2^15 (build_plan_host), so this bound never bites
constexpr int kInvalid = 0xffff;    // ragged-coordinate marker of a slot beyond the tile volume: fails every bound

// A CTA is persistent over (block, tile) work items.  The position of an element inside a tile is the same in every
// tile, so each thread reads ITS table entries (EPT elements: offsets in the input, in the output and in the staged
// tile) ONCE into registers; the tile loop itself is table-free: EPT independent coalesced loads in flight per
// thread, one barrier, EPT coalesced stores.  (The first version re-read the tables per element and tile: every
// global load hung off a table load, and the kernel sat at 60 % of the copy bandwidth on long-scoreboard stalls.)
template <int EPT, bool RAG, bool ACC>
__global__ void __launch_bounds__(kPT, (EPT <= 4 ? 4 : EPT <= 8 ? 3 : 2) - (RAG || ACC ? 1 : 0)) permute_kernel(const __grid_constant__ PermArgs a, const __grid_constant__ PermBatch b) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x;
    int r_off[EPT], w_off[EPT], w_pos[EPT], r_rag[RAG ? EPT : 1], w_rag[RAG ? EPT : 1];
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        const int e = tid + u * kPT;
        r_off[u] = w_off[u] = w_pos[u] = 0;
        if (RAG) r_rag[u] = w_rag[u] = kInvalid;
        if (e < a.V) {
            const int2 r = __ldg(a.rtab + e);
            const int4 w = __ldg(a.wtab + e);
            r_off[u] = r.x;
            w_off[u] = w.x;
            w_pos[u] = skew(w.y);
            if (RAG) { r_rag[u] = r.y; w_rag[u] = w.z; }
        }
    }
    // is slot u of this thread a real element of the tile (and, for ragged tiles, inside the block)?
    auto r_ok = [&](int u, int lim0, int lim1) {
        if constexpr (RAG) return (r_rag[u] & 0xffff) < lim0 && (r_rag[u] >> 16) < lim1;
        else return tid + u * kPT < a.V;
    };
    auto w_ok = [&](int u, int lim0, int lim1) {
        if constexpr (RAG) return (w_rag[u] & 0xffff) < lim0 && (w_rag[u] >> 16) < lim1;
        else return tid + u * kPT < a.V;
    };
    const long long total = a.ntiles * b.n;
    // (block, tile) work item -> source / destination base and the ragged limits of the tile
    // 32-bit index arithmetic: a block has < 2^31 elements, so its tile count fits an int, and the (block, tile) split is
    // one 32-bit division when the whole launch has < 2^31 work items (64-bit div/mod chains cost ~1000 cycles per tile,
    // 12 % of a 2048-element tile)
    const unsigned ntiles32 = (unsigned)a.ntiles;
    auto decode = [&](long long work, const double*& src, double*& dst, int& lim0, int& lim1) {
        const int blk = total < (1LL << 31) ? (int)((unsigned)work / ntiles32) : (int)(work / a.ntiles);
        unsigned t = (unsigned)(work - (long long)blk * a.ntiles);
        int bin = 0, bout = 0;
        lim0 = lim1 = kNoLimit;
#pragma unroll 1
        for (int d = 0; d < a.rank; ++d) {
            const unsigned q = t / (unsigned)a.ntile[d];
            const int c = (int)(t - q * (unsigned)a.ntile[d]);
            t = q;
            bin += c * a.tstep_in[d];
            bout += c * a.tstep_out[d];
            if (RAG) {
                if (d == a.rag_dim[0]) lim0 = min(a.rag_te[0], a.rag_ext[0] - c * a.rag_te[0]);
                if (d == a.rag_dim[1]) lim1 = min(a.rag_te[1], a.rag_ext[1] - c * a.rag_te[1]);
            }
        }
        src = (b.in ? b.in[blk] : b.in0) + bin;
        dst = (b.in ? b.out[blk] : b.out0) + bout;
    };
    // Software pipeline over the persistent tile loop: the loads of tile t+1 are issued BEFORE the stores of tile t (the
    // registers v[] are free as soon as tile t is parked in shared memory), so every CTA always has a tile's worth of
    // loads in flight while it writes -- the load -> barrier -> store -> barrier sequence no longer serialises them.
    long long work = blockIdx.x;
    const double* src = nullptr;
    double* dst = nullptr;
    int lim0 = kNoLimit, lim1 = kNoLimit;
    double v[EPT];
    if (work < total) {
        decode(work, src, dst, lim0, lim1);
#pragma unroll
        for (int u = 0; u < EPT; ++u)
            if (r_ok(u, lim0, lim1)) v[u] = __ldg(src + r_off[u]);
    }
    while (work < total) {
#pragma unroll
        for (int u = 0; u < EPT; ++u) sm[skew(tid + u * kPT)] = v[u];
        __syncthreads();
        double* __restrict__ cdst = dst;
        const int clim0 = lim0, clim1 = lim1;
        const long long next = work + gridDim.x;
        if (next < total) {
            decode(next, src, dst, lim0, lim1);
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (r_ok(u, lim0, lim1)) v[u] = __ldg(src + r_off[u]);
        }
        if (ACC) {
            double o[EPT];
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const bool ok = w_ok(u, clim0, clim1);
                if (ok) o[u] = cdst[w_off[u]];
            }
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const bool ok = w_ok(u, clim0, clim1);
                if (ok) cdst[w_off[u]] = b.alpha * sm[w_pos[u]] + b.beta * o[u];
            }
        } else {
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const bool ok = w_ok(u, clim0, clim1);
                if (ok) cdst[w_off[u]] = sm[w_pos[u]];
            }
        }
        __syncthreads();
        work = next;
    }
}

// ---- TMA variant: tiles are fetched by bulk copies (cp.async.bulk, SASS UBLKCP) issued by one warp into a ring of kBulkStages
// staged tiles guarded by full / empty mbarriers; the threads never load from global memory -- they wait for a tile, read it
// from shared memory in OUTPUT order and store it (coalesced), while the copies of the next kBulkStages - 1 tiles are in
// flight.  Registers hold only the output-side table entries.
constexpr int kBulkStages = 4;
struct BulkArgs {
    const int2* runtab;
    const int4* wtab;
    int nruns, te0, rs, stage_elems;
};
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(double* smem_dst, const double* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <int EPT, bool RAG, bool ACC>
__global__ void __launch_bounds__(kPT, (EPT <= 6 && !(RAG && ACC)) ? 3 : 2) permute_bulk_kernel(const __grid_constant__ PermArgs a, const __grid_constant__ PermBatch b,
                                                              const __grid_constant__ BulkArgs k) {
    extern __shared__ __align__(16) double sm[];
    __shared__ unsigned long long full[kBulkStages], empty[kBulkStages];
    __shared__ double* s_dst[kBulkStages];
    __shared__ int s_lim[kBulkStages][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kBulkStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kPT / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    int w_off[EPT], w_pos[EPT], w_rag[RAG ? EPT : 1];
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        const int e = tid + u * kPT;
        w_off[u] = w_pos[u] = 0;
        if (RAG) w_rag[u] = kInvalid;
        if (e < a.V) {
            const int4 w = __ldg(k.wtab + e);
            w_off[u] = w.x;
            w_pos[u] = w.y;
            if (RAG) w_rag[u] = w.z;
        }
    }
    auto w_ok = [&](int u, int lim0, int lim1) {
        if constexpr (RAG) return (w_rag[u] & 0xffff) < lim0 && (w_rag[u] >> 16) < lim1;
        else return tid + u * kPT < a.V;
    };
    __syncthreads();
    const long long total = a.ntiles * b.n;
    const unsigned ntiles32 = (unsigned)a.ntiles;
    auto decode = [&](long long work, const double*& src, double*& dst, int& lim0, int& lim1) {
        const int blk = total < (1LL << 31) ? (int)((unsigned)work / ntiles32) : (int)(work / a.ntiles);
        unsigned t = (unsigned)(work - (long long)blk * a.ntiles);
        int bin = 0, bout = 0;
        lim0 = lim1 = kNoLimit;
#pragma unroll 1
        for (int d = 0; d < a.rank; ++d) {
            const unsigned q = t / (unsigned)a.ntile[d];
            const int c = (int)(t - q * (unsigned)a.ntile[d]);
            t = q;
            bin += c * a.tstep_in[d];
            bout += c * a.tstep_out[d];
            if (RAG) {
                if (d == a.rag_dim[0]) lim0 = min(a.rag_te[0], a.rag_ext[0] - c * a.rag_te[0]);
                if (d == a.rag_dim[1]) lim1 = min(a.rag_te[1], a.rag_ext[1] - c * a.rag_te[1]);
            }
        }
        src = (b.in ? b.in[blk] : b.in0) + bin;
        dst = (b.in ? b.out[blk] : b.out0) + bout;
    };
    // warp 0: fetch item number j of this CTA (work = blockIdx.x + j * gridDim.x) into stage j % kBulkStages
    auto produce = [&](long long j) {
        const long long work = blockIdx.x + j * (long long)gridDim.x;
        if (work >= total) return;
        const int s = (int)(j % kBulkStages);
        mbar_wait(empty + s, (int)((j / kBulkStages) & 1) ^ 1);   // everybody has read what the stage held (first lap: at once)
        const double* src;
        double* dst;
        int lim0, lim1;
        decode(work, src, dst, lim0, lim1);
        double* stage = sm + (size_t)s * k.stage_elems;
        unsigned bytes = 0;
        for (int rr = lane; rr < k.nruns; rr += 32) {
            const int2 r = __ldg(k.runtab + rr);
            int len = k.te0;
            bool ok = true;
            if (RAG) {   // a ragged dimension 0 shortens the run, any other one drops whole runs
                if (a.rag_dim[0] == 0) len = min(len, lim0); else ok = ok && (r.y & 0xffff) < lim0;
                if (a.rag_dim[1] == 0) len = min(len, lim1); else ok = ok && (r.y >> 16) < lim1;
            }
            if (ok && len > 0) {
                bulk_g2s(stage + rr * k.rs, src + r.x, (unsigned)len * 8u, full + s);
                bytes += (unsigned)len * 8u;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
        if (lane == 0) {
            s_dst[s] = dst;
            s_lim[s][0] = lim0;
            s_lim[s][1] = lim1;
            mbar_arrive_expect_tx(full + s, bytes);   // the one pending arrival gates the phase until the byte count is known
        }
    };
    if (warp == 0)
        for (int j = 0; j < kBulkStages - 1; ++j) produce(j);
    for (long long j = 0;; ++j) {
        const long long work = blockIdx.x + j * (long long)gridDim.x;
        if (work >= total) break;
        if (warp == 0) produce(j + kBulkStages - 1);
        const int s = (int)(j % kBulkStages);
        mbar_wait(full + s, (int)((j / kBulkStages) & 1));
        const double* stage = sm + (size_t)s * k.stage_elems;
        double* __restrict__ cdst = s_dst[s];
        const int lim0 = s_lim[s][0], lim1 = s_lim[s][1];
        if (ACC) {
            double o[EPT];
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (w_ok(u, lim0, lim1)) o[u] = cdst[w_off[u]];
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (w_ok(u, lim0, lim1)) cdst[w_off[u]] = b.alpha * stage[w_pos[u]] + b.beta * o[u];
        } else {
#pragma unroll
            for (int u = 0; u < EPT; ++u)
                if (w_ok(u, lim0, lim1)) cdst[w_off[u]] = stage[w_pos[u]];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
}

// ---- cp.async ring variant: a tile is fetched straight into a ring of kRingStages staged tiles in shared memory (LDGSTS, no
// registers in between), so every CTA keeps kRingStages - 1 whole tiles of loads in flight while it stores the tile that has
// landed; ONE barrier per tile.  VLD: 16-byte cp.async when the input runs are even and 16-byte aligned; VST: 16-byte global
// stores (two 8-byte shared-memory reads of output-adjacent elements) when the output runs are.  Registers hold table entries only.
constexpr int kRingStages = 3;
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
// The staged position of every tile element comes from the plan (ptab / wtab.y, .w): the host searches a padded layout per
// (shape, permutation) that keeps the output-order reads of a warp off each other's banks (ring_layout below).

constexpr int ring_min_ctas(int ept, bool heavy) { return (ept <= 4 ? 4 : ept <= 8 ? 3 : 2) - (heavy ? 1 : 0); }   // heavy: ragged tiles or fused accumulate (more registers)

__device__ __forceinline__ void cp_async8s(unsigned saddr, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16s(unsigned saddr, const double* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ double lds_f64(unsigned saddr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(saddr) : "memory");
    return v;
}

// Instruction budget of the tile loop: every shared-memory address is a 32-bit byte address (stage base + the thread's byte
// offset from the plan table), the validity of a thread's slots is one bit mask when the tile is not ragged -- about 10 warp
// instructions per 32 elements moved (the first version spent 51 on 64-bit generic address arithmetic and parameter reloads).
template <int EPT, bool RAG, bool ACC, bool VLD, bool VST>
__global__ void __launch_bounds__(kPT, ring_min_ctas(EPT, RAG || ACC)) permute_ring_kernel(const __grid_constant__ PermArgs a, const __grid_constant__ PermBatch b, const int* __restrict__ ptab,
                                                                               const int4* __restrict__ wtab, int stage_elems) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NL = VLD ? EPT / 2 : EPT;   // load units per thread (elements or pairs)
    constexpr int NS = VST ? EPT / 2 : EPT;   // store units per thread
    const int tid = threadIdx.x;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
    const unsigned stage_bytes = (unsigned)stage_elems * 8u;
    int r_off[NL], r_rag[RAG ? NL : 1];
    unsigned r_pb[NL];                          // staged position of the load unit, in bytes
    int w_off[NS], w_rag[RAG ? NS : 1];
    unsigned w_pb0[NS], w_pb1[VST ? NS : 1];    // staged position(s) of the store unit, in bytes
    unsigned rmask = 0, wmask = 0;              // which slots of this thread are real elements of the tile
#pragma unroll
    for (int u = 0; u < NL; ++u) {
        const int e = (tid + u * kPT) * (VLD ? 2 : 1);
        r_off[u] = 0;
        r_pb[u] = 0;
        if (RAG) r_rag[u] = kInvalid;
        if (e < a.V) {
            const int2 r = __ldg(a.rtab + e);
            r_off[u] = r.x;
            r_pb[u] = (unsigned)__ldg(ptab + e) * 8u;
            if (RAG) r_rag[u] = r.y;
            rmask |= 1u << u;
        }
    }
#pragma unroll
    for (int u = 0; u < NS; ++u) {
        const int e = (tid + u * kPT) * (VST ? 2 : 1);
        w_off[u] = 0;
        w_pb0[u] = 0;
        if (VST) w_pb1[u] = 0;
        if (RAG) w_rag[u] = kInvalid;
        if (e < a.V) {
            const int4 w = __ldg(wtab + e);
            w_off[u] = w.x;
            w_pb0[u] = (unsigned)w.y * 8u;
            if (RAG) w_rag[u] = w.z;
            if (VST) w_pb1[u] = (unsigned)w.w * 8u;
            wmask |= 1u << u;
        }
    }
    auto r_ok = [&](int u, int lim0, int lim1) {
        if constexpr (RAG) return (r_rag[u] & 0xffff) < lim0 && (r_rag[u] >> 16) < lim1;
        else return (rmask >> u) & 1u;
    };
    auto w_ok = [&](int u, int lim0, int lim1) {
        if constexpr (RAG) return (w_rag[u] & 0xffff) < lim0 && (w_rag[u] >> 16) < lim1;
        else return (wmask >> u) & 1u;
    };
    const long long total = a.ntiles * b.n;
    const unsigned ntiles32 = (unsigned)a.ntiles;
    auto decode = [&](long long work, const double*& src, double*& dst, int& lim0, int& lim1) {
        const int blk = total < (1LL << 31) ? (int)((unsigned)work / ntiles32) : (int)(work / a.ntiles);
        unsigned t = (unsigned)(work - (long long)blk * a.ntiles);
        int bin = 0, bout = 0;
        lim0 = lim1 = kNoLimit;
#pragma unroll 1
        for (int d = 0; d < a.rank; ++d) {
            const unsigned q = t / (unsigned)a.ntile[d];
            const int c = (int)(t - q * (unsigned)a.ntile[d]);
            t = q;
            bin += c * a.tstep_in[d];
            bout += c * a.tstep_out[d];
            if (RAG) {
                if (d == a.rag_dim[0]) lim0 = min(a.rag_te[0], a.rag_ext[0] - c * a.rag_te[0]);
                if (d == a.rag_dim[1]) lim1 = min(a.rag_te[1], a.rag_ext[1] - c * a.rag_te[1]);
            }
        }
        src = (b.in ? b.in[blk] : b.in0) + bin;
        dst = (b.in ? b.out[blk] : b.out0) + bout;
    };
    // destination and ragged limits of the tiles in flight: a register ring shifted once per tile (kRingStages == 3)
    double *d0 = nullptr, *d1 = nullptr, *d2 = nullptr;
    int l00 = kNoLimit, l01 = kNoLimit, l10 = kNoLimit, l11 = kNoLimit, l20 = kNoLimit, l21 = kNoLimit;
    auto issue = [&](long long work, unsigned sstage, double*& dst, int& lim0, int& lim1) {
        if (work < total) {
            const double* src;
            decode(work, src, dst, lim0, lim1);
#pragma unroll
            for (int u = 0; u < NL; ++u)
                if (r_ok(u, lim0, lim1)) {
                    if (VLD) cp_async16s(sstage + r_pb[u], src + r_off[u]);
                    else cp_async8s(sstage + r_pb[u], src + r_off[u]);
                }
        }
        cp_async_commit();
    };
    const long long w0 = blockIdx.x, step = gridDim.x;
    issue(w0, sbase, d0, l00, l01);
    issue(w0 + step, sbase + stage_bytes, d1, l10, l11);
    int stage = 0;
    const double alpha = b.alpha, beta = b.beta;
    for (long long work = w0; work < total; work += step) {
        cp_async_wait<kRingStages - 2>();   // this thread's copies of the current tile have landed ...
        __syncthreads();                     // ... everybody's have, and everybody is done reading the stage refilled next
        int nstage = stage + 2;
        if (nstage >= kRingStages) nstage -= kRingStages;
        issue(work + 2 * step, sbase + (unsigned)nstage * stage_bytes, d2, l20, l21);
        const unsigned st = sbase + (unsigned)stage * stage_bytes;
        double* __restrict__ cdst = d0;
        if constexpr (VST) {
            if (ACC) {
                double2 o[NS];
#pragma unroll
                for (int u = 0; u < NS; ++u)
                    if (w_ok(u, l00, l01)) o[u] = *reinterpret_cast<const double2*>(cdst + w_off[u]);
#pragma unroll
                for (int u = 0; u < NS; ++u)
                    if (w_ok(u, l00, l01)) {
                        double2 v;
                        v.x = alpha * lds_f64(st + w_pb0[u]) + beta * o[u].x;
                        v.y = alpha * lds_f64(st + w_pb1[u]) + beta * o[u].y;
                        *reinterpret_cast<double2*>(cdst + w_off[u]) = v;
                    }
            } else {
#pragma unroll
                for (int u = 0; u < NS; ++u)
                    if (w_ok(u, l00, l01)) {
                        double2 v;
                        v.x = lds_f64(st + w_pb0[u]);
                        v.y = lds_f64(st + w_pb1[u]);
                        *reinterpret_cast<double2*>(cdst + w_off[u]) = v;
                    }
            }
        } else {
            if (ACC) {
                double o[NS];
#pragma unroll
                for (int u = 0; u < NS; ++u)
                    if (w_ok(u, l00, l01)) o[u] = cdst[w_off[u]];
#pragma unroll
                for (int u = 0; u < NS; ++u)
                    if (w_ok(u, l00, l01)) cdst[w_off[u]] = alpha * lds_f64(st + w_pb0[u]) + beta * o[u];
            } else {
#pragma unroll
                for (int u = 0; u < NS; ++u)
                    if (w_ok(u, l00, l01)) cdst[w_off[u]] = lds_f64(st + w_pb0[u]);
            }
        }
        d0 = d1; l00 = l10; l01 = l11;
        d1 = d2; l10 = l20; l11 = l21;
        if (++stage == kRingStages) stage = 0;
    }
    cp_async_wait<0>();
}

template <int EPT, bool VLD, bool VST>
int launch_perm_ring2(const PermArgs& a, const PermBatch& b, const int* ptab, const int4* wtab, bool acc, int grid, size_t smem, int stage_elems,
                      cudaStream_t st) {
    const bool rag = a.rag_dim[0] >= 0;
    static bool attr_set = false;
    if (!attr_set) {
        SIP_CUDA(cudaFuncSetAttribute(permute_ring_kernel<EPT, true, true, VLD, VST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        SIP_CUDA(cudaFuncSetAttribute(permute_ring_kernel<EPT, true, false, VLD, VST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        SIP_CUDA(cudaFuncSetAttribute(permute_ring_kernel<EPT, false, true, VLD, VST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        SIP_CUDA(cudaFuncSetAttribute(permute_ring_kernel<EPT, false, false, VLD, VST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        attr_set = true;
    }
    if (rag) {
        if (acc) permute_ring_kernel<EPT, true, true, VLD, VST><<<grid, kPT, smem, st>>>(a, b, ptab, wtab, stage_elems);
        else permute_ring_kernel<EPT, true, false, VLD, VST><<<grid, kPT, smem, st>>>(a, b, ptab, wtab, stage_elems);
    } else {
        if (acc) permute_ring_kernel<EPT, false, true, VLD, VST><<<grid, kPT, smem, st>>>(a, b, ptab, wtab, stage_elems);
        else permute_ring_kernel<EPT, false, false, VLD, VST><<<grid, kPT, smem, st>>>(a, b, ptab, wtab, stage_elems);
    }
    return SIPGPU_OK;
}
template <int EPT>
int launch_perm_ring(const PermArgs& a, const PermBatch& b, const int* ptab, const int4* wtab, bool acc, bool vld, bool vst, int grid, size_t smem,
                     int stage_elems, cudaStream_t st) {
    if (vld) return vst ? launch_perm_ring2<EPT, true, true>(a, b, ptab, wtab, acc, grid, smem, stage_elems, st)
                        : launch_perm_ring2<EPT, true, false>(a, b, ptab, wtab, acc, grid, smem, stage_elems, st);
    return vst ? launch_perm_ring2<EPT, false, true>(a, b, ptab, wtab, acc, grid, smem, stage_elems, st)
               : launch_perm_ring2<EPT, false, false>(a, b, ptab, wtab, acc, grid, smem, stage_elems, st);
}

template <int EPT>
int launch_perm_bulk(const PermArgs& a, const PermBatch& b, const BulkArgs& k, bool acc, int grid, size_t smem, cudaStream_t st) {
    const bool rag = a.rag_dim[0] >= 0;
    static bool attr_set = false;
    if (!attr_set) {
        SIP_CUDA(cudaFuncSetAttribute(permute_bulk_kernel<EPT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        SIP_CUDA(cudaFuncSetAttribute(permute_bulk_kernel<EPT, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        SIP_CUDA(cudaFuncSetAttribute(permute_bulk_kernel<EPT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        SIP_CUDA(cudaFuncSetAttribute(permute_bulk_kernel<EPT, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    if (rag) {
        if (acc) permute_bulk_kernel<EPT, true, true><<<grid, kPT, smem, st>>>(a, b, k);
        else permute_bulk_kernel<EPT, true, false><<<grid, kPT, smem, st>>>(a, b, k);
    } else {
        if (acc) permute_bulk_kernel<EPT, false, true><<<grid, kPT, smem, st>>>(a, b, k);
        else permute_bulk_kernel<EPT, false, false><<<grid, kPT, smem, st>>>(a, b, k);
    }
    return SIPGPU_OK;
}

template <int EPT>
int launch_perm(const PermArgs& a, const PermBatch& b, bool acc, int grid, size_t smem, cudaStream_t st) {
    const bool rag = a.rag_dim[0] >= 0;
    if (rag) {
        if (acc) permute_kernel<EPT, true, true><<<grid, kPT, smem, st>>>(a, b);
        else permute_kernel<EPT, true, false><<<grid, kPT, smem, st>>>(a, b);
    } else {
        if (acc) permute_kernel<EPT, false, true><<<grid, kPT, smem, st>>>(a, b);
        else permute_kernel<EPT, false, false><<<grid, kPT, smem, st>>>(a, b);
    }
    return SIPGPU_OK;
}

inline int skew_host(int e) { return e + (e >> 5); }
// d = alpha * s + beta * d for the degenerate (identity) permutation
int ew_axpby(double* d, const double* s, long long n, double alpha, double beta) {
    if (beta == 0.0) return ew_scale_copy(d, s, n, alpha);
    if (beta != 1.0) SIP_TRY(ew_scale(d, n, beta));
    return ew_axpy(d, s, n, alpha);
}

// TMA path (cp.async.bulk): a tile arrives as `nruns` bulk copies of `te0` contiguous input elements each, laid out run by run
// in shared memory (run stride te0 + 2 doubles: 16-byte aligned, and consecutive runs start two banks apart)
struct BulkPlan {
    bool ok = false;
    const int2* runtab = nullptr;   // [nruns] {input offset of the run inside the tile, ragged coordinates of the run}
    const int4* wtab = nullptr;     // output order: {out offset, position in the staged tile, ragged coords, -}
    int nruns = 0, te0 = 0, rs = 0;
};
struct PlanEntry {
    PermArgs args;
    BulkPlan bulk;
    bool vld_ok = false, vst_ok = false;   // ring variant: 16-byte loads / stores possible (given 16-byte aligned blocks)
    const int* ring_ptab = nullptr;        // ring variant: staged position of tile element e (input order)
    const int4* ring_wtab = nullptr;       //   output order: {out offset, staged position, ragged coords, staged position of the next element}
    int ring_stage_elems = 0;
};

// Staged layout of a tile for the ring variant: p(e) = e + pad1 * (e / R1) + pad2 * (e / R2), R = the tile strides of the dimensions
// (where a warp's output-order walk jumps), pads searched so that the 8-byte shared-memory reads of every half-warp spread over
// the 16 bank pairs.  `even`: 16-byte staging keeps pairs aligned (even pads only; same-parity reads are then 2-way conflicted at best).
// sp_of[e_out] = input-order index of output-order element e_out; pairs: the threads read elements (2q, 2q + 1).
static long long ring_read_cost(const std::vector<int>& pos, const std::vector<int>& sp_of, bool pairs) {
    const int V = (int)sp_of.size();
    long long cost = 0;
    const int unit = pairs ? 2 : 1;
    for (int part = 0; part < unit; ++part)
        for (int base = 0; base * unit < V; base += 16) {   // one half-warp: 16 consecutive units
            int cnt[16] = {0};
            int worst = 0;
            for (int l = 0; l < 16; ++l) {
                const int e = (base + l) * unit + part;
                if (e >= V) break;
                const int c = ++cnt[pos[sp_of[e]] & 15];
                if (c > worst) worst = c;
            }
            cost += worst;
        }
    return cost;
}
static void ring_layout(int rank, const int* te, const std::vector<int>& sp_of, bool even, bool pairs, std::vector<int>& pos, int* stage_elems) {
    const int V = (int)sp_of.size();
    std::vector<int> strides;
    { int s = 1; for (int d = 0; d < rank; ++d) { if (s > 1 && te[d] > 1) strides.push_back(s); s *= te[d]; } }
    strides.push_back(32);
    std::sort(strides.begin(), strides.end());
    strides.erase(std::unique(strides.begin(), strides.end()), strides.end());
    const int step = even ? 2 : 1;
    auto fill = [&](int r1, int p1, int r2, int p2, std::vector<int>& out) {
        out.resize(V);
        for (int e = 0; e < V; ++e) out[e] = e + p1 * (e / r1) + (r2 ? p2 * (e / r2) : 0);
    };
    long long best = -1;
    int b_r1 = 32, b_p1 = 0, b_r2 = 0, b_p2 = 0;
    std::vector<int> cand;
    auto consider = [&](int r1, int p1, int r2, int p2) {
        if (even && ((r1 & 1) || (r2 & 1))) return;   // a pad may only fall between pairs
        fill(r1, p1, r2, p2, cand);
        if ((long long)cand[V - 1] + 2 > (long long)V + V / 4 + 64) return;   // bounded shared-memory overhead
        const long long c = ring_read_cost(cand, sp_of, pairs);
        if (best < 0 || c < best) { best = c; b_r1 = r1; b_p1 = p1; b_r2 = r2; b_p2 = p2; }
    };
    consider(32, 0, 0, 0);
    for (int r1 : strides)
        for (int p1 = step; p1 <= 8; p1 += step) consider(r1, p1, 0, 0);
    const long long ideal = (V / (pairs ? 2 : 1) + 15) / 16 * (pairs ? 2 : 1) * ((even || pairs) ? 2 : 1);
    if (best > ideal + ideal / 8)   // two levels of padding
        for (size_t i = 0; i < strides.size(); ++i)
            for (size_t j = i + 1; j < strides.size(); ++j)
                for (int p1 = 0; p1 <= 8; p1 += step)
                    for (int p2 = step; p2 <= 8; p2 += step) consider(strides[i], p1, strides[j], p2);
    fill(b_r1, b_p1, b_r2, b_p2, pos);
    *stage_elems = (pos[V - 1] + 2 + 1) & ~1;
}
std::unordered_map<std::string, PlanEntry>& cache() {
    static std::unordered_map<std::string, PlanEntry> c;
    return c;
}

int build_plan_host(const PermShape& ps, PermArgs* out, std::vector<int2>& rt, std::vector<int4>& wt, int* te_out = nullptr) {
    const int r = ps.rank;
    int oorder[kMaxRank];  // dims by ascending output stride
    for (int i = 0; i < r; ++i) oorder[i] = i;
    for (int i = 1; i < r; ++i)
        for (int j = i; j > 0 && ps.out_stride[oorder[j]] < ps.out_stride[oorder[j - 1]]; --j) std::swap(oorder[j], oorder[j - 1]);
    int te[kMaxRank];
    for (int i = 0; i < r; ++i) te[i] = 1;
    auto vol = [&]() { long long v = 1; for (int i = 0; i < r; ++i) v *= te[i]; return v; };
    // grow a run (over `order`) until it holds >= need elements
    auto grow_run = [&](const int* order, int need) {
        long long cur = 1;
        for (int i = 0; i < r && cur < need; ++i) {
            const int d = order[i];
            int c = (int)((need + cur - 1) / cur);
            if (c > ps.ext[d]) c = ps.ext[d];
            // a short dimension is taken whole (run still <= 64 elements): full-length contiguous runs and no ragged
            // remainder tile, e.g. extent 50 is one run of 50 rather than 32 + 18
            if (c < ps.ext[d] && cur * ps.ext[d] <= 64) c = ps.ext[d];
            // otherwise prefer the next divisor of the extent (no ragged remainder) while the run stays <= 64
            for (int q = c; q < ps.ext[d] && cur * q <= 64; ++q)
                if (ps.ext[d] % q == 0) { c = q; break; }
            if (c > te[d]) te[d] = c;
            cur *= te[d];
        }
    };
    int iorder[kMaxRank];
    for (int i = 0; i < r; ++i) iorder[i] = i;
    grow_run(iorder, 32);
    grow_run(oorder, 32);
    // widen towards ~kMaxTile: alternately lengthen the input run and the output run
    auto widen = [&](const int* order) {
        for (int i = 0; i < r; ++i) {
            const int d = order[i];
            if (te[d] < ps.ext[d]) {
                int c = te[d] * 2;
                if (c > ps.ext[d]) c = ps.ext[d];
                // step to a divisor of the extent when one is near (keeps the dimension free of ragged tiles)
                for (int q = te[d] + 1; q <= 3 * te[d] && q <= ps.ext[d]; ++q)
                    if (ps.ext[d] % q == 0 && ps.ext[d] % te[d] == 0) { c = q; break; }
                if (vol() / te[d] * c > kWidenTo) return false;
                te[d] = c;
                return true;
            }
        }
        return false;
    };
    for (int it = 0; it < 32; ++it) {
        bool a = widen(iorder), b = widen(oorder);
        if (!a && !b) break;
    }
    // a tile that overflows the cap (pathological extents) falls back to shrinking the slowest tile dims
    while (vol() > kMaxTile) {
        int d = -1;
        for (int i = r - 1; i >= 0; --i) if (te[i] > 1) { d = i; break; }
        if (d < 0) break;
        te[d] = (te[d] + 1) / 2;
    }
    const int V = (int)vol();

    PermArgs a;
    memset(&a, 0, sizeof(a));
    a.rank = r;
    a.V = V;
    a.rag_dim[0] = a.rag_dim[1] = -1;
    long long ntiles = 1;
    int nr = 0;
    for (int d = 0; d < r; ++d) {
        a.ntile[d] = (ps.ext[d] + te[d] - 1) / te[d];
        a.tstep_in[d] = te[d] * ps.in_stride[d];
        a.tstep_out[d] = te[d] * ps.out_stride[d];
        ntiles *= a.ntile[d];
        if (ps.ext[d] % te[d] != 0) {
            if (nr >= 2 || te[d] >= (1 << 15)) { nr = 3; break; }
            a.rag_dim[nr] = d; a.rag_te[nr] = te[d]; a.rag_ext[nr] = ps.ext[d];
            ++nr;
        }
    }
    if (nr > 2) return SIPGPU_E_ARG;  // cannot happen with the growth rules above (<= 2 frontier dims); guarded anyway
    a.ntiles = ntiles;

    rt.assign(V, make_int2(0, 0));
    wt.assign(V, make_int4(0, 0, 0, 0));
    int tin[kMaxRank];  // stride of dim d in the tile's input-order linearisation
    { int s = 1; for (int d = 0; d < r; ++d) { tin[d] = s; s *= te[d]; } }
    auto pack = [&](const int* idx) {
        int p = 0;
        if (a.rag_dim[0] >= 0) p |= idx[a.rag_dim[0]];
        if (a.rag_dim[1] >= 0) p |= idx[a.rag_dim[1]] << 16;
        return p;
    };
    for (int e = 0; e < V; ++e) {
        int idx[kMaxRank], x = e, off = 0;
        for (int d = 0; d < r; ++d) { idx[d] = x % te[d]; x /= te[d]; off += idx[d] * ps.in_stride[d]; }
        rt[e] = make_int2(off, pack(idx));
    }
    for (int e = 0; e < V; ++e) {
        int idx[kMaxRank], x = e, off = 0, sp = 0;
        for (int i = 0; i < r; ++i) {
            const int d = oorder[i];
            idx[d] = x % te[d]; x /= te[d];
            off += idx[d] * ps.out_stride[d];
            sp += idx[d] * tin[d];
        }
        wt[e] = make_int4(off, sp, pack(idx), 0);
    }
    if (te_out)
        for (int d = 0; d < kMaxRank; ++d) te_out[d] = d < r ? te[d] : 1;
    *out = a;
    return SIPGPU_OK;
}

int build_plan(const PermShape& ps, PlanEntry* out) {
    PermArgs a;
    std::vector<int2> rt;
    std::vector<int4> wt;
    int te[kMaxRank];
    SIP_TRY(build_plan_host(ps, &a, rt, wt, te));
    const int V = a.V;
    // ---- TMA path: dimension 0 is the input's contiguous one; every run (te[0] elements) must be a whole number of
    // 16-byte units at a 16-byte aligned address: even tile extent and even extent along it (then every other input stride,
    // a product of extents that contains ext[0], is even as well) ----
    BulkPlan bp;
    if (ps.in_stride[0] == 1 && te[0] >= 4 && te[0] % 2 == 0 && ps.ext[0] % 2 == 0 && V % te[0] == 0 && V / te[0] <= 1024) {
        bp.te0 = te[0];
        bp.rs = te[0] + 2;
        bp.nruns = V / te[0];
        std::vector<int2> runs((size_t)bp.nruns);
        for (int rr = 0; rr < bp.nruns; ++rr) runs[rr] = rt[(size_t)rr * te[0]];   // first element of the run: {offset, ragged coords}
        std::vector<int4> wb(wt);
        for (int e = 0; e < V; ++e) {   // staged position: input-order index sp -> run sp / te0, element sp % te0
            const int sp = wt[e].y;
            wb[e].y = (sp / te[0]) * bp.rs + sp % te[0];
        }
        int2* d_runs = reinterpret_cast<int2*>(pool_alloc(sizeof(int2) * runs.size(), true));
        int4* d_wb = reinterpret_cast<int4*>(pool_alloc(sizeof(int4) * wb.size(), true));
        if (!d_runs || !d_wb) return SIPGPU_E_NOMEM;
        SIP_CUDA(cudaMemcpyAsync(d_runs, runs.data(), sizeof(int2) * runs.size(), cudaMemcpyHostToDevice, ctx().stream));
        SIP_CUDA(cudaMemcpyAsync(d_wb, wb.data(), sizeof(int4) * wb.size(), cudaMemcpyHostToDevice, ctx().stream));
        SIP_CUDA(cudaStreamSynchronize(ctx().stream));
        bp.runtab = d_runs;
        bp.wtab = d_wb;
        bp.ok = true;
    }
    out->bulk = bp;
    {   // 16-byte accesses of the ring variant: the tile extent and the block extent along the contiguous dimension must be even
        // (then every other stride and every tile step on that side is even as well)
        int o0 = 0;
        for (int d = 0; d < ps.rank; ++d)
            if (ps.out_stride[d] < ps.out_stride[o0]) o0 = d;
        out->vld_ok = ps.in_stride[0] == 1 && te[0] % 2 == 0 && ps.ext[0] % 2 == 0;
        out->vst_ok = ps.out_stride[o0] == 1 && te[o0] % 2 == 0 && ps.ext[o0] % 2 == 0;
        std::vector<int> sp_of((size_t)V), pos;
        for (int e = 0; e < V; ++e) sp_of[e] = wt[e].y;
        ring_layout(ps.rank, te, sp_of, out->vld_ok, out->vst_ok, pos, &out->ring_stage_elems);
        std::vector<int4> wr(wt);
        for (int e = 0; e < V; ++e) {
            wr[e].y = pos[wt[e].y];
            wr[e].w = e + 1 < V ? pos[wt[e + 1].y] : 0;
        }
        int* d_pt = reinterpret_cast<int*>(pool_alloc(sizeof(int) * V, true));
        int4* d_wr = reinterpret_cast<int4*>(pool_alloc(sizeof(int4) * V, true));
        if (!d_pt || !d_wr) return SIPGPU_E_NOMEM;
        SIP_CUDA(cudaMemcpyAsync(d_pt, pos.data(), sizeof(int) * V, cudaMemcpyHostToDevice, ctx().stream));
        SIP_CUDA(cudaMemcpyAsync(d_wr, wr.data(), sizeof(int4) * V, cudaMemcpyHostToDevice, ctx().stream));
        SIP_CUDA(cudaStreamSynchronize(ctx().stream));
        out->ring_ptab = d_pt;
        out->ring_wtab = d_wr;
    }
    int2* d_rt = reinterpret_cast<int2*>(pool_alloc(sizeof(int2) * V, true));
    int4* d_wt = reinterpret_cast<int4*>(pool_alloc(sizeof(int4) * V, true));
    if (!d_rt || !d_wt) return SIPGPU_E_NOMEM;
    // plan tables are built once per (shape, permutation); a blocking upload keeps the host vectors simple
    SIP_CUDA(cudaMemcpyAsync(d_rt, rt.data(), sizeof(int2) * V, cudaMemcpyHostToDevice, ctx().stream));
    SIP_CUDA(cudaMemcpyAsync(d_wt, wt.data(), sizeof(int4) * V, cudaMemcpyHostToDevice, ctx().stream));
    SIP_CUDA(cudaStreamSynchronize(ctx().stream));
    a.rtab = d_rt;
    a.wtab = d_wt;
    out->args = a;
    return SIPGPU_OK;
}

}  // namespace

void permute_cache_clear() { cache().clear(); }

static int& permute_bulk_flag() {
    static int v = [] { const char* e = getenv("SIPGPU_PERMUTE_BULK"); return e ? atoi(e) : 3; }();
    return v;
}
bool permute_bulk_enabled() { return permute_bulk_flag() == 1; }
int permute_variant() { return permute_bulk_flag(); }
static int& permute_vec_flag() {
    static int v = [] { const char* e = getenv("SIPGPU_PERMUTE_VEC"); return e ? atoi(e) : -1; }();
    return v;
}
int permute_vec_mode() { return permute_vec_flag(); }   // ring variant, 16-byte cp.async / stores where eligible: 1 always, 0 never, -1 by tile size
void permute_set_vec(int on) { permute_vec_flag() = on; }   // 0: register-staged tiles, 1: TMA bulk copies, 2: cp.async ring
void permute_set_bulk(int on) { permute_bulk_flag() = on; }

// Host-only view of the plan for CPU tests of the tiling logic (no device needed).
// meta = {rank, V, ntiles, rag_dim0, rag_te0, rag_ext0, rag_dim1, rag_te1, rag_ext1, ntile[6], tstep_in[6], tstep_out[6]}
int permute_plan_debug(int rank, const int* ext, const int* transp, long long* meta, int* rtab, int* wtab, int cap) {
    PermShape ps;
    SIP_TRY(build_perm_shape(rank, ext, transp, &ps));
    meta[0] = ps.rank;
    if (ps.rank <= 1) { meta[1] = 0; meta[2] = 0; return SIPGPU_OK; }
    PermArgs a;
    std::vector<int2> rt;
    std::vector<int4> wt;
    SIP_TRY(build_plan_host(ps, &a, rt, wt));
    if (a.V > cap) return SIPGPU_E_ARG;
    meta[1] = a.V; meta[2] = a.ntiles;
    for (int i = 0; i < 2; ++i) { meta[3 + 3 * i] = a.rag_dim[i]; meta[4 + 3 * i] = a.rag_te[i]; meta[5 + 3 * i] = a.rag_ext[i]; }
    for (int d = 0; d < kMaxRank; ++d) { meta[9 + d] = a.ntile[d]; meta[15 + d] = a.tstep_in[d]; meta[21 + d] = a.tstep_out[d]; }
    for (int e = 0; e < a.V; ++e) {
        rtab[2 * e] = rt[e].x; rtab[2 * e + 1] = rt[e].y;
        wtab[3 * e] = wt[e].x; wtab[3 * e + 1] = wt[e].y; wtab[3 * e + 2] = wt[e].z;
    }
    return SIPGPU_OK;
}

// n blocks of identical shape and permutation: out_i = alpha * permuted(in_i) + beta * out_i in ONE launch.
// alpha = 1, beta = 0 is the plain transpose (tensor_block_copy_); beta != 0 is the fused permute-accumulate of
// `X[b,j,a,i] += R[a,i,b,j]` bodies (interpreter.cpp:1874-1997 permutes into a temp block and adds).
int permute_batched(int n, int rank, const int* ext, const int* transp, const double* const* in, double* const* out,
                    double alpha, double beta) {
    if (n < 0 || rank < 0 || !in || !out) return SIPGPU_E_ARG;
    if (n == 0) return SIPGPU_OK;
    SIP_TRY(ensure_init());
    Ctx& c = ctx();
    for (int i = 0; i < n; ++i)
        if (!in[i] || !out[i]) return SIPGPU_E_ARG;
    const bool acc = !(alpha == 1.0 && beta == 0.0);
    PermShape ps;
    if (rank == 0) { ps.rank = 0; ps.total = 1; }
    else SIP_TRY(build_perm_shape(rank, ext, transp, &ps));
    if (ps.rank <= 1) {  // identity after collapsing: straight copy (F90:478-491) or axpby
        for (int i = 0; i < n; ++i) {
            const double* src = in[i];
            double* dst = out[i];
            const long long cnt = ps.total;
            auto go = [src, dst, cnt, acc, alpha, beta]() -> int {
                if (!acc) SIP_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)cnt, cudaMemcpyDeviceToDevice, ctx().stream));
                else SIP_TRY(ew_axpby(dst, src, cnt, alpha, beta));
                return SIPGPU_OK;
            };
            if (Capture* cap = capture()) cap->steps.push_back(go);
            else SIP_TRY(go());
        }
        return SIPGPU_OK;
    }
    int keyv[1 + 3 * kMaxRank] = {ps.rank};
    for (int i = 0; i < kMaxRank; ++i) {
        keyv[1 + i] = i < ps.rank ? ps.ext[i] : 0;
        keyv[1 + kMaxRank + i] = i < ps.rank ? ps.in_stride[i] : 0;
        keyv[1 + 2 * kMaxRank + i] = i < ps.rank ? ps.out_stride[i] : 0;
    }
    std::string key(reinterpret_cast<const char*>(keyv), sizeof(keyv));
    auto it = cache().find(key);
    if (it == cache().end()) {
        PlanEntry pe;
        SIP_TRY(build_plan(ps, &pe));
        it = cache().emplace(key, pe).first;
    }
    const PermArgs& a = it->second.args;
    PermBatch b;
    memset(&b, 0, sizeof(b));
    b.n = n;
    b.alpha = alpha;
    b.beta = beta;
    if (n == 1) {
        b.in0 = in[0];
        b.out0 = out[0];
    } else {
        void *h, *d;
        SIP_TRY(desc_alloc(sizeof(void*) * 2 * (size_t)n, &h, &d));
        memcpy(h, in, sizeof(void*) * n);
        memcpy((char*)h + sizeof(void*) * n, out, sizeof(void*) * n);
        SIP_TRY(desc_commit(h, d, sizeof(void*) * 2 * (size_t)n));
        b.in = (const double* const*)d;
        b.out = (double* const*)((char*)d + sizeof(void*) * n);
    }
    int ept = 16;
    for (int cand : {4, 6, 8, 10, 12})
        if (a.V <= cand * kPT) { ept = cand; break; }
    // Kernel choice by evidence (profiles/r02_permute_variants.txt, all 23 rank-4 permutations at 16^4 .. 64^4 and 50x20x50x20):
    // plain permutes run fastest through the cp.async ring with 8-byte staging (conflict-free searched layout); tiles of more
    // than 2560 elements (EPT >= 12: long unmerged leading runs, e.g. 64-element first dimension) want the 16-byte loads / stores
    // (+20 %: half the instructions and table registers); the fused accumulate is faster register-staged except on those big
    // tiles; TMA bulk copies only win (+4 %) when the leading dimensions stay merged -- not selected automatically.
    const int variant = permute_variant();
    const bool big = ept >= 12;
    const bool use_ring = variant == 2 || (variant == 3 && (!acc || big));
    if (use_ring) {   // ---- cp.async ring ----
        bool vld = it->second.vld_ok, vst = it->second.vst_ok;
        for (int i = 0; i < n && (vld || vst); ++i) {
            vld = vld && (((uintptr_t)in[i]) & 15) == 0;
            vst = vst && (((uintptr_t)out[i]) & 15) == 0;
        }
        if (permute_vec_mode() == 0 || (permute_vec_mode() < 0 && !big)) vld = vst = false;
        const int stage_elems = it->second.ring_stage_elems;
        const int* ptab = it->second.ring_ptab;
        const int4* wtab = it->second.ring_wtab;
        const size_t smem = sizeof(double) * (size_t)kRingStages * stage_elems;
        long long per_sm = std::max<long long>(1, std::min<long long>(ring_min_ctas(ept, acc || a.rag_dim[0] >= 0), (long long)(220 * 1024) / (long long)(smem + 1024)));
        long long grid = a.ntiles * n;
        if (grid > c.num_sms * per_sm) grid = c.num_sms * per_sm;
        auto go = [a, b, ptab, wtab, acc, vld, vst, ept, grid, smem, stage_elems]() -> int {
            cudaStream_t st = ctx().stream;
            switch (ept) {
                case 4: SIP_TRY(launch_perm_ring<4>(a, b, ptab, wtab, acc, vld, vst, (int)grid, smem, stage_elems, st)); break;
                case 6: SIP_TRY(launch_perm_ring<6>(a, b, ptab, wtab, acc, vld, vst, (int)grid, smem, stage_elems, st)); break;
                case 8: SIP_TRY(launch_perm_ring<8>(a, b, ptab, wtab, acc, vld, vst, (int)grid, smem, stage_elems, st)); break;
                case 10: SIP_TRY(launch_perm_ring<10>(a, b, ptab, wtab, acc, vld, vst, (int)grid, smem, stage_elems, st)); break;
                case 12: SIP_TRY(launch_perm_ring<12>(a, b, ptab, wtab, acc, vld, vst, (int)grid, smem, stage_elems, st)); break;
                default: SIP_TRY(launch_perm_ring<16>(a, b, ptab, wtab, acc, vld, vst, (int)grid, smem, stage_elems, st)); break;
            }
            SIP_CUDA(cudaGetLastError());
            count_launch();
            return SIPGPU_OK;
        };
        if (Capture* cap = capture()) { cap->steps.push_back(go); return SIPGPU_OK; }
        return go();
    }
    // ---- TMA path: bulk copies need 16-byte aligned runs, i.e. 16-byte aligned blocks on top of the plan's conditions ----
    const BulkPlan& bp = it->second.bulk;
    bool bulk = bp.ok && variant == 1;
    for (int i = 0; i < n && bulk; ++i) bulk = (((uintptr_t)in[i]) & 15) == 0;
    if (bulk) {
        BulkArgs k;
        k.runtab = bp.runtab;
        k.wtab = bp.wtab;
        k.nruns = bp.nruns;
        k.te0 = bp.te0;
        k.rs = bp.rs;
        k.stage_elems = bp.nruns * bp.rs;
        const size_t smem = sizeof(double) * (size_t)kBulkStages * k.stage_elems;
        if (smem <= 160 * 1024) {
            const long long reg_cap = (ept <= 6 && !(acc && a.rag_dim[0] >= 0)) ? 3 : 2;
            long long per_sm = std::max<long long>(1, std::min<long long>(reg_cap, (long long)(210 * 1024) / (long long)(smem + 2048)));
            long long grid = a.ntiles * n;
            if (grid > c.num_sms * per_sm) grid = c.num_sms * per_sm;
            auto go = [a, b, k, acc, ept, grid, smem]() -> int {
                cudaStream_t st = ctx().stream;
                switch (ept) {
                    case 4: SIP_TRY(launch_perm_bulk<4>(a, b, k, acc, (int)grid, smem, st)); break;
                    case 6: SIP_TRY(launch_perm_bulk<6>(a, b, k, acc, (int)grid, smem, st)); break;
                    case 8: SIP_TRY(launch_perm_bulk<8>(a, b, k, acc, (int)grid, smem, st)); break;
                    case 10: SIP_TRY(launch_perm_bulk<10>(a, b, k, acc, (int)grid, smem, st)); break;
                    case 12: SIP_TRY(launch_perm_bulk<12>(a, b, k, acc, (int)grid, smem, st)); break;
                    default: SIP_TRY(launch_perm_bulk<16>(a, b, k, acc, (int)grid, smem, st)); break;
                }
                SIP_CUDA(cudaGetLastError());
                count_launch();
                return SIPGPU_OK;
            };
            if (Capture* cap = capture()) { cap->steps.push_back(go); return SIPGPU_OK; }
            return go();
        }
    }
    const size_t smem = sizeof(double) * (size_t)(skew_host(ept * kPT) + 1);  // every thread parks all its EPT slots
    // persistent CTAs: as many as fit per SM (registers: 4-6 of 256 threads; shared memory: 227 KB / tile)
    long long per_sm = (long long)(200 * 1024) / (long long)(smem + 1024);
    const long long reg_cap = (ept <= 4 ? 4 : ept <= 8 ? 3 : 2) - ((acc || a.rag_dim[0] >= 0) ? 1 : 0);
    if (per_sm > reg_cap) per_sm = reg_cap;
    if (per_sm < 1) per_sm = 1;
    long long grid = a.ntiles * n;
    if (grid > c.num_sms * per_sm) grid = c.num_sms * per_sm;
    auto go = [a, b, acc, ept, grid, smem]() -> int {
        cudaStream_t st = ctx().stream;
        switch (ept) {
            case 4: SIP_TRY(launch_perm<4>(a, b, acc, (int)grid, smem, st)); break;
            case 6: SIP_TRY(launch_perm<6>(a, b, acc, (int)grid, smem, st)); break;
            case 8: SIP_TRY(launch_perm<8>(a, b, acc, (int)grid, smem, st)); break;
            case 10: SIP_TRY(launch_perm<10>(a, b, acc, (int)grid, smem, st)); break;
            case 12: SIP_TRY(launch_perm<12>(a, b, acc, (int)grid, smem, st)); break;
            default: SIP_TRY(launch_perm<16>(a, b, acc, (int)grid, smem, st)); break;
        }
        SIP_CUDA(cudaGetLastError());
        count_launch();
        return SIPGPU_OK;
    };
    if (Capture* cap = capture()) { cap->steps.push_back(go); return SIPGPU_OK; }
    return go();
}

int permute_block(int rank, const int* ext, const int* transp, const double* in, double* out) {
    if (rank < 0 || !in || !out) return SIPGPU_E_ARG;
    return permute_batched(1, rank, ext, transp, &in, &out, 1.0, 0.0);
}

}  // namespace sipgpu

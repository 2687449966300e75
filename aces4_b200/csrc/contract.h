// contract.h -- launch interface of the fused DMMA contraction kernel (contract.cu).
#pragma once
#include <vector>

#include "plan.h"

namespace sipgpu {

struct ContractArgs {
    const Problem* probs;    // device work-list, or nullptr: use the inline p0/pair0/s0 (single block, no upload)
    const Pair* pairs;       // device
    const Shape* shapes;     // device
    const int* tile_prefix;  // device, nprob+1 entries (exclusive prefix sum of tiles per block)
    int nprob;
    int total_tiles;
    double alpha, beta;
    int atomic;  // epilogue adds alpha*acc with red.global.add.f64 (several problems share a destination: split-K)
    // Lockstep throttle (uniform work-lists only): every kSyncEvery ring stages the producers of all CTAs of a tile round
    // announce their progress in sync_ctr[round * sync_q + step] and wait until the whole round has reached the step
    // kSyncWindow steps back, so CTAs that share operand panels stay within a few hundred k-elements of each other and
    // the panels are fetched from DRAM once instead of once per drifted-apart CTA.  sync_q = steps per tile, 0 = off.
    int sync_q;
    int* sync_ctr;
    Problem p0;
    Pair pair0;
    Shape s0;
};

// CTA tile of the kernel variant (for the host-side tile count)
void contract_tile_dims(int tile, int* bm, int* bn);
// variant = (a_kc, b_kc): which operand tiles are K-contiguous; vec: 16-byte cp.async (Shape::vec and 16-byte
// aligned operand pointers)
int launch_contract(const ContractArgs& a, bool a_kc, bool b_kc, bool vec, int tile);
int contract_pick_tile(int M, int N);
long long contract_tile_count(int M, int N, int tile);
constexpr int kContractKWin = 2048;  // contracted elements per k window (Cfg::KWIN)
constexpr int kSyncEvery = 32;   // ring stages (of 16 contracted elements) between two lockstep points
constexpr int kSyncWindow = 2;   // allowed lead, in lockstep points
constexpr int kSmallTile = 2;  // 64x64, for launches that cannot fill the SMs with large tiles
int dmma_probe(int iters, double* tflops);

// lowint.cu: the bandwidth-shaped kernel for low arithmetic intensity contractions (N <= 64 after the launcher's operand
// swap, flops per algorithmic byte below the roofline ridge)
bool lowint_eligible(const Shape& s);
int lowint_launch(const Shape& s, int n, const std::vector<Pair>& pairs, const std::vector<int>& chain, double* const* D,
                  double alpha, double beta, bool dense_d);
void lowint_cache_clear();
void lowint_set_max_intensity(double flops_per_byte);  // negative: the kernel is never chosen
void lowint_set_slab(int on);   // TMA-fed slab path for small results of long contractions (default on)
void lowint_set_scope(int scope);  // 0 never, 1 where it wins (default), 2 every shape with N <= 64 below the intensity limit

}  // namespace sipgpu

// worklist.h -- deferred op stream: record the per-block ops a pardo body issues one opcode at a time, then
// schedule them as a handful of batched launches (worklist.cu).  SURVEY.md section 8f row 3.
#pragma once
#include <functional>
#include <initializer_list>

#include "common.h"

namespace sipgpu {

enum WlMode { WL_R = 1, WL_W = 2, WL_RW = 3, WL_ATOMIC = 4 };
struct WlRange {
    const void* p;
    size_t bytes;
    int mode;
};

// elementwise op codes of the batched kernel (same arithmetic as elementwise.cu)
enum WlEwOp { WL_FILL = 0, WL_SCALE, WL_SCALE_COPY, WL_INCR, WL_AXPY, WL_ADDSUB, WL_REDADD, WL_REDINCR };

bool wl_active();  // recording (entry points divert into the wl_rec_* functions)
bool wl_dry();     // recording without a device (host-only planning, CPU tests)
// Schedule and launch everything recorded so far; recording stays on.  Blocking entry points call this first.
int wl_flush();
void wl_replay_cache_clear();  // drops every captured stream (finalize; descriptors live in pool memory)
void wl_set_copy_bulk(int on);  // whole-block copies as TMA bulk transfers (-1: environment / default on)
void wl_tuning_changed();       // launch policy changed: captured streams of the old policy must not be replayed

// exclusive: a WL_REDADD whose destination no other process can reach (single-rank array) -- the scheduler may turn it
// into the fused accumulate of the producing contraction
int wl_rec_ew(int op, double* d, const double* a, const double* b, long long n, double f, bool exclusive = false);
int wl_rec_permute(int rank, const int* ext, const int* transp, const double* in, double* out, double alpha, double beta);
int wl_rec_contract(const int* ptrn, const double* L, int lrank, const int* lext, const double* R, int rrank,
                    const int* rext, double* D, int drank, const int* dext, double alpha, double beta);
// any other asynchronous op: runs `fn` at its scheduled position; `ranges` declare what it reads / writes
int wl_rec_opaque(std::function<int()> fn, std::initializer_list<WlRange> ranges);
// pool hooks: blocks allocated while recording are known temporaries; frees are deferred to the flush
double* wl_alloc(long long n, int zero);
int wl_free(double* p);

// launchers implemented in abi.cu / worklist.cu that the scheduler calls
int contract_chained(int n, const int* ptrn, int lrank, int rrank, int drank, const int* lext, const int* rext,
                     const int* dext, const int* chain_start, const double* const* L, const double* const* R,
                     double* const* D, double alpha, double beta);

}  // namespace sipgpu

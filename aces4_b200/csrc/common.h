// common.h -- shared state, error handling and launch bookkeeping of libsipgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <functional>
#include <vector>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sipgpu.h"

namespace sipgpu {

constexpr int kMaxRank = SIPGPU_MAX_RANK;
constexpr int kNumSMs = 148;  // B200

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

// Evaluate a CUDA runtime call; on failure record the message and return SIPGPU_E_CUDA from the caller.
#define SIP_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return ::sipgpu::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define SIP_TRY(call)                 \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != SIPGPU_OK) return rc__; \
    } while (0)

struct Ctx {
    bool inited = false;
    int device = -1;
    int num_sms = kNumSMs;
    cudaStream_t stream = nullptr;       // compute stream: every kernel of the library runs here
    cudaStream_t copy_stream = nullptr;  // staging copies that may overlap compute
    long long launches = 0;              // kernels launched by this library
    // small pinned + device rings for descriptors / tables of batched launches
    void* h_scratch = nullptr;
    void* d_scratch = nullptr;
    size_t scratch_bytes = 0;
    size_t scratch_off = 0;
    double* d_reduce = nullptr;  // reduction partials + result
    double* h_reduce = nullptr;  // pinned
};
Ctx& ctx();
int ensure_init();  // lazily initialises on device 0 (or SIPGPU_DEVICE / LOCAL_RANK) -- fails without a GPU

// Reserve `bytes` in the descriptor rings; returns host and device views of the same slot.  The slot is
// valid until the ring wraps (the ring synchronises the stream when it wraps).
int scratch_reserve(size_t bytes, void** h, void** d);

inline void count_launch(int n = 1) { ctx().launches += n; }

// ---- prepared launches (sipgpu_plan_*): while a Capture is open, descriptor uploads go to device memory the capture owns
// (not the ring, whose slots are recycled) and the kernel launch sites push a closure instead of launching; replaying the
// closures re-issues exactly the same launches with no host-side marshalling.
struct Capture {
    std::vector<std::function<int()>> steps;
    std::vector<void*> device;  // pool blocks holding descriptors, freed with the capture
    bool failed = false;        // a launch site that cannot be captured was reached: the plan falls back to the eager path
};
Capture*& capture();
// descriptor memory of one launch: ring slot (eager) or capture-owned (capturing); commit = the host->device copy
int desc_alloc(size_t bytes, void** h, void** d);
int desc_commit(void* h, void* d, size_t bytes);

// ---- per-entry-point trace (sipgpu_trace_*): the counterpart of the reference's Tracer (src/sip/worker/tracer.h:41-50, per-pc
// call counts and wall time of the interpreter loop): calls and host wall time per C-ABI entry point.  Off by default; one
// predictable branch per call when off.
extern bool g_trace_on;
int trace_slot(const char* name);              // registers `name` once (call sites keep the slot in a static)
void trace_add(int slot, double host_seconds);
struct TraceScope {
    int slot;
    double t0;
    explicit TraceScope(int s);
    ~TraceScope();
};
#define SIP_TRACE(name)                                                   \
    static const int sip_trace_slot__ = ::sipgpu::trace_slot(name);       \
    ::sipgpu::TraceScope sip_trace_scope__(::sipgpu::g_trace_on ? sip_trace_slot__ : -1)

// bumped by every device allocation made OUTSIDE the pool (distributed-array slabs): cached free-memory figures are stale then
long long& mem_epoch();

// device pool (pool.cu)
double* pool_alloc(size_t bytes, bool fresh = false);   // fresh: never a recycled block (long-lived plan tables)
int pool_free(void* p);

}  // namespace sipgpu

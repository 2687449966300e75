// core.cu -- library state: device selection, streams, error text, descriptor rings and the resident
// device block pool.
//
// The pool replaces the reference's per-block heap traffic on the hot path: `new double[size]` per temp
// block per loop iteration (block.cpp:28-79), three allocate/deallocate pairs inside every contraction
// (tensor_dil_omp.F90:731-751,788-790) and cudaMalloc/cudaFree per block and 3 x 40 MB per contraction in
// the legacy GPU backend (gpu_super_instructions.cu:183-204,400-402).  Blocks come from large HBM arenas
// (bump allocation) and are recycled through exact-size free lists -- SIAL programs use a handful of
// distinct block sizes, so a freed block is almost always reused by the next allocation of the same shape.
#include <stdarg.h>
#include <stdlib.h>
#include <time.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include <algorithm>
#include <functional>

#include "common.h"
#include "worklist.h"

namespace sipgpu {

namespace {
char g_err[512] = "";
Ctx g_ctx;

struct Arena {
    char* base;
    size_t size, used;
};
struct Pool {
    std::vector<Arena> arenas;
    std::unordered_map<size_t, std::vector<void*>> free_lists;  // rounded size -> blocks
    std::unordered_map<void*, size_t> live;                     // block -> rounded size
    size_t reserved = 0, in_use = 0;
    size_t default_arena = (size_t)256 << 20;
} g_pool;

constexpr size_t kAlign = 512;
inline size_t round_up(size_t b) { return (b + kAlign - 1) / kAlign * kAlign; }

int add_arena(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("device pool: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return SIPGPU_E_NOMEM;
    }
    g_pool.arenas.push_back({(char*)p, bytes, 0});
    g_pool.reserved += bytes;
    return SIPGPU_OK;
}
}  // namespace

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("%s:%d: %s -> %s", file, line, what, cudaGetErrorString(e));
    cudaGetLastError();  // clear the sticky flag of non-fatal errors
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return SIPGPU_E_NODEVICE;
    return SIPGPU_E_CUDA;
}

Ctx& ctx() { return g_ctx; }

static int init_on(int device) {
    Ctx& c = g_ctx;
    if (c.inited) {
        if (device >= 0 && device != c.device) {
            set_error("sipgpu already initialised on device %d (asked for %d)", c.device, device);
            return SIPGPU_E_STATE;
        }
        return SIPGPU_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); libsipgpu has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return SIPGPU_E_NODEVICE;
    }
    if (device < 0) {
        const char* env = getenv("SIPGPU_DEVICE");
        if (!env) env = getenv("LOCAL_RANK");
        device = env ? atoi(env) % ndev : 0;
    }
    if (device >= ndev) {
        set_error("device %d out of range (%d devices)", device, ndev);
        return SIPGPU_E_ARG;
    }
    SIP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SIP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; libsipgpu is built for sm_100a only", device, prop.major, prop.minor);
        return SIPGPU_E_NODEVICE;
    }
    c.device = device;
    c.num_sms = prop.multiProcessorCount;
    SIP_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    SIP_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    c.scratch_bytes = (size_t)8 << 20;
    SIP_CUDA(cudaMallocHost(&c.h_scratch, c.scratch_bytes));
    SIP_CUDA(cudaMalloc(&c.d_scratch, c.scratch_bytes));
    c.scratch_off = 0;
    SIP_CUDA(cudaMalloc(&c.d_reduce, sizeof(double) * 2048));
    SIP_CUDA(cudaMemset(c.d_reduce, 0, sizeof(double) * 2048));
    {  // d_reduce[2] is a resident 1.0 (unit scalar operand for the scalar-block kernels)
        const double one = 1.0;
        SIP_CUDA(cudaMemcpy(c.d_reduce + 2, &one, sizeof(double), cudaMemcpyHostToDevice));
    }
    SIP_CUDA(cudaMallocHost(&c.h_reduce, sizeof(double) * 16));
    c.inited = true;
    return SIPGPU_OK;
}

int ensure_init() { return g_ctx.inited ? SIPGPU_OK : init_on(-1); }

int scratch_reserve(size_t bytes, void** h, void** d) {
    Ctx& c = g_ctx;
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes > c.scratch_bytes) {
        set_error("descriptor ring too small for %zu bytes", bytes);
        return SIPGPU_E_NOMEM;
    }
    if (c.scratch_off + bytes > c.scratch_bytes) {
        // wrap: everything queued so far must have consumed its descriptors
        SIP_CUDA(cudaStreamSynchronize(c.stream));
        c.scratch_off = 0;
    }
    *h = (char*)c.h_scratch + c.scratch_off;
    *d = (char*)c.d_scratch + c.scratch_off;
    c.scratch_off += bytes;
    return SIPGPU_OK;
}

bool g_trace_on = false;
namespace {
struct TraceRow {
    const char* name;
    long long calls = 0;
    double host_s = 0.0;
};
std::vector<TraceRow>& trace_rows() {
    static std::vector<TraceRow> r;
    return r;
}
double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
}  // namespace
int trace_slot(const char* name) {
    trace_rows().push_back(TraceRow{name});
    return (int)trace_rows().size() - 1;
}
void trace_add(int slot, double host_seconds) {
    TraceRow& r = trace_rows()[(size_t)slot];
    r.calls++;
    r.host_s += host_seconds;
}
TraceScope::TraceScope(int s) : slot(s), t0(s >= 0 ? now_s() : 0.0) {}
TraceScope::~TraceScope() {
    if (slot >= 0) trace_add(slot, now_s() - t0);
}

Capture*& capture() {
    static Capture* c = nullptr;
    return c;
}
int desc_alloc(size_t bytes, void** h, void** d) {
    Capture* cap = capture();
    if (!cap) return scratch_reserve(bytes, h, d);
    *h = malloc(bytes ? bytes : 1);
    *d = pool_alloc(bytes ? bytes : 1, true);   // plan-owned, long-lived: fresh space (see pool_alloc), never a freed temp's slot
    if (!*h || !*d) { free(*h); return SIPGPU_E_NOMEM; }
    cap->device.push_back(*d);
    return SIPGPU_OK;
}
int desc_commit(void* h, void* d, size_t bytes) {
    SIP_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
    if (capture()) {  // pageable staging: the copy has left the buffer when the call returns
        SIP_CUDA(cudaStreamSynchronize(g_ctx.stream));
        free(h);
    }
    return SIPGPU_OK;
}

// fresh: carve new arena space even when a recycled block of the size is free -- for tables that live as long as the library
// (plan caches).  Taking a freed temp's slot would shift the addresses the next identical recording gets and defeat the
// replay cache for one round.
double* pool_alloc(size_t bytes, bool fresh) {
    if (ensure_init() != SIPGPU_OK) return nullptr;
    const size_t sz = round_up(bytes ? bytes : 1);
    auto fl = g_pool.free_lists.find(sz);
    void* p = nullptr;
    if (!fresh && fl != g_pool.free_lists.end() && !fl->second.empty()) {
        // lowest free address first (min-heap): the address a block gets depends only on WHICH blocks are live, not on
        // the order earlier frees happened in -- so the same program point of every CC iteration sees the same addresses,
        // which is what lets the deferred op stream recognise a repeated recording (worklist.cu replay cache)
        std::pop_heap(fl->second.begin(), fl->second.end(), std::greater<void*>());
        p = fl->second.back();
        fl->second.pop_back();
    } else {
        for (auto& a : g_pool.arenas)
            if (a.size - a.used >= sz) { p = a.base + a.used; a.used += sz; break; }
        if (!p) {
            size_t want = sz > g_pool.default_arena ? sz : g_pool.default_arena;
            if (add_arena(want) != SIPGPU_OK) {
                if (want == sz || add_arena(sz) != SIPGPU_OK) return nullptr;
            }
            Arena& a = g_pool.arenas.back();
            p = a.base + a.used;
            a.used += sz;
        }
    }
    g_pool.live[p] = sz;
    g_pool.in_use += sz;
    return (double*)p;
}

int pool_free(void* p) {
    if (!p) return SIPGPU_OK;
    auto it = g_pool.live.find(p);
    if (it == g_pool.live.end()) {
        set_error("sipgpu_block_free: %p is not a live pool block", p);
        return SIPGPU_E_ARG;
    }
    // Stream order makes recycling safe: every kernel that used the block was enqueued on the compute stream
    // before this call, and the next user is enqueued after it.
    auto& fl = g_pool.free_lists[it->second];
    fl.push_back(p);
    std::push_heap(fl.begin(), fl.end(), std::greater<void*>());
    g_pool.in_use -= it->second;
    g_pool.live.erase(it);
    return SIPGPU_OK;
}

void permute_cache_clear();
void lowint_cache_clear();
void wl_replay_cache_clear();

static int finalize_all() {
    Ctx& c = g_ctx;
    if (!c.inited) return SIPGPU_OK;
    cudaStreamSynchronize(c.stream);
    cudaStreamSynchronize(c.copy_stream);
    wl_replay_cache_clear();
    permute_cache_clear();
    lowint_cache_clear();
    for (auto& a : g_pool.arenas) cudaFree(a.base);
    g_pool = Pool();
    cudaFreeHost(c.h_scratch);
    cudaFree(c.d_scratch);
    cudaFree(c.d_reduce);
    cudaFreeHost(c.h_reduce);
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.copy_stream);
    const long long l = c.launches;
    c = Ctx();
    c.launches = l;
    return SIPGPU_OK;
}

}  // namespace sipgpu

using namespace sipgpu;

extern "C" {

int sipgpu_init(int device) {
    SIP_TRACE("sipgpu_init"); return init_on(device); }
int sipgpu_finalize(void) {
    SIP_TRACE("sipgpu_finalize");
    if (sipgpu_wl_recording()) sipgpu_wl_end();
    return finalize_all();
}
int sipgpu_device(void) { return g_ctx.inited ? g_ctx.device : -1; }
const char* sipgpu_last_error(void) { return g_err; }
int sipgpu_sync(void) {
    SIP_TRACE("sipgpu_sync");
    SIP_TRY(wl_flush());
    SIP_TRY(ensure_init());
    SIP_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return SIPGPU_OK;
}
void* sipgpu_stream(void) { return ensure_init() == SIPGPU_OK ? (void*)g_ctx.stream : nullptr; }
long long sipgpu_kernel_launches(void) { return g_ctx.launches; }

int sipgpu_pool_reserve(size_t bytes) {
    SIP_TRACE("sipgpu_pool_reserve");
    SIP_TRY(ensure_init());
    return add_arena(round_up(bytes));
}
double* sipgpu_block_alloc(long long n, int zero) {
    SIP_TRACE("sipgpu_block_alloc");
    if (n < 0) return nullptr;
    if (wl_active()) return wl_alloc(n, zero);  // a temp of the recorded stream; zeroing becomes a recorded fill
    double* p = pool_alloc(sizeof(double) * (size_t)n);
    if (p && zero && n > 0) {
        if (cudaMemsetAsync(p, 0, sizeof(double) * (size_t)n, g_ctx.stream) != cudaSuccess) {
            cudaGetLastError();
            pool_free(p);
            return nullptr;
        }
    }
    return p;
}
int sipgpu_block_free(double* p) {
    SIP_TRACE("sipgpu_block_free"); return wl_active() ? wl_free(p) : pool_free(p); }
}  // extern C (reopened below)
namespace sipgpu { long long& mem_epoch() { static long long e = 0; return e; } }
extern "C" {
int sipgpu_pool_stats(size_t* reserved, size_t* in_use, size_t* n_live) {
    if (reserved) *reserved = g_pool.reserved;
    if (in_use) *in_use = g_pool.in_use;
    if (n_live) *n_live = g_pool.live.size();
    return SIPGPU_OK;
}
int sipgpu_h2d(double* g_dst, const double* h_src, long long n) {
    SIP_TRACE("sipgpu_h2d");
    SIP_TRY(wl_flush());
    SIP_TRY(ensure_init());
    if (n < 0 || (n && (!g_dst || !h_src))) return SIPGPU_E_ARG;
    SIP_CUDA(cudaMemcpyAsync(g_dst, h_src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, g_ctx.stream));
    return SIPGPU_OK;
}
int sipgpu_d2h(double* h_dst, const double* g_src, long long n) {
    SIP_TRACE("sipgpu_d2h");
    SIP_TRY(wl_flush());
    SIP_TRY(ensure_init());
    if (n < 0 || (n && (!h_dst || !g_src))) return SIPGPU_E_ARG;
    SIP_CUDA(cudaMemcpyAsync(h_dst, g_src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, g_ctx.stream));
    SIP_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return SIPGPU_OK;
}
void* sipgpu_host_alloc(size_t bytes) {
    if (ensure_init() != SIPGPU_OK) return nullptr;
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
int sipgpu_host_free(void* p) {
    SIP_TRACE("sipgpu_host_free");
    if (p) SIP_CUDA(cudaFreeHost(p));
    return SIPGPU_OK;
}

// ---- per-entry-point trace ----
int sipgpu_trace_enable(int on) {
    g_trace_on = on != 0;
    return SIPGPU_OK;
}
int sipgpu_trace_reset(void) {
    for (auto& r : trace_rows()) { r.calls = 0; r.host_s = 0.0; }
    return SIPGPU_OK;
}
// rows with at least one call since the last reset, most host time first; returns how many there are (may exceed cap)
int sipgpu_trace_report(int cap, const char** names, long long* calls, double* host_seconds) {
    std::vector<const TraceRow*> rows;
    for (const auto& r : trace_rows())
        if (r.calls > 0) rows.push_back(&r);
    std::sort(rows.begin(), rows.end(), [](const TraceRow* a, const TraceRow* b) { return a->host_s > b->host_s; });
    for (int i = 0; i < (int)rows.size() && i < cap; ++i) {
        if (names) names[i] = rows[i]->name;
        if (calls) calls[i] = rows[i]->calls;
        if (host_seconds) host_seconds[i] = rows[i]->host_s;
    }
    return (int)rows.size();
}

}  // extern "C"

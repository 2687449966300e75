// mirror.cu -- host/device coherence of one block: the lazy_gpu_* transitions of the reference's BlockManager.
//
// Reference behaviour followed (not code): src/sip/dynamic_data/block_manager.cpp:340-441
// (lazy_gpu_{read,write,update}_on_{device,host}) over the four status bits of sip::Block
// (block.h:198-204: onHost, onGPU, dirtyOnHost, dirtyOnGPU).  A block lives on the host, on the device or on both; a
// copy is made only when the side that is about to be used is missing or stale:
//     read_on_device   : bring the device copy up to date;                      afterwards onGPU, !dirtyOnHost
//     write_on_device  : as read (the old contents may be partially kept), and mark   dirtyOnGPU
//     update_on_device : as write
//     read/write/update_on_host : the mirror image
// "dirty on both sides" and "allocated on neither side" are errors in the reference (sip::fail); here they return
// SIPGPU_E_STATE -- except write_on_* of a block that exists nowhere, which creates it on that side (the reference
// replaces the block by a fresh one, :360-363, :411-414).  The transition itself is a pure function
// (sipgpu_mirror_transition) so that it is covered without a device; sipgpu_mirror_* applies it to real buffers.
#include "common.h"
#include "worklist.h"

namespace sipgpu {
enum { ON_HOST = 1, ON_GPU = 2, DIRTY_HOST = 4, DIRTY_GPU = 8 };
enum { OP_READ_DEV = 0, OP_WRITE_DEV, OP_UPDATE_DEV, OP_READ_HOST, OP_WRITE_HOST, OP_UPDATE_HOST };
enum { ACT_NONE = 0, ACT_H2D = 1, ACT_D2H = 2, ACT_ALLOC_DEV_H2D = 3, ACT_ALLOC_HOST_D2H = 4, ACT_NEW_DEV = 5, ACT_NEW_HOST = 6 };
}  // namespace sipgpu

struct sipgpu_mirror {
    double* host = nullptr;
    double* dev = nullptr;
    long long n = 0;
    int bits = 0;
    bool own_host = false;
};

using namespace sipgpu;

extern "C" {

int sipgpu_mirror_transition(int bits, int op, int* new_bits, int* action) {
    SIP_TRACE("sipgpu_mirror_transition");
    if (op < OP_READ_DEV || op > OP_UPDATE_HOST || bits < 0 || bits > 15) return SIPGPU_E_ARG;
    const bool to_dev = op <= OP_UPDATE_DEV;
    const int here = to_dev ? ON_GPU : ON_HOST;
    const int dirty_here = to_dev ? DIRTY_GPU : DIRTY_HOST, dirty_there = to_dev ? DIRTY_HOST : DIRTY_GPU;
    const bool writes = op != OP_READ_DEV && op != OP_READ_HOST;
    const bool fresh_ok = op == OP_WRITE_DEV || op == OP_WRITE_HOST;
    int act = ACT_NONE;
    if (!(bits & ON_HOST) && !(bits & ON_GPU)) {
        if (!fresh_ok) {
            set_error("block allocated neither on host or gpu");
            return SIPGPU_E_STATE;
        }
        act = to_dev ? ACT_NEW_DEV : ACT_NEW_HOST;
        bits = 0;
    } else if (!(bits & here)) {
        act = to_dev ? ACT_ALLOC_DEV_H2D : ACT_ALLOC_HOST_D2H;
    } else if (bits & dirty_there) {
        // (also when the block is dirty on BOTH sides: the reference tests the other side's dirty bit first, so its
        // "dirty on host & gpu" failure branch is unreachable and the other side's copy wins)
        act = to_dev ? ACT_H2D : ACT_D2H;
    }
    bits |= here;
    bits &= ~dirty_there;
    if (writes) bits |= dirty_here;
    if (new_bits) *new_bits = bits;
    if (action) *action = act;
    return SIPGPU_OK;
}

int sipgpu_mirror_create(double* host_or_null, long long n, sipgpu_mirror** out) {
    SIP_TRACE("sipgpu_mirror_create");
    if (n < 0 || !out) return SIPGPU_E_ARG;
    sipgpu_mirror* m = new sipgpu_mirror();
    m->host = host_or_null;
    m->n = n;
    m->bits = host_or_null ? ON_HOST : 0;
    *out = m;
    return SIPGPU_OK;
}
int sipgpu_mirror_destroy(sipgpu_mirror* m) {
    SIP_TRACE("sipgpu_mirror_destroy");
    if (!m) return SIPGPU_OK;
    // inside a recording the ops that touch the device side have not run yet: defer the free like sipgpu_block_free
    if (m->dev) { if (wl_active()) wl_free(m->dev); else pool_free(m->dev); }
    if (m->own_host && m->host) wl_flush();  // recorded ops may still name the pinned host side (h2d / d2h)
    if (m->own_host && m->host) cudaFreeHost(m->host);
    delete m;
    return SIPGPU_OK;
}
int sipgpu_mirror_status(const sipgpu_mirror* m) { return m ? m->bits : -1; }
double* sipgpu_mirror_host_ptr(sipgpu_mirror* m) { return m ? m->host : nullptr; }
double* sipgpu_mirror_device_ptr(sipgpu_mirror* m) { return m ? m->dev : nullptr; }

// applies one lazy_gpu_* transition; returns the pointer of the side that may now be used, NULL on error
double* sipgpu_mirror_access(sipgpu_mirror* m, int op) {
    SIP_TRACE("sipgpu_mirror_access");
    if (!m) return nullptr;
    int nb = 0, act = 0;
    if (sipgpu_mirror_transition(m->bits, op, &nb, &act) != SIPGPU_OK) return nullptr;
    // sipgpu.h promises that blocking calls flush an open recording: a host access must see what the recorded writers
    // of the device side produce, and a copy towards the device must not overtake recorded readers of the old contents
    if ((act != ACT_NONE || op >= OP_READ_HOST) && wl_flush() != SIPGPU_OK) return nullptr;
    if (ensure_init() != SIPGPU_OK) return nullptr;
    const size_t bytes = sizeof(double) * (size_t)(m->n > 0 ? m->n : 1);
    if ((act == ACT_ALLOC_DEV_H2D || act == ACT_NEW_DEV) && !m->dev) {
        m->dev = pool_alloc(bytes);
        if (!m->dev) return nullptr;
        if (act == ACT_NEW_DEV && cudaMemsetAsync(m->dev, 0, bytes, ctx().stream) != cudaSuccess) return nullptr;  // _gpu_allocate zero-fills
    }
    if ((act == ACT_ALLOC_HOST_D2H || act == ACT_NEW_HOST) && !m->host) {
        void* h = nullptr;
        if (cudaMallocHost(&h, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        m->host = (double*)h;
        m->own_host = true;
        if (act == ACT_NEW_HOST) memset(h, 0, bytes);
    }
    if (act == ACT_H2D || act == ACT_ALLOC_DEV_H2D) {
        if (cudaMemcpyAsync(m->dev, m->host, sizeof(double) * (size_t)m->n, cudaMemcpyHostToDevice, ctx().stream) != cudaSuccess) return nullptr;
        // a pageable host buffer may be reused by the caller right away: make the copy complete before returning
        if (!m->own_host && cudaStreamSynchronize(ctx().stream) != cudaSuccess) return nullptr;
    } else if (act == ACT_D2H || act == ACT_ALLOC_HOST_D2H) {
        if (cudaMemcpyAsync(m->host, m->dev, sizeof(double) * (size_t)m->n, cudaMemcpyDeviceToHost, ctx().stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx().stream) != cudaSuccess)
            return nullptr;
    } else if (op >= OP_READ_HOST && (m->bits & ON_GPU)) {
        // the host side is current, but kernels that READ the device copy may still be in flight before a host write
        if (op != OP_READ_HOST && cudaStreamSynchronize(ctx().stream) != cudaSuccess) return nullptr;
    }
    m->bits = nb;
    return op <= OP_UPDATE_DEV ? m->dev : m->host;
}

}  // extern "C"

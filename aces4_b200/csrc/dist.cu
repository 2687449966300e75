// dist.cu -- distributed / served arrays over NVLink peer memory (boundary 4).
//
// Reference protocol being replaced: SialOpsParallel::get / put_replace / put_accumulate
// (sial_ops_parallel.cpp:132-171,232-284,332-408) talk to dedicated single-threaded server ranks over MPI:
// 3 messages + 2 acks per put, a temp buffer and a serial `data_[i] += to_add[i]` loop at the owner
// (server_block.cpp:53-62).  Here there are no servers: every GPU owns the blocks whose number is
// congruent to its rank (block-cyclic, data_distribution.cpp:74-82; numbering array_table.cpp:50-97, LAST
// index fastest), keeps them resident in one HBM slab, and exports the slab through CUDA IPC.  A get is a
// peer read, a put a peer write, and put += a red.global.add.f64 stream straight into the owner's HBM over
// NVLink -- many writers are safe because the adds are atomic and commutative (the reference's arrival
// order is non-deterministic too).  Section barriers are the caller's (stream sync + process barrier).
#include <algorithm>
#include <map>
#include <unordered_map>
#include <vector>

#include "elementwise.h"
#include "worklist.h"

struct sipgpu_array {
    int rank = 0;
    int my_rank = 0, world = 1;
    int nseg[sipgpu::kMaxRank];
    std::vector<int> seg_ext[sipgpu::kMaxRank];
    long long slice[sipgpu::kMaxRank];  // block-number strides, last index fastest
    long long nblocks = 0;
    std::vector<long long> block_off;   // element offset of each block inside its owner's slab
    std::vector<long long> slab_elems;  // per owner
    std::vector<double*> base;          // per owner: slab base as mapped into this process
    std::vector<char> opened;           // base[r] came from cudaIpcOpenMemHandle
    // race detection (distributed_block_consistency.cpp): what THIS rank did to which block since the last barrier
    bool track = false;
    std::unordered_map<long long, int> touched;  // block number -> SIPGPU_ACCESS_* bits
};

using namespace sipgpu;

namespace {
constexpr long long kBlockAlign = 32;  // doubles (256 B): keeps every block 16-byte-vector and sector aligned

int check_idx(const sipgpu_array* a, const int* idx) {
    if (!a || !idx) return SIPGPU_E_ARG;
    for (int i = 0; i < a->rank; ++i)
        if (idx[i] < 1 || idx[i] > a->nseg[i]) return SIPGPU_E_ARG;
    return SIPGPU_OK;
}
}  // namespace

extern "C" {

// host-only layout arithmetic (no device needed): array_table.cpp:50-97, data_distribution.cpp:74-82
long long sipgpu_layout_block_number(int rank, const int* nseg, const int* idx) {
    if (rank < 1 || rank > kMaxRank || !nseg || !idx) return -1;
    long long s = 1, b = 0;
    for (int p = rank - 1; p >= 0; --p) {
        if (idx[p] < 1 || idx[p] > nseg[p]) return -1;
        b += s * (idx[p] - 1);
        s *= nseg[p];
    }
    return b;
}
int sipgpu_layout_block_owner(long long block_number, int world) {
    return (block_number < 0 || world < 1) ? -1 : (int)(block_number % world);
}

int sipgpu_array_create(int rank, const int* nseg, const int* seg_ext, int my_rank, int world, sipgpu_array** out) {
    SIP_TRACE("sipgpu_array_create");
    if (rank < 1 || rank > kMaxRank || !nseg || !seg_ext || !out || world < 1 || my_rank < 0 || my_rank >= world)
        return SIPGPU_E_ARG;
    SIP_TRY(ensure_init());
    sipgpu_array* a = new sipgpu_array();
    a->rank = rank;
    a->my_rank = my_rank;
    a->world = world;
    const int* e = seg_ext;
    long long nb = 1;
    for (int i = 0; i < rank; ++i) {
        if (nseg[i] < 1) { delete a; return SIPGPU_E_ARG; }
        a->nseg[i] = nseg[i];
        a->seg_ext[i].assign(e, e + nseg[i]);
        e += nseg[i];
        nb *= nseg[i];
    }
    long long s = 1;
    for (int p = rank - 1; p >= 0; --p) { a->slice[p] = s; s *= nseg[p]; }  // array_table.cpp:50-73
    a->nblocks = nb;
    a->block_off.resize(nb);
    a->slab_elems.assign(world, 0);
    for (long long b = 0; b < nb; ++b) {
        long long rem = b, sz = 1;
        for (int i = 0; i < rank; ++i) {  // num2id, array_table.cpp:83-97
            const long long q = rem / a->slice[i];
            sz *= a->seg_ext[i][q];
            rem -= q * a->slice[i];
        }
        const int owner = (int)(b % world);
        a->block_off[b] = a->slab_elems[owner];
        a->slab_elems[owner] += (sz + kBlockAlign - 1) / kBlockAlign * kBlockAlign;
    }
    a->base.assign(world, nullptr);
    a->opened.assign(world, 0);
    const size_t bytes = sizeof(double) * (size_t)(a->slab_elems[my_rank] > 0 ? a->slab_elems[my_rank] : kBlockAlign);
    void* p = nullptr;
    // a dedicated cudaMalloc (not the pool): the IPC handle must name exactly this allocation
    cudaError_t ce = cudaMalloc(&p, bytes);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        set_error("distributed array slab: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(ce));
        delete a;
        return SIPGPU_E_NOMEM;
    }
    a->base[my_rank] = (double*)p;
    ++mem_epoch();
    SIP_CUDA(cudaMemsetAsync(p, 0, bytes, ctx().stream));  // new server blocks are zero (disk_backed_block_map.cpp:150-183)
    // With peers the zero fill must have landed before the slab can be exported: a peer's first put / put += may arrive
    // while this rank's stream is still busy with earlier work and would otherwise be wiped by the late memset (the
    // reference needs no barrier between create and the first put either).
    if (world > 1) SIP_CUDA(cudaStreamSynchronize(ctx().stream));
    *out = a;
    return SIPGPU_OK;
}

int sipgpu_array_destroy(sipgpu_array* a) {
    SIP_TRACE("sipgpu_array_destroy");
    if (!a) return SIPGPU_OK;
    cudaStreamSynchronize(ctx().stream);
    for (int r = 0; r < a->world; ++r) {
        if (!a->base[r]) continue;
        if (a->opened[r]) cudaIpcCloseMemHandle(a->base[r]);
        else if (r == a->my_rank) { cudaFree(a->base[r]); ++mem_epoch(); }
    }
    delete a;
    return SIPGPU_OK;
}

int sipgpu_array_export(sipgpu_array* a, void* handle_bytes) {
    SIP_TRACE("sipgpu_array_export");
    if (!a || !handle_bytes) return SIPGPU_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == SIPGPU_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    SIP_CUDA(cudaIpcGetMemHandle(&h, a->base[a->my_rank]));
    memcpy(handle_bytes, &h, sizeof(h));
    return SIPGPU_OK;
}

int sipgpu_array_attach(sipgpu_array* a, int peer_rank, const void* handle_bytes, int peer_device) {
    SIP_TRACE("sipgpu_array_attach");
    if (!a || peer_rank < 0 || peer_rank >= a->world || !handle_bytes) return SIPGPU_E_ARG;
    if (peer_rank == a->my_rank) return SIPGPU_OK;
    (void)peer_device;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_bytes, sizeof(h));
    void* p = nullptr;
    SIP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    a->base[peer_rank] = (double*)p;
    a->opened[peer_rank] = 1;
    return SIPGPU_OK;
}

long long sipgpu_array_block_number(const sipgpu_array* a, const int* idx) {
    if (check_idx(a, idx) != SIPGPU_OK) return -1;
    long long b = 0;
    for (int i = 0; i < a->rank; ++i) b += a->slice[i] * (idx[i] - 1);  // array_table.cpp:75-81 (lower = 1)
    return b;
}
int sipgpu_array_block_owner(const sipgpu_array* a, long long b) {
    if (!a || b < 0 || b >= a->nblocks) return -1;
    return (int)(b % a->world);  // data_distribution.cpp:74-82
}
long long sipgpu_array_block_size(const sipgpu_array* a, const int* idx) {
    if (check_idx(a, idx) != SIPGPU_OK) return -1;
    long long sz = 1;
    for (int i = 0; i < a->rank; ++i) sz *= a->seg_ext[i][idx[i] - 1];
    return sz;
}
double* sipgpu_array_block_ptr(sipgpu_array* a, const int* idx) {
    const long long b = sipgpu_array_block_number(a, idx);
    if (b < 0) return nullptr;
    double* base = a->base[b % a->world];
    if (!base) {
        set_error("distributed array: slab of rank %d is not attached", (int)(b % a->world));
        return nullptr;
    }
    return base + a->block_off[b];
}

int sipgpu_array_get(sipgpu_array* a, const int* idx, double* g_dst) {
    SIP_TRACE("sipgpu_array_get");
    double* src = sipgpu_array_block_ptr(a, idx);
    if (!src || !g_dst) return src ? SIPGPU_E_ARG : SIPGPU_E_STATE;
    if (a->track) a->touched[sipgpu_array_block_number(a, idx)] |= SIPGPU_ACCESS_GET;
    const long long n = sipgpu_array_block_size(a, idx);
    if (wl_active()) return wl_rec_ew(WL_SCALE_COPY, g_dst, src, nullptr, n, 1.0);  // peer loads inside the batched kernel
    // peer read over NVLink when the owner is remote (UVA resolves the direction)
    SIP_CUDA(cudaMemcpyAsync(g_dst, src, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx().stream));
    return SIPGPU_OK;
}
int sipgpu_array_put(sipgpu_array* a, const int* idx, const double* g_src) {
    SIP_TRACE("sipgpu_array_put");
    double* dst = sipgpu_array_block_ptr(a, idx);
    if (!dst || !g_src) return dst ? SIPGPU_E_ARG : SIPGPU_E_STATE;
    if (a->track) a->touched[sipgpu_array_block_number(a, idx)] |= SIPGPU_ACCESS_PUT;
    const long long n = sipgpu_array_block_size(a, idx);
    if (wl_active()) return wl_rec_ew(WL_SCALE_COPY, dst, g_src, nullptr, n, 1.0);
    SIP_CUDA(cudaMemcpyAsync(dst, g_src, sizeof(double) * (size_t)n, cudaMemcpyDefault, ctx().stream));
    return SIPGPU_OK;
}
int sipgpu_array_put_accumulate(sipgpu_array* a, const int* idx, const double* g_src) {
    SIP_TRACE("sipgpu_array_put_accumulate");
    double* dst = sipgpu_array_block_ptr(a, idx);
    if (!dst || !g_src) return dst ? SIPGPU_E_ARG : SIPGPU_E_STATE;
    if (a->track) a->touched[sipgpu_array_block_number(a, idx)] |= SIPGPU_ACCESS_PUT_ACCUMULATE;
    if (wl_active()) return wl_rec_ew(WL_REDADD, dst, g_src, nullptr, sipgpu_array_block_size(a, idx), 1.0, a->world == 1);
    return ew_red_add(dst, g_src, sipgpu_array_block_size(a, idx));
}
// One barrier section's worth of get / put += in ONE call and ONE launch: the n blocks become n descriptors of the batched
// elementwise kernel (16-byte peer loads for get, red.global.add.f64 into the owner's slab for put +=).  The reference
// issues these one MPI message per block (sial_ops_parallel.cpp:132-171, 332-408); a pardo body that prefetches its section
// (or a section's worth of accumulates at its end) maps onto this.  Inside an open recording the ops just join it.
static int array_many(sipgpu_array* a, int n, const int* idx, double* const* blk, bool get) {
    if (!a || n < 0 || (n > 0 && (!idx || !blk))) return SIPGPU_E_ARG;
    if (n == 0) return SIPGPU_OK;
    const bool own = !wl_active();
    if (own) SIP_TRY(sipgpu_wl_begin(0));
    int rc = SIPGPU_OK;
    for (int i = 0; i < n && rc == SIPGPU_OK; ++i) {
        const int* ix = idx + (size_t)i * a->rank;
        rc = get ? sipgpu_array_get(a, ix, blk[i]) : sipgpu_array_put_accumulate(a, ix, blk[i]);
    }
    if (own) {
        const int rc2 = sipgpu_wl_end();
        if (rc == SIPGPU_OK) rc = rc2;
    }
    return rc;
}
int sipgpu_array_get_many(sipgpu_array* a, int n, const int* idx, double* const* g_dst) {
    SIP_TRACE("sipgpu_array_get_many");
    return array_many(a, n, idx, g_dst, true);
}
int sipgpu_array_put_accumulate_many(sipgpu_array* a, int n, const int* idx, const double* const* g_src) {
    SIP_TRACE("sipgpu_array_put_accumulate_many");
    return array_many(a, n, idx, const_cast<double* const*>(g_src), false);
}
// put_initialize / put_increment / put_scale (sial_ops_parallel.cpp:412-528): one scalar applied to one block at its
// owner.  The reference sends {value, block id} to the server, which runs a serial loop; here the elementwise kernel
// runs on the caller's stream directly on the owner's (possibly peer-mapped) block.  For the race detector they count
// as put, put_accumulate and put respectively (distributed_block_consistency.cpp:60).
int sipgpu_array_put_initialize(sipgpu_array* a, const int* idx, double value) {
    SIP_TRACE("sipgpu_array_put_initialize");
    double* dst = sipgpu_array_block_ptr(a, idx);
    if (!dst) return SIPGPU_E_STATE;
    if (a->track) a->touched[sipgpu_array_block_number(a, idx)] |= SIPGPU_ACCESS_PUT;
    const long long n = sipgpu_array_block_size(a, idx);
    return wl_active() ? wl_rec_ew(WL_FILL, dst, nullptr, nullptr, n, value) : ew_fill(dst, n, value);
}
int sipgpu_array_put_increment(sipgpu_array* a, const int* idx, double delta) {
    SIP_TRACE("sipgpu_array_put_increment");
    double* dst = sipgpu_array_block_ptr(a, idx);
    if (!dst) return SIPGPU_E_STATE;
    if (a->track) a->touched[sipgpu_array_block_number(a, idx)] |= SIPGPU_ACCESS_PUT_ACCUMULATE;
    const long long n = sipgpu_array_block_size(a, idx);
    // With several ranks put_increment shares a block with other workers' put_increment / put += inside one barrier
    // section (it counts as PUT_ACCUMULATE for the race rules, distributed_block_consistency.cpp:60); the reference
    // serialises them at the server.  A plain read-modify-write would lose updates, so the increment is atomic too.
    if (a->world > 1)
        return wl_active() ? wl_rec_ew(WL_REDINCR, dst, nullptr, nullptr, n, delta) : ew_red_increment(dst, n, delta);
    return wl_active() ? wl_rec_ew(WL_INCR, dst, nullptr, nullptr, n, delta) : ew_increment(dst, n, delta);
}
int sipgpu_array_put_scale(sipgpu_array* a, const int* idx, double factor) {
    SIP_TRACE("sipgpu_array_put_scale");
    double* dst = sipgpu_array_block_ptr(a, idx);
    if (!dst) return SIPGPU_E_STATE;
    if (a->track) a->touched[sipgpu_array_block_number(a, idx)] |= SIPGPU_ACCESS_PUT;
    const long long n = sipgpu_array_block_size(a, idx);
    return wl_active() ? wl_rec_ew(WL_SCALE, dst, nullptr, nullptr, n, factor) : ew_scale(dst, n, factor);
}
int sipgpu_array_fill_local(sipgpu_array* a, double v) {
    SIP_TRACE("sipgpu_array_fill_local");
    if (!a) return SIPGPU_E_ARG;
    if (wl_active()) return wl_rec_ew(WL_FILL, a->base[a->my_rank], nullptr, nullptr, a->slab_elems[a->my_rank], v);
    return ew_fill(a->base[a->my_rank], a->slab_elems[a->my_rank], v);
}
// ---- race detection between barriers (distributed_block_consistency.cpp:25-175) ----
// The reference checks every get / put / put += at the block's server against a state table.  Read as a whole, the
// table accepts the accesses of one barrier section to one block iff  (a) a single worker made all of them, or
// (b) all of them are GETs, or (c) all of them are PUT_ACCUMULATEs  -- the order inside a section does not matter.
// Without servers the check is made at the barrier instead: every rank hands over what it touched
// (sipgpu_array_section_accesses), the summaries are exchanged by the caller's barrier, and each rank validates the
// union with sipgpu_consistency_validate (host-only arithmetic).
int sipgpu_array_track_accesses(sipgpu_array* a, int on) {
    SIP_TRACE("sipgpu_array_track_accesses");
    if (!a) return SIPGPU_E_ARG;
    a->track = on != 0;
    a->touched.clear();
    return SIPGPU_OK;
}
long long sipgpu_array_section_accesses(sipgpu_array* a, long long cap, long long* block_numbers, int* access_bits) {
    SIP_TRACE("sipgpu_array_section_accesses");
    if (!a) return -1;
    std::map<long long, int> sorted(a->touched.begin(), a->touched.end());
    long long k = 0;
    for (const auto& kv : sorted) {
        if (k < cap && block_numbers && access_bits) { block_numbers[k] = kv.first; access_bits[k] = kv.second; }
        ++k;
    }
    return k;
}
int sipgpu_array_section_reset(sipgpu_array* a) {
    SIP_TRACE("sipgpu_array_section_reset");
    if (!a) return SIPGPU_E_ARG;
    a->touched.clear();  // sip_barrier: DistributedBlockConsistency::reset_consistency_status
    return SIPGPU_OK;
}
int sipgpu_consistency_validate(long long n, const long long* block_numbers, const int* access_bits, const int* workers,
                                long long* bad_block) {
    SIP_TRACE("sipgpu_consistency_validate");
    if (n < 0 || (n && (!block_numbers || !access_bits || !workers))) return SIPGPU_E_ARG;
    struct Agg { int bits = 0, worker = -1; bool multiple = false; };
    std::map<long long, Agg> agg;
    for (long long i = 0; i < n; ++i) {
        if (access_bits[i] == 0) continue;
        Agg& g = agg[block_numbers[i]];
        g.bits |= access_bits[i];
        if (g.worker >= 0 && g.worker != workers[i]) g.multiple = true;
        g.worker = workers[i];
    }
    for (const auto& kv : agg) {
        const Agg& g = kv.second;
        if (g.multiple && g.bits != SIPGPU_ACCESS_GET && g.bits != SIPGPU_ACCESS_PUT_ACCUMULATE) {
            if (bad_block) *bad_block = kv.first;
            set_error("inconsistent block %lld: accessed by several workers between two barriers with access bits %d "
                      "(only all-GET or all-PUT_ACCUMULATE may be shared)", kv.first, g.bits);
            return SIPGPU_E_STATE;
        }
    }
    return SIPGPU_OK;
}

// ---- persistence of the rank's slab (array_file.h:53-70 structure; one file pair per rank instead of MPI-IO) ----
// data file  : <int chunk_size = slab elements><int num_servers = world><double>*          (ArrayFile header_val_t = int)
// index file : <offset DENSE_INDEX = 77><offset nblocks><offset per block number>*        (offset_val_t = long long)
//              offset = byte position of the block in THIS rank's data file, -1 (ABSENT_BLOCK_OFFSET) for blocks of peers
int sipgpu_array_save(sipgpu_array* a, const char* data_path, const char* index_path) {
    SIP_TRACE("sipgpu_array_save");
    if (!a || !data_path || !index_path) return SIPGPU_E_ARG;
    SIP_TRY(wl_flush());
    SIP_TRY(ensure_init());
    const long long n = a->slab_elems[a->my_rank];
    if (n > 0x7fffffffLL) {
        set_error("array_save: slab of %lld elements exceeds the int header of the file format; save per block range", n);
        return SIPGPU_E_ARG;
    }
    FILE* f = fopen(data_path, "wb");
    if (!f) { set_error("array_save: cannot open %s", data_path); return SIPGPU_E_ARG; }
    const int header[2] = {(int)n, a->world};
    bool ok = fwrite(header, sizeof(int), 2, f) == 2;
    void* h = nullptr;
    const size_t stage = (size_t)32 << 20;
    if (cudaMallocHost(&h, stage) != cudaSuccess) { fclose(f); return cuda_fail(cudaGetLastError(), "cudaMallocHost", __FILE__, __LINE__); }
    int rc = SIPGPU_OK;
    for (long long off = 0; off < n && ok; off += (long long)(stage / 8)) {
        const size_t cnt = (size_t)std::min<long long>(n - off, (long long)(stage / 8));
        if (cudaMemcpyAsync(h, a->base[a->my_rank] + off, cnt * 8, cudaMemcpyDeviceToHost, ctx().stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx().stream) != cudaSuccess) {
            rc = cuda_fail(cudaGetLastError(), "array_save d2h", __FILE__, __LINE__);
            break;
        }
        ok = fwrite(h, 8, cnt, f) == cnt;
    }
    cudaFreeHost(h);
    fclose(f);
    if (rc != SIPGPU_OK) return rc;
    FILE* fi = fopen(index_path, "wb");
    if (!fi) { set_error("array_save: cannot open %s", index_path); return SIPGPU_E_ARG; }
    std::vector<long long> index(2 + (size_t)a->nblocks, -1);
    index[0] = 77;
    index[1] = a->nblocks;
    for (long long b = 0; b < a->nblocks; ++b)
        if (b % a->world == a->my_rank) index[2 + b] = (long long)(2 * sizeof(int)) + 8 * a->block_off[b];
    ok = ok && fwrite(index.data(), sizeof(long long), index.size(), fi) == index.size();
    fclose(fi);
    if (!ok) { set_error("array_save: short write"); return SIPGPU_E_ARG; }
    return SIPGPU_OK;
}
int sipgpu_array_load(sipgpu_array* a, const char* data_path, const char* index_path) {
    SIP_TRACE("sipgpu_array_load");
    if (!a || !data_path || !index_path) return SIPGPU_E_ARG;
    SIP_TRY(wl_flush());
    SIP_TRY(ensure_init());
    FILE* fi = fopen(index_path, "rb");
    if (!fi) { set_error("array_load: cannot open %s", index_path); return SIPGPU_E_ARG; }
    std::vector<long long> index(2 + (size_t)a->nblocks);
    const bool iok = fread(index.data(), sizeof(long long), index.size(), fi) == index.size();
    fclose(fi);
    if (!iok || index[0] != 77 || index[1] != a->nblocks) {
        set_error("array_load: index of %s does not describe this array (dense index of %lld blocks expected)", index_path, a->nblocks);
        return SIPGPU_E_ARG;
    }
    for (long long b = 0; b < a->nblocks; ++b) {
        const long long want = (b % a->world == a->my_rank) ? (long long)(2 * sizeof(int)) + 8 * a->block_off[b] : -1;
        if (index[2 + b] != want) {
            set_error("array_load: block %lld is at offset %lld in the file, this layout needs %lld (different segments or world size)",
                      b, index[2 + b], want);
            return SIPGPU_E_ARG;
        }
    }
    FILE* f = fopen(data_path, "rb");
    if (!f) { set_error("array_load: cannot open %s", data_path); return SIPGPU_E_ARG; }
    int header[2] = {0, 0};
    const long long n = a->slab_elems[a->my_rank];
    if (fread(header, sizeof(int), 2, f) != 2 || header[0] != (int)n || header[1] != a->world) {  // array_file.cpp:183-190
        fclose(f);
        set_error("array_load: header of %s (chunk_size %d, servers %d) does not match this array (%lld, %d)", data_path, header[0],
                  header[1], n, a->world);
        return SIPGPU_E_ARG;
    }
    void* h = nullptr;
    const size_t stage = (size_t)32 << 20;
    if (cudaMallocHost(&h, stage) != cudaSuccess) { fclose(f); return cuda_fail(cudaGetLastError(), "cudaMallocHost", __FILE__, __LINE__); }
    int rc = SIPGPU_OK;
    for (long long off = 0; off < n; off += (long long)(stage / 8)) {
        const size_t cnt = (size_t)std::min<long long>(n - off, (long long)(stage / 8));
        if (fread(h, 8, cnt, f) != cnt) { set_error("array_load: %s is truncated", data_path); rc = SIPGPU_E_ARG; break; }
        if (cudaMemcpyAsync(a->base[a->my_rank] + off, h, cnt * 8, cudaMemcpyHostToDevice, ctx().stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx().stream) != cudaSuccess) {
            rc = cuda_fail(cudaGetLastError(), "array_load h2d", __FILE__, __LINE__);
            break;
        }
    }
    cudaFreeHost(h);
    fclose(f);
    return rc;
}

double* sipgpu_array_local_base(sipgpu_array* a) { return a ? a->base[a->my_rank] : nullptr; }
size_t sipgpu_array_local_bytes(const sipgpu_array* a) {
    return a ? sizeof(double) * (size_t)a->slab_elems[a->my_rank] : 0;
}

}  // extern "C"

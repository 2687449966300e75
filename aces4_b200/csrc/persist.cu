// persist.cu -- persistence of device-resident arrays.
//
// Reference behaviour followed (not code):
//  * WorkerPersistentArrayManager::set_persistent / save_marked_arrays / restore_persistent
//    (src/sip/dynamic_data/worker_persistent_array_manager.cpp:34-155): at the end of a SIAL program the arrays marked
//    persistent are moved into label-keyed maps -- scalars by value, contiguous (static) arrays and distributed arrays by
//    OWNERSHIP TRANSFER, no copy -- and the next SIAL program takes them back by label (the entry is erased on restore;
//    re-using a label overwrites with a warning).  Here the objects stay resident in HBM across SIAL programs.
//  * checkpoint_persistent / init_from_checkpoint (:157-260): scalars and contiguous arrays are written to one file with
//    the little-endian stream conventions of setup/io_utils.cpp:44-70,114-155 --
//        int32 nscalars, nscalars x (string label, float64 value),
//        int32 narrays,  narrays  x (string label, int32 rank (= MAX_RANK), int32 count + rank x int32 dims,
//                                     int32 count + count x float64 data)
//    where a string is int32 length (INCLUDING the trailing NUL) + bytes.  The same bytes are written / read here, so a
//    checkpoint produced by either side restores on the other.
//  * server-side array files (src/sip/mpi/array_file.h:53-70): data file = <int chunk_size><int num_servers><double>*,
//    index file = <index_type><n><offset>* with DENSE_INDEX = 77 (one offset per block number, absent = -1).  The
//    reference writes one file collectively with MPI-IO; without MPI every rank writes its own slab with the same
//    header/index structure (chunk_size = slab elements, one chunk per rank).
#include <map>
#include <string>
#include <vector>

#include "elementwise.h"
#include "worklist.h"

struct sipgpu_array;
extern "C" {
long long sipgpu_array_block_number(const sipgpu_array* a, const int* idx);
}

namespace sipgpu {
namespace {

struct Contig {
    double* dev = nullptr;
    int rank = 0;
    int ext[kMaxRank] = {1, 1, 1, 1, 1, 1};
    long long n = 0;
};
std::map<std::string, double> g_scalars;
std::map<std::string, Contig> g_contig;
std::map<std::string, sipgpu_array*> g_arrays;

struct Writer {
    FILE* f;
    bool ok = true;
    void raw(const void* p, size_t n) { ok = ok && fwrite(p, 1, n, f) == n; }
    void i32(int v) { raw(&v, 4); }
    void f64(double v) { raw(&v, 8); }
    void str(const std::string& s) {  // io_utils.cpp:44-52: trailing blanks trimmed, length counts the NUL
        size_t end = s.find_last_not_of(' ');
        const std::string t = end == std::string::npos ? "" : s.substr(0, end + 1);
        i32((int)t.size() + 1);
        raw(t.c_str(), t.size() + 1);
    }
};
struct Reader {
    FILE* f;
    bool ok = true;
    void raw(void* p, size_t n) { ok = ok && fread(p, 1, n, f) == n; }
    int i32() { int v = 0; raw(&v, 4); return v; }
    double f64() { double v = 0; raw(&v, 8); return v; }
    std::string str() {
        const int len = i32();
        if (!ok || len < 0 || len > (1 << 20)) { ok = false; return ""; }
        std::vector<char> b((size_t)len + 1, 0);
        raw(b.data(), (size_t)len);
        return std::string(b.data());
    }
};

// device <-> file through a pinned staging buffer
constexpr size_t kStage = (size_t)32 << 20;
int dev_to_file(Writer& w, const double* dev, long long n) {
    void* h = nullptr;
    SIP_CUDA(cudaMallocHost(&h, kStage));
    int rc = SIPGPU_OK;
    for (long long off = 0; off < n && rc == SIPGPU_OK; off += (long long)(kStage / 8)) {
        const size_t cnt = (size_t)std::min<long long>(n - off, (long long)(kStage / 8));
        if (cudaMemcpyAsync(h, dev + off, cnt * 8, cudaMemcpyDeviceToHost, ctx().stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx().stream) != cudaSuccess) {
            rc = cuda_fail(cudaGetLastError(), "persist d2h", __FILE__, __LINE__);
            break;
        }
        w.raw(h, cnt * 8);
    }
    cudaFreeHost(h);
    return rc;
}
int file_to_dev(Reader& r, double* dev, long long n) {
    void* h = nullptr;
    SIP_CUDA(cudaMallocHost(&h, kStage));
    int rc = SIPGPU_OK;
    for (long long off = 0; off < n && rc == SIPGPU_OK; off += (long long)(kStage / 8)) {
        const size_t cnt = (size_t)std::min<long long>(n - off, (long long)(kStage / 8));
        r.raw(h, cnt * 8);
        if (!r.ok) { rc = SIPGPU_E_ARG; break; }
        if (cudaMemcpyAsync(dev + off, h, cnt * 8, cudaMemcpyHostToDevice, ctx().stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx().stream) != cudaSuccess) {
            rc = cuda_fail(cudaGetLastError(), "persist h2d", __FILE__, __LINE__);
            break;
        }
    }
    cudaFreeHost(h);
    return rc;
}

}  // namespace
}  // namespace sipgpu

using namespace sipgpu;

extern "C" {

// ---- label registry (hand-off between SIAL programs) ----
int sipgpu_persist_scalar(const char* label, double value) {
    SIP_TRACE("sipgpu_persist_scalar");
    if (!label) return SIPGPU_E_ARG;
    g_scalars[label] = value;  // a repeated label overwrites (worker_persistent_array_manager.cpp:124-130)
    return SIPGPU_OK;
}
int sipgpu_restore_scalar(const char* label, double* value) {
    SIP_TRACE("sipgpu_restore_scalar");
    if (!label || !value) return SIPGPU_E_ARG;
    auto it = g_scalars.find(label);
    if (it == g_scalars.end()) {
        set_error("restore_persistent: scalar with label %s not found", label);
        return SIPGPU_E_STATE;
    }
    *value = it->second;
    g_scalars.erase(it);
    return SIPGPU_OK;
}
int sipgpu_persist_contiguous(const char* label, double* dev_block, int rank, const int* ext) {
    SIP_TRACE("sipgpu_persist_contiguous");
    if (!label || !dev_block || rank < 0 || rank > kMaxRank || (rank && !ext)) return SIPGPU_E_ARG;
    SIP_TRY(wl_flush());
    Contig c;
    c.dev = dev_block;
    c.rank = rank;
    c.n = 1;
    for (int i = 0; i < rank; ++i) {
        if (ext[i] < 1) return SIPGPU_E_ARG;
        c.ext[i] = ext[i];
        c.n *= ext[i];
    }
    auto it = g_contig.find(label);
    if (it != g_contig.end()) {  // "Overwriting previously saved array": the old block goes back to the pool
        if (it->second.dev != dev_block) pool_free(it->second.dev);
        g_contig.erase(it);
    }
    g_contig[label] = c;
    return SIPGPU_OK;
}
int sipgpu_restore_contiguous(const char* label, double** dev_block, int* rank, int* ext6) {
    SIP_TRACE("sipgpu_restore_contiguous");
    if (!label || !dev_block) return SIPGPU_E_ARG;
    auto it = g_contig.find(label);
    if (it == g_contig.end()) {
        set_error("restore_persistent: contiguous array with label %s not found", label);
        return SIPGPU_E_STATE;
    }
    *dev_block = it->second.dev;
    if (rank) *rank = it->second.rank;
    if (ext6) memcpy(ext6, it->second.ext, sizeof(int) * kMaxRank);
    g_contig.erase(it);
    return SIPGPU_OK;
}
int sipgpu_persist_array(const char* label, sipgpu_array* a) {
    SIP_TRACE("sipgpu_persist_array");
    if (!label || !a) return SIPGPU_E_ARG;
    SIP_TRY(wl_flush());
    auto it = g_arrays.find(label);
    if (it != g_arrays.end() && it->second != a) sipgpu_array_destroy(it->second);
    g_arrays[label] = a;
    return SIPGPU_OK;
}
int sipgpu_restore_array(const char* label, sipgpu_array** a) {
    SIP_TRACE("sipgpu_restore_array");
    if (!label || !a) return SIPGPU_E_ARG;
    auto it = g_arrays.find(label);
    if (it == g_arrays.end()) {
        set_error("restore_persistent: distributed/served array with label %s not found", label);
        return SIPGPU_E_STATE;
    }
    *a = it->second;
    g_arrays.erase(it);
    return SIPGPU_OK;
}
int sipgpu_persist_count(int* nscalars, int* ncontiguous, int* narrays) {
    SIP_TRACE("sipgpu_persist_count");
    if (nscalars) *nscalars = (int)g_scalars.size();
    if (ncontiguous) *ncontiguous = (int)g_contig.size();
    if (narrays) *narrays = (int)g_arrays.size();
    return SIPGPU_OK;
}

// ---- checkpoint file of the scalars + contiguous arrays, reference byte format ----
int sipgpu_persist_checkpoint(const char* filename) {
    SIP_TRACE("sipgpu_persist_checkpoint");
    if (!filename) return SIPGPU_E_ARG;
    if (!g_contig.empty()) SIP_TRY(ensure_init());
    SIP_TRY(wl_flush());
    FILE* f = fopen(filename, "wb");
    if (!f) {
        set_error("checkpoint: cannot open %s for writing", filename);
        return SIPGPU_E_ARG;
    }
    Writer w{f};
    w.i32((int)g_scalars.size());
    for (const auto& kv : g_scalars) {
        w.str(kv.first);
        w.f64(kv.second);
    }
    w.i32((int)g_contig.size());
    int rc = SIPGPU_OK;
    for (const auto& kv : g_contig) {
        const Contig& c = kv.second;
        if (c.n > 0x7fffffffLL) { rc = SIPGPU_E_ARG; set_error("checkpoint: array %s exceeds the int32 element count of the format", kv.first.c_str()); break; }
        w.str(kv.first);
        w.i32(kMaxRank);  // the reference always writes MAX_RANK dims, trailing 1s
        w.i32(kMaxRank);
        w.raw(c.ext, sizeof(int) * kMaxRank);
        w.i32((int)c.n);
        rc = dev_to_file(w, c.dev, c.n);
        if (rc != SIPGPU_OK) break;
    }
    const bool ok = w.ok;
    fclose(f);
    if (rc == SIPGPU_OK && !ok) {
        set_error("checkpoint: short write to %s", filename);
        rc = SIPGPU_E_ARG;
    }
    return rc;
}
int sipgpu_persist_init_from_checkpoint(const char* filename) {
    SIP_TRACE("sipgpu_persist_init_from_checkpoint");
    if (!filename) return SIPGPU_E_ARG;
    if (!g_scalars.empty() || !g_contig.empty()) {  // CHECKs of init_from_checkpoint (:204-205)
        set_error("init_from_checkpoint: persistent maps are not empty");
        return SIPGPU_E_STATE;
    }
    FILE* f = fopen(filename, "rb");
    if (!f) {
        set_error("init_from_checkpoint: cannot open %s", filename);
        return SIPGPU_E_ARG;
    }
    Reader r{f};
    int rc = SIPGPU_OK;
    const int ns = r.i32();
    for (int i = 0; i < ns && r.ok; ++i) {
        const std::string name = r.str();
        const double v = r.f64();
        if (r.ok) g_scalars[name] = v;
    }
    const int na = r.ok ? r.i32() : 0;
    for (int i = 0; i < na && r.ok && rc == SIPGPU_OK; ++i) {
        const std::string name = r.str();
        int rank = r.i32();
        const int cnt = r.i32();
        if (!r.ok || cnt < 0 || cnt > kMaxRank) { r.ok = false; break; }
        Contig c;
        rank = cnt;
        for (int d = 0; d < cnt; ++d) c.ext[d] = r.i32();
        c.rank = rank;
        const int n = r.i32();
        long long vol = 1;
        for (int d = 0; d < cnt; ++d) vol *= c.ext[d];
        if (!r.ok || n < 0 || vol != n) { r.ok = false; break; }
        c.n = n;
        rc = ensure_init();
        if (rc != SIPGPU_OK) break;
        c.dev = pool_alloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        if (!c.dev) { rc = SIPGPU_E_NOMEM; break; }
        rc = file_to_dev(r, c.dev, n);
        if (rc != SIPGPU_OK) { pool_free(c.dev); break; }
        g_contig[name] = c;
    }
    fclose(f);
    if (rc == SIPGPU_OK && !r.ok) {
        set_error("init_from_checkpoint: malformed input %s", filename);
        rc = SIPGPU_E_ARG;
    }
    return rc;
}

}  // extern "C"

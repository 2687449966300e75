// sial_ops_device.hpp -- SialOpsDevice: the third implementation of the SialOps method set, next to
// SialOpsSequential and SialOpsParallel (src/sip/worker/sial_ops_parallel.h; selected at interpreter.h:300-304).
//
// Header-only C++ over the C ABI of libsipgpu (sipgpu.h).  It mirrors the public method set of SialOpsParallel
// (sial_ops_parallel.h:47-73; sial_ops_parallel.cpp: get :132-171, put_replace :232-284, put_accumulate :332-408,
// put_initialize/increment/scale :412-528, sip_barrier :39-99, collective_sum :549-565, create/delete_distributed,
// destroy_served / request / prequest / prepare / prepare_accumulate :531-547, assert_same :570-591, broadcast_static
// :611-617, set_persistent / restore_persistent :629-692, end_program :694-719) with the same argument meaning, but
// without servers: every worker owns the blocks whose number is congruent to its rank and the others reach them over
// NVLink.  Block identity is (array id, segment numbers) exactly as in sip::BlockId (block_id.h:40-250); the array's
// segment table comes from the index table (index_table.cpp:27-70), passed in once at create time.
//
// The process-level plumbing (exchange of IPC handles, barrier, all-reduce of one double) is injected as three
// callbacks, so the class binds to MPI (the reference's SIPMPIAttr communicator), NCCL or torch.distributed alike.
#pragma once
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "sipgpu.h"

namespace sipgpu {

class SialOpsDevice {
public:
    struct Comm {
        int rank = 0, size = 1;
        // all-gather of `bytes` opaque bytes per rank into out (size * bytes), rank order
        std::function<void(const void* mine, void* out, size_t bytes)> allgather;
        std::function<void()> barrier;
        std::function<double(double)> allreduce_sum;
        // broadcast of n doubles in HOST memory from rank `root` (MPI_Bcast); only assert_same / broadcast_static use it
        std::function<void(double* host, long long n, int root)> bcast;
    };

    explicit SialOpsDevice(Comm comm, bool check_races = false) : comm_(std::move(comm)), check_races_(check_races) {
        int dev = -1, rank = comm_.rank;
        check(_init_gpu(&dev, &rank), "_init_gpu");
    }
    ~SialOpsDevice() {
        for (auto& kv : arrays_) sipgpu_array_destroy(kv.second);
    }
    SialOpsDevice(const SialOpsDevice&) = delete;
    SialOpsDevice& operator=(const SialOpsDevice&) = delete;

    // create_distributed / served array: seg_ext[d] = extents of the segments of dimension d
    void create_distributed(int array_id, const std::vector<std::vector<int>>& seg_ext) {
        std::vector<int> nseg, flat;
        for (const auto& d : seg_ext) {
            nseg.push_back((int)d.size());
            flat.insert(flat.end(), d.begin(), d.end());
        }
        sipgpu_array* a = nullptr;
        check(sipgpu_array_create((int)seg_ext.size(), nseg.data(), flat.data(), comm_.rank, comm_.size, &a), "sipgpu_array_create");
        if (comm_.size > 1) {
            std::vector<unsigned char> mine(SIPGPU_IPC_HANDLE_BYTES), all((size_t)SIPGPU_IPC_HANDLE_BYTES * comm_.size);
            check(sipgpu_array_export(a, mine.data()), "sipgpu_array_export");
            comm_.allgather(mine.data(), all.data(), SIPGPU_IPC_HANDLE_BYTES);
            for (int r = 0; r < comm_.size; ++r)
                if (r != comm_.rank) check(sipgpu_array_attach(a, r, all.data() + (size_t)r * SIPGPU_IPC_HANDLE_BYTES, r), "sipgpu_array_attach");
        }
        if (check_races_) check(sipgpu_array_track_accesses(a, 1), "sipgpu_array_track_accesses");
        arrays_[array_id] = a;
    }
    void delete_distributed(int array_id) {
        auto it = arrays_.find(array_id);
        if (it == arrays_.end()) return;
        sipgpu_array_destroy(it->second);
        arrays_.erase(it);
        drop_cache(array_id);
    }

    // get: device address of the block, valid until the next sip_barrier.  A block this rank owns is used in place; a
    // remote block is fetched once per barrier section into the worker-side cache (sial_ops_parallel.cpp:41-47).
    const double* get(int array_id, const int* idx) {
        sipgpu_array* a = array(array_id);
        const long long num = sipgpu_array_block_number(a, idx);
        if (num < 0) throw std::runtime_error("SialOpsDevice::get: bad block id");
        if (sipgpu_array_block_owner(a, num) == comm_.rank && !check_races_) return sipgpu_array_block_ptr(a, idx);
        auto key = std::make_pair(array_id, num);
        auto it = cache_.find(key);
        if (it != cache_.end()) return it->second;
        double* tmp = sipgpu_block_alloc(sipgpu_array_block_size(a, idx), 0);
        if (!tmp) fail("sipgpu_block_alloc");
        check(sipgpu_array_get(a, idx, tmp), "sipgpu_array_get");
        cache_[key] = tmp;
        return tmp;
    }
    void put_replace(int array_id, const int* idx, const double* src) { check(sipgpu_array_put(array(array_id), idx, src), "sipgpu_array_put"); }
    void put_accumulate(int array_id, const int* idx, const double* src) {
        check(sipgpu_array_put_accumulate(array(array_id), idx, src), "sipgpu_array_put_accumulate");
    }
    void put_initialize(int array_id, const int* idx, double v) { check(sipgpu_array_put_initialize(array(array_id), idx, v), "put_initialize"); }
    void put_increment(int array_id, const int* idx, double v) { check(sipgpu_array_put_increment(array(array_id), idx, v), "put_increment"); }
    void put_scale(int array_id, const int* idx, double v) { check(sipgpu_array_put_scale(array(array_id), idx, v), "put_scale"); }

    // Served arrays are distributed arrays here (no server ranks): the served-array methods are the same aliases the
    // reference has (sial_ops_parallel.cpp:531-547); prequest is unsupported there as well (:538-540).
    void destroy_served(int array_id) { delete_distributed(array_id); }
    const double* request(int array_id, const int* idx) { return get(array_id, idx); }
    void prequest(int, const int*, const int*) { throw std::runtime_error("PREQUEST Not supported !"); }
    void prepare(int array_id, const int* idx, const double* src) { put_replace(array_id, idx, src); }
    void prepare_accumulate(int array_id, const int* idx, const double* src) { put_accumulate(array_id, idx, src); }

    // set_persistent / restore_persistent of a distributed or served array (:629-692 + the server side,
    // worker_persistent_array_manager.cpp:34-155): the resident array moves into / out of the library's label registry;
    // nothing is copied and the slabs stay mapped on every rank.  restore_distributed (:51) is the same hand-over.
    void set_persistent(int array_id, const std::string& label) {
        auto it = arrays_.find(array_id);
        if (it == arrays_.end()) throw std::runtime_error("SialOpsDevice::set_persistent: unknown array id " + std::to_string(array_id));
        drop_cache(array_id);
        check(sipgpu_persist_array(label.c_str(), it->second), "sipgpu_persist_array");
        arrays_.erase(it);
    }
    void restore_persistent(int array_id, const std::string& label) {
        if (arrays_.count(array_id)) delete_distributed(array_id);
        sipgpu_array* a = nullptr;
        check(sipgpu_restore_array(label.c_str(), &a), "sipgpu_restore_array");
        if (check_races_) check(sipgpu_array_track_accesses(a, 1), "sipgpu_array_track_accesses");
        arrays_[array_id] = a;
    }
    void restore_distributed(int array_id, const std::string& label) { restore_persistent(array_id, label); }
    void set_persistent_scalar(const std::string& label, double value) { check(sipgpu_persist_scalar(label.c_str(), value), "sipgpu_persist_scalar"); }
    double restore_persistent_scalar(const std::string& label) {
        double v = 0.0;
        check(sipgpu_restore_scalar(label.c_str(), &v), "sipgpu_restore_scalar");
        return v;
    }

    // assert_same (:570-591): every worker takes rank 0's value; a worker whose own value is not nearlyEqual (:593-609,
    // relative 5e-5) to it fails.  Returns the agreed value.
    double assert_same(double mine) {
        if (comm_.size <= 1 || !comm_.bcast) return mine;
        double v = mine;
        comm_.bcast(&v, 1, 0);
        if (comm_.rank != 0 && !nearly_equal(v, mine, .00005)) throw std::runtime_error("values are too far apart");
        return v;
    }
    // broadcast_static (:611-617): a static (contiguous, replicated) array resident on every rank's device takes the
    // contents of `source_worker`'s copy.  Staged through host memory: static arrays are small (coefficient / Fock
    // matrices) and this is not on the contraction path.
    void broadcast_static(double* dev_block, long long n, int source_worker) {
        if (comm_.size <= 1 || !comm_.bcast || n <= 0) return;
        std::vector<double> host((size_t)n);
        if (comm_.rank == source_worker) check(sipgpu_d2h(host.data(), dev_block, n), "sipgpu_d2h");
        comm_.bcast(host.data(), n, source_worker);
        if (comm_.rank != source_worker) {
            check(sipgpu_h2d(dev_block, host.data(), n), "sipgpu_h2d");
            check(sipgpu_sync(), "sipgpu_sync");   // `host` dies at the end of this scope
        }
    }
    // end_program (:694-719): the implicit final barrier; there are no servers to notify
    void end_program() { sip_barrier(); }

    // pardo entry / exit: the per-block calls in between are recorded and launched batched (sipgpu.h boundary 3c)
    void begin_pardo() { check(sipgpu_wl_begin(0), "sipgpu_wl_begin"); }
    void end_pardo() { check(sipgpu_wl_end(), "sipgpu_wl_end"); }

    // sip_barrier (:39-99): finish the device work of this section, drop the cached remote blocks, optionally run the
    // race detector of distributed_block_consistency.cpp over everybody's accesses, then the workers' barrier.
    void sip_barrier() {
        if (sipgpu_wl_recording()) check(sipgpu_wl_flush(), "sipgpu_wl_flush");
        check(sipgpu_sync(), "sipgpu_sync");
        for (auto& kv : cache_) sipgpu_block_free(kv.second);
        cache_.clear();
        if (check_races_) validate_section();
        if (comm_.barrier) comm_.barrier();
    }
    double collective_sum(double partial) { return comm_.allreduce_sum ? comm_.allreduce_sum(partial) : partial; }

    sipgpu_array* array(int array_id) {
        auto it = arrays_.find(array_id);
        if (it == arrays_.end()) throw std::runtime_error("SialOpsDevice: unknown array id " + std::to_string(array_id));
        return it->second;
    }

private:
    void validate_section() {
        for (auto& kv : arrays_) {
            const long long n = sipgpu_array_section_accesses(kv.second, 0, nullptr, nullptr);
            // fixed-size exchange: count first, then the padded lists
            long long mine_n = n;
            std::vector<long long> counts(comm_.size, n);
            if (comm_.size > 1) comm_.allgather(&mine_n, counts.data(), sizeof(long long));
            long long maxn = 0;
            for (long long c : counts) maxn = c > maxn ? c : maxn;
            std::vector<long long> blocks((size_t)maxn + 1, -1), all_blocks;
            std::vector<int> bits((size_t)maxn + 1, 0), all_bits;
            sipgpu_array_section_accesses(kv.second, n, blocks.data(), bits.data());
            all_blocks.resize((size_t)(maxn + 1) * comm_.size);
            all_bits.resize((size_t)(maxn + 1) * comm_.size);
            if (comm_.size > 1) {
                comm_.allgather(blocks.data(), all_blocks.data(), sizeof(long long) * (size_t)(maxn + 1));
                comm_.allgather(bits.data(), all_bits.data(), sizeof(int) * (size_t)(maxn + 1));
            } else {
                all_blocks = blocks;
                all_bits = bits;
            }
            std::vector<int> workers(all_blocks.size());
            for (size_t i = 0; i < workers.size(); ++i) workers[i] = (int)(i / (size_t)(maxn + 1));
            long long bad = -1;
            check(sipgpu_consistency_validate((long long)all_blocks.size(), all_blocks.data(), all_bits.data(), workers.data(), &bad),
                  "race detected (distributed_block_consistency)");
            sipgpu_array_section_reset(kv.second);
        }
    }
    void drop_cache(int array_id) {
        for (auto it = cache_.begin(); it != cache_.end();)
            if (it->first.first == array_id) { sipgpu_block_free(it->second); it = cache_.erase(it); } else ++it;
    }
    static bool nearly_equal(double a, double b, double epsilon) {   // SialOpsParallel::nearlyEqual, :593-609
        const double abs_a = a < 0 ? -a : a, abs_b = b < 0 ? -b : b, diff = a - b < 0 ? b - a : a - b;
        if (a == b) return true;
        if (a == 0 || b == 0 || diff < 2.2250738585072014e-308) return diff < epsilon * 2.2250738585072014e-308;
        return diff / (abs_a + abs_b) < epsilon;
    }
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string(what) + ": " + sipgpu_last_error()); }
    void check(int rc, const char* what) {
        if (rc != SIPGPU_OK) fail(what);
    }

    Comm comm_;
    bool check_races_;
    std::map<int, sipgpu_array*> arrays_;
    std::map<std::pair<int, long long>, double*> cache_;
};

}  // namespace sipgpu

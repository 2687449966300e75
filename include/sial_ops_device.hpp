// sial_ops_device.hpp -- SialOpsDevice: the third implementation of the SialOps method set, next to
// SialOpsSequential and SialOpsParallel (src/sip/worker/sial_ops_parallel.h; selected at interpreter.h:300-304).
//
// Header-only C++ over the C ABI of libsipgpu (sipgpu.h).  It mirrors the methods of SialOpsParallel that move blocks
// (sial_ops_parallel.cpp: get :132-171, put_replace :232-284, put_accumulate :332-408, put_initialize/increment/scale
// :412-528, sip_barrier :39-99, collective_sum :549-565, create/delete_distributed) with the same argument meaning, but
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

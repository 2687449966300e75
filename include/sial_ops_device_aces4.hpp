// sial_ops_device_aces4.hpp -- INTEGRATION.md level 2, the file a maintainer drops next to src/sip/worker/sial_ops_parallel.h:
// SialOpsDevice behind the REFERENCE'S OWN method signatures (sial_ops_parallel.h:47-73) -- sip::BlockId& for the block,
// sip::Block::BlockPtr for the data, `int pc` for the program counter the wait timers use -- so that the interpreter's
//      #if defined(HAVE_MPI) ... SialOpsParallel sial_ops_;       (interpreter.h:300-304)
// can become `SialOpsDeviceAces4 sial_ops_;` under HAVE_CUDA with no change to any call site
// (interpreter.cpp:611-655: sial_ops_.get(id, pc), .put_replace(id, block, pc), .put_accumulate(id, block, pc), ...).
//
// It compiles against the reference's headers (block_id.h, block.h built with HAVE_CUDA: the device half of sip::Block of
// INTEGRATION level 1 holds the block's data, Block::get_gpu_data()) and libsipgpu's C ABI; nothing else of Aces4 is needed,
// which is why `make -C oracle ref_l2` can build and the tests can run it without the interpreter (no MPI, no .siox).
// What the reference's SialOpsParallel takes from SipTables -- the segment extents of an array's dimensions
// (sip_tables.h:96-140 / index_table.cpp:27-70) -- arrives through one callback given at construction.
#pragma once
#include <functional>
#include <vector>

#include "block.h"
#include "block_id.h"
#include "sial_ops_device.hpp"

namespace sip {

class SialOpsDeviceAces4 {
public:
    // segments_of(array_id) = extents of the segments of every dimension of the array (from the index table)
    using SegmentTable = std::function<std::vector<std::vector<int>>(int array_id)>;
    // where collective_sum leaves its result: the interpreter's scalar table (data_manager_.set_scalar_value)
    using ScalarSink = std::function<void(int slot, double value)>;

    SialOpsDeviceAces4(sipgpu::SialOpsDevice::Comm comm, SegmentTable segments_of, ScalarSink set_scalar, bool check_races = false)
        : ops_(std::move(comm), check_races), segments_of_(std::move(segments_of)), set_scalar_(std::move(set_scalar)) {}

    void sip_barrier(int /*pc*/) { ops_.sip_barrier(); }

    void create_distributed(int array_id, int /*pc*/) {
        const std::vector<std::vector<int>> segs = segments_of_(array_id);
        rank_[array_id] = (int)segs.size();
        ops_.create_distributed(array_id, segs);
    }
    void delete_distributed(int array_id, int /*pc*/) { ops_.delete_distributed(array_id); }
    void destroy_served(int array_id, int pc) { delete_distributed(array_id, pc); }

    // get / request: the block becomes readable on the device until the next barrier (sial_ops_parallel.cpp:132-171);
    // the interpreter then asks for it with get_block_for_reading
    void get(BlockId& id, int /*pc*/) { ops_.get(id.array_id(), idx(id)); }
    void request(BlockId& id, int pc) { get(id, pc); }
    const double* get_block_for_reading(const BlockId& id, int /*pc*/) { return ops_.get(id.array_id(), idx(id)); }

    void put_replace(BlockId& id, const Block::BlockPtr block, int /*pc*/) { ops_.put_replace(id.array_id(), idx(id), device(block)); }
    void put_accumulate(BlockId& id, const Block::BlockPtr block, int /*pc*/) { ops_.put_accumulate(id.array_id(), idx(id), device(block)); }
    void prepare(BlockId& id, Block::BlockPtr block, int pc) { put_replace(id, block, pc); }
    void prepare_accumulate(BlockId& id, Block::BlockPtr block, int pc) { put_accumulate(id, block, pc); }
    void put_initialize(BlockId& id, double value, int /*pc*/) { ops_.put_initialize(id.array_id(), idx(id), value); }
    void put_increment(BlockId& id, double value, int /*pc*/) { ops_.put_increment(id.array_id(), idx(id), value); }
    void put_scale(BlockId& id, double value, int /*pc*/) { ops_.put_scale(id.array_id(), idx(id), value); }

    void collective_sum(double rhs_value, int dest_array_slot) { set_scalar_(dest_array_slot, ops_.collective_sum(rhs_value)); }
    void end_program() { ops_.end_program(); }

    // pardo entry / exit: the deferred op stream of INTEGRATION level 3
    void begin_pardo() { ops_.begin_pardo(); }
    void end_pardo() { ops_.end_pardo(); }

    sipgpu::SialOpsDevice& device_ops() { return ops_; }

private:
    // BlockId -> the 1-based segment numbers of the array's dimensions (block_id.h:40-250: index_values_[MAX_RANK], unused = -1)
    const int* idx(const BlockId& id) {
        const int r = rank_.at(id.array_id());
        for (int i = 0; i < r; ++i) scratch_[i] = id.index_values(i);
        return scratch_;
    }
    // the device half of a sip::Block (block.h:135-193 under HAVE_CUDA)
    static const double* device(const Block::BlockPtr block) {
        double* p = block->get_gpu_data();
        if (!p) throw std::runtime_error("SialOpsDeviceAces4: block has no device data (lazy_gpu_write_on_device first)");
        return p;
    }

    sipgpu::SialOpsDevice ops_;
    SegmentTable segments_of_;
    ScalarSink set_scalar_;
    std::map<int, int> rank_;
    int scratch_[MAX_RANK];
};

}  // namespace sip

// oracle/ref_shim/level2/aces4_ref_l2_shim.cpp -- TEST INFRASTRUCTURE: INTEGRATION.md level 2, executed.
//
// include/sial_ops_device_aces4.hpp (SialOpsDevice behind the reference's SialOpsParallel signatures) compiled against the
// REFERENCE'S headers and objects -- sip::BlockId (block_id.cpp), sip::Block with its device half (block.cpp built with
// HAVE_CUDA on the level-1 replacement header) -- and driven the way interpreter.cpp:611-655 drives `sial_ops_`:
// BlockId for the block, Block::BlockPtr for the data, a pc.  One C entry point runs the closed forms of the reference's
// Sial tests (test/test_sial.cpp: put_test :282-318, get :583, put_accumulate_stress :1072-1113, put_initialize /
// increment / scale) and reports what it read back.
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "sial_ops_device_aces4.hpp"

static std::string g_err;

extern "C" {

const char* aces4ref_l2_last_error() { return g_err.c_str(); }

// nseg segments of `seg` elements in both dimensions of a rank-2 distributed array (id 7).
// out[0] = number of blocks checked, out[1] = number of mismatching elements, out[2] = collective_sum result seen in the
// scalar sink, out[3] = value read back after put_initialize / increment / scale.  Returns 0, or -1 with aces4ref_l2_last_error().
int aces4ref_l2_selftest(int nseg, int seg, int reps, double* out) {
    try {
        using namespace sip;
        double sink_value = 0.0;
        int sink_slot = -1;
        sipgpu::SialOpsDevice::Comm comm;   // one worker: no callbacks
        SialOpsDeviceAces4 ops(
            comm, [&](int) { return std::vector<std::vector<int>>(2, std::vector<int>(nseg, seg)); },
            [&](int slot, double v) { sink_slot = slot; sink_value = v; }, /*check_races=*/true);
        const int array_id = 7, pc = 0;
        ops.create_distributed(array_id, pc);
        const BlockShape shape = [&] { segment_size_array_t e; for (int i = 0; i < MAX_RANK; ++i) e[i] = 1; e[0] = seg; e[1] = seg; return BlockShape(e, 2); }();
        auto id_of = [&](int i, int j) { std::vector<int> v = {i, j}; return BlockId(array_id, 2, v); };
        // put_accumulate_stress: pardo k: put c[i,j] += a; += aa; += a; += aa   with a = i, aa = j   (recorded as one pardo)
        ops.begin_pardo();
        std::vector<Block::BlockPtr> keep;
        for (int k = 0; k < reps; ++k)
            for (int i = 1; i <= nseg; ++i)
                for (int j = 1; j <= nseg; ++j) {
                    Block::BlockPtr a = Block::new_gpu_block(shape), aa = Block::new_gpu_block(shape);
                    a->gpu_fill((double)i);
                    aa->gpu_fill((double)j);
                    BlockId id = id_of(i, j);
                    ops.put_accumulate(id, a, pc);
                    ops.put_accumulate(id, aa, pc);
                    ops.put_accumulate(id, a, pc);
                    ops.put_accumulate(id, aa, pc);
                    keep.push_back(a);
                    keep.push_back(aa);
                }
        ops.end_pardo();
        for (Block::BlockPtr b : keep) delete b;
        ops.sip_barrier(pc);
        long long blocks = 0, bad = 0;
        std::vector<double> h((size_t)seg * seg);
        for (int i = 1; i <= nseg; ++i)
            for (int j = 1; j <= nseg; ++j) {
                BlockId id = id_of(i, j);
                ops.get(id, pc);
                if (sipgpu_d2h(h.data(), ops.get_block_for_reading(id, pc), (long long)h.size()) != 0) throw std::runtime_error(sipgpu_last_error());
                for (double x : h) bad += x != reps * (2.0 * i + 2.0 * j);
                ++blocks;
            }
        ops.sip_barrier(pc);
        // put_replace of a block, then the scalar block ops at the owner
        BlockId id22 = id_of(nseg, nseg);
        Block::BlockPtr b = Block::new_gpu_block(shape);
        b->gpu_fill(1.0);
        ops.put_replace(id22, b, pc);
        ops.sip_barrier(pc);
        ops.put_increment(id22, 0.5, pc);
        ops.sip_barrier(pc);
        ops.put_scale(id22, 4.0, pc);
        ops.sip_barrier(pc);
        if (sipgpu_d2h(h.data(), ops.get_block_for_reading(id22, pc), (long long)h.size()) != 0) throw std::runtime_error(sipgpu_last_error());
        double rb = h[0];
        for (double x : h) bad += x != rb;
        ops.sip_barrier(pc);
        BlockId id11 = id_of(1, 1);
        ops.put_initialize(id11, -2.5, pc);
        ops.sip_barrier(pc);
        if (sipgpu_d2h(h.data(), ops.get_block_for_reading(id11, pc), (long long)h.size()) != 0) throw std::runtime_error(sipgpu_last_error());
        for (double x : h) bad += x != -2.5;
        ops.collective_sum(1.25, 3);
        delete b;
        ops.sip_barrier(pc);
        ops.delete_distributed(array_id, pc);
        out[0] = (double)blocks;
        out[1] = (double)bad;
        out[2] = sink_slot == 3 ? sink_value : std::nan("");
        out[3] = rb;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"

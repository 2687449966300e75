// SialOpsDevice (include/sial_ops_device.hpp) on one rank: the reference's Sial.put_test / put_accumulate_stress /
// get closed forms (test/test_sial.cpp:282-318, 583, 1072-1113) driven through the class the way the interpreter's
// SialOps calls would, including a recorded pardo and the race detector.  Prints "ok" and returns 0 on success.
#include <cmath>
#include <cstdio>
#include <vector>

#include "sial_ops_device.hpp"

#define REQUIRE(c)                                                      \
    do {                                                                \
        if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

int main() {
    sipgpu::SialOpsDevice::Comm comm;  // single worker: no callbacks needed
    sipgpu::SialOpsDevice ops(comm, /*check_races=*/true);
    const std::vector<std::vector<int>> segs = {{2, 3, 2}, {2, 3, 2}};
    ops.create_distributed(7, segs);
    // put_accumulate_stress: pardo k(1..20): put c[i,j] += a; += aa; += a; += aa with a = i, aa = j
    ops.begin_pardo();
    for (int k = 0; k < 20; ++k)
        for (int i = 1; i <= 3; ++i)
            for (int j = 1; j <= 3; ++j) {
                const int idx[2] = {i, j};
                const long long n = sipgpu_array_block_size(ops.array(7), idx);
                double* a = sipgpu_block_alloc(n, 0);
                double* aa = sipgpu_block_alloc(n, 0);
                sipgpu_block_fill(a, n, (double)i);
                sipgpu_block_fill(aa, n, (double)j);
                ops.put_accumulate(7, idx, a);
                ops.put_accumulate(7, idx, aa);
                ops.put_accumulate(7, idx, a);
                ops.put_accumulate(7, idx, aa);
                sipgpu_block_free(a);
                sipgpu_block_free(aa);
            }
    ops.end_pardo();
    ops.sip_barrier();
    for (int i = 1; i <= 3; ++i)
        for (int j = 1; j <= 3; ++j) {
            const int idx[2] = {i, j};
            const long long n = sipgpu_array_block_size(ops.array(7), idx);
            std::vector<double> h((size_t)n);
            REQUIRE(sipgpu_d2h(h.data(), ops.get(7, idx), n) == 0);
            for (double x : h) REQUIRE(x == 20.0 * (2 * i + 2 * j));
        }
    ops.sip_barrier();
    // put_initialize / increment / scale, then collective_sum of a per-worker partial
    const int idx[2] = {2, 2};
    ops.put_initialize(7, idx, 1.0);
    ops.put_increment(7, idx, 0.5);
    ops.put_scale(7, idx, 4.0);
    ops.sip_barrier();
    double norm = 0;
    REQUIRE(sipgpu_block_norm2(ops.get(7, idx), 9, &norm) == 0);
    REQUIRE(std::fabs(norm - 9 * 36.0) < 1e-12);
    REQUIRE(ops.collective_sum(1.25) == 1.25);
    ops.sip_barrier();
    ops.delete_distributed(7);
    printf("ok\n");
    return 0;
}

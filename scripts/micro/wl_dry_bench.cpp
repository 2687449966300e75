// wl_dry_bench.cpp -- host cost of the deferred op stream (aces4_b200/csrc/worklist.cu), no GPU needed.
// The op-at-a-time call stream of two LCCD pardo bodies (rlccd_rhf.sialx: hhladder_ab :342-355, phladder_ab pardo 3
// :537-556; same loops as wl_pardo_bench.cpp) is recorded in DRY mode -- fake block addresses, nothing launched -- so
// that what is timed is exactly the host work per recorded op: the entry points' argument checks, the recording, and
// the scheduling passes A-C of the flush (emission needs a device and is not included).
//   build: g++ -O2 -std=c++17 wl_dry_bench.cpp -I../../include -L../../aces4_b200/lib -lsipgpu -Wl,-rpath,... -o wl_dry_bench
//   usage: wl_dry_bench [o_seg n_o_segs v_seg n_v_segs [reps]]      prints one JSON line
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "sipgpu.h"

#define CK(call)                                                                       \
    do {                                                                               \
        int rc__ = (call);                                                             \
        if (rc__ != 0) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, sipgpu_last_error()); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

static int so = 16, no = 3, sv = 16, nv = 6;

// a made-up resident block of array `arr`: distinct, non-overlapping address ranges (never dereferenced in dry mode)
static double* blk(int arr, int a, int b, int c, int d, long long n) {
    const long long num = ((a * 16 + b) * 16 + c) * 16 + d;
    return (double*)(((unsigned long long)(arr + 1) << 40) + (unsigned long long)num * (unsigned long long)n * 8ull);
}
static void ext4(int k0, int k1, int k2, int k3, int* e) {
    const int k[4] = {k0, k1, k2, k3};
    for (int d = 0; d < 4; ++d) e[d] = k[d] ? sv : so;
}
static long long vol(const int* e) { return (long long)e[0] * e[1] * e[2] * e[3]; }

static void hhladder() {
    int eD[4], eL[4], eR[4];
    ext4(1, 0, 1, 0, eD); ext4(1, 0, 1, 0, eL); ext4(0, 0, 0, 0, eR);
    const int dl[4] = {1, 2, 3, 4}, ll[4] = {1, 5, 3, 6}, rl[4] = {2, 5, 4, 6};
    for (int j1 = 1; j1 <= no; ++j1)
        for (int i1 = 1; i1 <= no; ++i1)
            for (int b = 1; b <= nv; ++b)
                for (int a = 1; a <= nv; ++a) {
                    const double* L = blk(0, a, i1, b, j1, vol(eL));
                    for (int i = 1; i <= no; ++i)
                        for (int j = 1; j <= no; ++j) {
                            double* T = sipgpu_block_alloc(vol(eD), 0);
                            CK(sipgpu_block_contract_labels(4, eD, dl, T, 4, eL, ll, L, 4, eR, rl, blk(1, i, i1, j, j1, vol(eR)), 1.0, 0.0));
                            CK(sipgpu_block_accumulate(blk(2, a, i, b, j, vol(eD)), T, vol(eD)));
                            CK(sipgpu_block_free(T));
                        }
                }
}

static void phring3() {
    int eD[4], eL[4], eR[4];
    ext4(1, 0, 1, 0, eD); ext4(1, 0, 1, 0, eL); ext4(1, 1, 0, 0, eR);
    const int dl[4] = {1, 2, 3, 4}, ll[4] = {1, 5, 8, 4}, rl[4] = {3, 8, 5, 2}, pl[4] = {3, 4, 1, 2};
    for (int b1 = 1; b1 <= nv; ++b1)
        for (int i1 = 1; i1 <= no; ++i1)
            for (int j = 1; j <= no; ++j)
                for (int a = 1; a <= nv; ++a) {
                    const double* L = blk(0, a, i1, b1, j, vol(eL));
                    for (int i = 1; i <= no; ++i)
                        for (int b = 1; b <= nv; ++b) {
                            double* T = sipgpu_block_alloc(vol(eD), 0);
                            CK(sipgpu_block_contract_labels(4, eD, dl, T, 4, eL, ll, L, 4, eR, rl, blk(3, b, b1, i1, i, vol(eR)), 1.0, 0.0));
                            CK(sipgpu_block_scale(T, vol(eD), -1.0));
                            double* T2 = sipgpu_block_alloc(vol(eD), 0);
                            CK(sipgpu_block_permute_labels(4, eD, pl, dl, T, T2));
                            CK(sipgpu_block_accumulate(blk(2, a, i, b, j, vol(eD)), T, vol(eD)));
                            CK(sipgpu_block_accumulate(blk(2, b, j, a, i, vol(eD)), T2, vol(eD)));
                            CK(sipgpu_block_free(T));
                            CK(sipgpu_block_free(T2));
                        }
                }
}

int main(int argc, char** argv) {
    if (argc > 4) { so = atoi(argv[1]); no = atoi(argv[2]); sv = atoi(argv[3]); nv = atoi(argv[4]); }
    const int reps = argc > 5 ? atoi(argv[5]) : 5;
    double best = 1e30;
    long long ops = 0, sched = 0, fused = 0, chains = 0;
    for (int r = 0; r < reps; ++r) {
        const auto t0 = std::chrono::steady_clock::now();
        long long st[9], n = 0, s = 0, f = 0, c = 0;
        void (*pardos[2])() = {hhladder, phring3};
        for (auto body : pardos) {
            CK(sipgpu_wl_begin(1));
            body();
            CK(sipgpu_wl_end());
            CK(sipgpu_wl_stats(st));
            n += st[0]; s += st[1]; f += st[4]; c += st[5];
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (sec < best) best = sec;
        ops = n; sched = s; fused = f; chains = c;
    }
    printf("{\"workload\": \"hhladder_ab + phladder_ab pardo 3, dry recording, o=%dx%d v=%dx%d\", \"ops_recorded\": %lld, "
           "\"ops_scheduled\": %lld, \"fused_accumulates\": %lld, \"chains\": %lld, \"seconds\": %.6f, \"us_per_op\": %.4f}\n",
           no, so, nv, sv, ops, sched, fused, chains, best, best / (double)ops * 1e6);
    return 0;
}

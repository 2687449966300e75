// dmma_ablate.cu -- progressive ablation of the contraction kernel's inner loop on sm_100a: where does the gap between
// the DMMA issue-rate peak and the kernel's pipe utilisation come from?  Development probe (not part of the library).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dmma_ablate dmma_ablate.cu && ./dmma_ablate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(double* s, const double* g) {
    unsigned d = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(g));
}

// MODE 0: registers only, same a/b for all 32 accumulators
// MODE 1: registers only, a[MF] x b[NF] distinct fragments (register-bank pattern of the real kernel)
// MODE 2: + fragments re-loaded from shared memory every k-step (LDS.64, conflict-free layout), no barrier
// MODE 3: + __syncthreads every 8 k-steps
// MODE 4: + cp.async 16B loads of the next stage (64 KB per stage) + wait_group + barrier  (= real main loop)
template <int MODE, int MF, int NF, int WARPS, int STAGES = 2, int LDIV = 2, bool HBM = false>
__global__ void __launch_bounds__(WARPS * 32, 1) k(int stages, const double* __restrict__ gsrc, double* sink) {
    extern __shared__ double sm[];
    constexpr int LDK = 36, BM = 128, BN = 128;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    constexpr int WM = BM / (MF * 8);
    const int wm = (warp % WM) * MF * 8, wn = (warp / WM) * NF * 8;
    for (int i = tid; i < STAGES * (BM + BN) * LDK; i += WARPS * 32) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double acc[MF][NF][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double a[2][MF], b[2][NF];
#pragma unroll
    for (int i = 0; i < MF; ++i) a[0][i] = a[1][i] = 1.0 + i * 1e-9 + tid * 1e-12;
#pragma unroll
    for (int j = 0; j < NF; ++j) b[0][j] = b[1][j] = 1.0 - j * 1e-9;
    const double* gp = gsrc + (size_t)blockIdx.x * 8192 + tid * 2;
    for (int s = 0; s < stages; ++s) {
        const double* as = sm + (s % STAGES) * (BM + BN) * LDK;
        const double* bs = as + BM * LDK;
        double* ls = sm + ((s + STAGES - 1) % STAGES) * (BM + BN) * LDK;
        if (MODE >= 4 && MODE != 5) { asm volatile("cp.async.wait_group %0;\n" ::"n"(STAGES - 2)); }
        if (MODE >= 3) __syncthreads();
        auto frags = [&](int kk, int buf) {
            if (MODE >= 2) {
#pragma unroll
                for (int mi = 0; mi < MF; ++mi) a[buf][mi] = as[(wm + mi * 8 + g) * LDK + kk * 4 + t4];
#pragma unroll
                for (int ni = 0; ni < NF; ++ni) b[buf][ni] = bs[(wn + ni * 8 + g) * LDK + kk * 4 + t4];
            }
        };
        frags(0, 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            if (kk + 1 < 8) frags(kk + 1, (kk + 1) & 1);
#pragma unroll
            for (int mi = 0; mi < MF; ++mi) {
#pragma unroll
                for (int ni = 0; ni < NF; ++ni) {
                    if (MODE == 0) dmma(acc[mi][ni][0], acc[mi][ni][1], a[0][0], b[0][0]);
                    else dmma(acc[mi][ni][0], acc[mi][ni][1], a[kk & 1][mi], b[kk & 1][ni]);
                }
                if (MODE >= 4) {
                    // 64 KB per stage / (WARPS*32 threads * 16 B) items, spread over the first half of the groups
                    constexpr int ITEMS = 65536 / (WARPS * 32 * 16);
                    constexpr int GROUPS = 8 * MF / LDIV;
                    const int grp = kk * MF + mi;
                    if (grp < GROUPS) {
#pragma unroll
                        for (int it = grp * ITEMS / GROUPS; it < (grp + 1) * ITEMS / GROUPS; ++it) {
                            const int e = (it * WARPS * 32 + tid) * 2;  // element index in the stage (128x32 A then B)
                            cp16(ls + (e / 32) * LDK + (e % 32), gp + (HBM ? (size_t)(s & 511) * 8192 * 148 : (size_t)(s & 63) * 1024 * 148) + it * 512);
                        }
                    }
                }
            }
        }
        if (MODE >= 4) asm volatile("cp.async.commit_group;\n" ::);
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) t += acc[i][j][0] + acc[i][j][1];
    if (t == 123.456) sink[0] = t;
}

template <int MODE, int MF, int NF, int WARPS, int STAGES = 2, int LDIV = 2, bool HBM = false>
void run(const char* name, const double* g, double* sink) {
    const int stages = 4000;
    const size_t smem = STAGES * (128 + 128) * 36 * 8;
    cudaFuncSetAttribute(k<MODE, MF, NF, WARPS, STAGES, LDIV, HBM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<MODE, MF, NF, WARPS, STAGES, LDIV, HBM><<<148, WARPS * 32, smem>>>(stages / 10, g, sink);
    cudaEventRecord(e0);
    k<MODE, MF, NF, WARPS, STAGES, LDIV, HBM><<<148, WARPS * 32, smem>>>(stages, g, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 148.0 * WARPS * stages * 8 * MF * NF * 512.0;
    printf("%-58s %7.2f TFLOP/s  (%s)\n", name, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double *g, *sink;
    cudaMalloc(&g, (size_t)5 << 30);
    cudaMemset(g, 0, (size_t)5 << 30);
    cudaMalloc(&sink, 8);
    run<3, 8, 4, 8>("3 LDS frags + barrier per 32 k, 8 warps", g, sink);
    run<4, 8, 4, 8, 2, 2>("4 + cp.async (L2 hits) 2 stages, loads in 1st half", g, sink);
    run<4, 8, 4, 8, 2, 1>("4 2 stages, loads spread over whole stage", g, sink);
    run<4, 8, 4, 8, 2, 4>("4 2 stages, loads in 1st quarter", g, sink);
    run<5, 8, 4, 8, 2, 2>("5 2 stages, loads issued, never waited (no wait_group)", g, sink);
    run<4, 8, 4, 8, 3, 2>("4 3 stages, loads in 1st half", g, sink);
    run<4, 8, 4, 8, 3, 1>("4 3 stages, loads spread", g, sink);
    run<4, 8, 4, 8, 2, 2, true>("4 2 stages, HBM stream (4.8 GB ring)", g, sink);
    run<4, 8, 4, 8, 3, 2, true>("4 3 stages, HBM stream", g, sink);
    run<4, 8, 4, 8, 3, 1, true>("4 3 stages, HBM stream, spread", g, sink);
    return 0;
}

// oracle/ref_shim/aces4_ref_gpu_shim.cpp -- TEST INFRASTRUCTURE.  Never linked into, imported by or shipped with the
// product (libsipgpu.so).
//
// extern "C" names for the REFERENCE's legacy CUDA backend, src/sip/cuda/gpu_super_instructions.cu (C++ linkage there;
// written for sm_13, switched off in the reference's CMake).  oracle/Makefile (target `ref_gpu`) compiles that file
// UNMODIFIED, where it lies under /root/reference, with nvcc for sm_100a and links it with this shim and cuBLAS into
// oracle/_ref/libaces4_ref_gpu.so.  It is the one implementation of this path's contraction and permutation that is
// reference code AND runs in this environment (the CPU path is Fortran): reorderScatter of both operands into scratch
// buffers, cublasDgemm, reorderGather into the destination (gpu_super_instructions.cu:369-586), three cudaMalloc of 40 MB
// and seven device synchronisations per call, so blocks are limited to 40 MB (5 242 880 doubles).
//
// Used by tests/test_gpu_vs_ref_cuda.py (product `_gpu_contract` / `_gpu_permute` == reference `_gpu_contract` /
// `_gpu_permute` on the same device buffers) and scripts/ref_gpu_baseline.py (the legacy path timed on the same B200).
// The reference code calls exit(EXIT_FAILURE) on any CUDA error (cuda_check.h:37-43): callers run it in a child process.
#include "config.h"
#define HAVE_CUDA 1
#include "gpu_super_instructions.h"

extern "C" {
int aces4ref_gpu_init(int rank) {
    int dev = -1;
    _init_gpu(&dev, &rank);
    return dev;
}
double* aces4ref_gpu_allocate(int n) { return _gpu_allocate(n); }
void aces4ref_gpu_free(double* p) { _gpu_free(p); }
void aces4ref_gpu_host_to_device(double* h, double* g, int n) { _gpu_host_to_device(h, g, n); }
void aces4ref_gpu_device_to_host(double* h, double* g, int n) { _gpu_device_to_host(h, g, n); }
void aces4ref_gpu_device_to_device(double* dst, double* src, int n) { _gpu_device_to_device(dst, src, n); }
void aces4ref_gpu_double_memset(double* g, double v, int n) { _gpu_double_memset(g, v, n); }
void aces4ref_gpu_selfmultiply(double* x, double alpha, int n) { _gpu_selfmultiply(x, alpha, n); }
void aces4ref_gpu_axpy(double* y, double* x, double alpha, int n) { _gpu_axpy(y, x, alpha, n); }
void aces4ref_gpu_permute(double* y, int ny, const int* yDims, const int* yInds, double* x, int nx, const int* xDims,
                          const int* xInds) {
    _gpu_permute(y, ny, yDims, yInds, x, nx, xDims, xInds);
}
void aces4ref_gpu_contract(double* y, int ny, const int* yDims, const int* yInds, double* x1, int n1, const int* x1Dims,
                           const int* x1Inds, double* x2, int n2, const int* x2Dims, const int* x2Inds) {
    _gpu_contract(y, ny, yDims, yInds, x1, n1, x1Dims, x1Inds, x2, n2, x2Dims, x2Inds);
}
}

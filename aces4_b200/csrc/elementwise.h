// elementwise.h -- launch interface of the HBM-bound block kernels (elementwise.cu).
#pragma once
#include "common.h"

namespace sipgpu {

int ew_fill(double* d, long long n, double v);
int ew_scale(double* d, long long n, double f);
int ew_scale_copy(double* d, const double* s, long long n, double f);
int ew_increment(double* d, long long n, double delta);
int ew_axpy(double* d, const double* s, long long n, double f);
int ew_add_sub(double* d, const double* l, const double* r, long long n, double sign);
int ew_dot(const double* a, const double* b, long long n, double* result_host);  // blocking
// d_out[0] = prev_scale*d_out[0] + sum a*b, asynchronous (device result)
int ew_dot_device(const double* a, const double* b, long long n, double* d_out, double prev_scale);
int ew_slice(int rank, const double* t, const int* t_ext, double* s, const int* s_ext, const int* beg);
int ew_insert(int rank, double* t, const int* t_ext, const double* s, const int* s_ext, const int* beg);
int ew_red_add(double* d, const double* s, long long n);
int ew_red_increment(double* d, long long n, double delta);  // d[i] += delta, atomically (many-writer put_increment)
int ew_fill_hash(double* d, long long n, unsigned long long seed, unsigned long long tag, double scale);
int ew_copy_probe(double* d, const double* s, long long n);
// worklist.cu: d_i = beta * d_i for n blocks in ONE launch (beta = 0: zero fill)
int ew_scale_many(int n, double* const* d, const long long* cnt, double beta);

// permute.cu
int permute_block(int rank, const int* ext, const int* transp, const double* in, double* out);
int permute_batched(int n, int rank, const int* ext, const int* transp, const double* const* in, double* const* out,
                    double alpha, double beta);
void permute_set_bulk(int variant);  // permute kernel: 0 register-staged tiles, 1 TMA bulk copies where the plan allows it, 2 cp.async ring, 3 chosen per launch (default)
void permute_set_vec(int on);        // ring variant: 16-byte cp.async / stores where the runs are even and aligned: 1 always, 0 never, -1 for big tiles (default)
int permute_plan_debug(int rank, const int* ext, const int* transp, long long* meta, int* rtab, int* wtab, int cap);

// superinstr.cu: elementwise CC super-instructions on device blocks
int si_set_int_array(const char* name, int n, const int* values);
int si_energy_denominator_rhf(int rank, const int* index_values, const int* ext, double* data, int fock_rank,
                              const int* fock_ext, const double* fock, const double* d_shift = nullptr);
int si_stripi(int rank, const int* iv0, const int* ext0, const double* x, const int* iv1, const int* ext1, double* y);
int si_anti_symm_o(int rank, const int* iv, const int* ext, double* x);
int si_anti_symm_v(int rank, const int* iv, const int* ext, double* x);
int si_return_sval(int rank, const int* ext, const double* data, double* d_scalar);
int si_invert_diagonal(int rank0, int rank1, const int* ext, double* a1, const double* a2);
int si_invert_diagonal_asym(int rank0, int rank1, const int* iv, const int* ext, double* a1, const double* a2);
int si_return_diagonal_elements(int rank, const int* ext, double* x);

}  // namespace sipgpu

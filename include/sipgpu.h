/*
 * sipgpu.h -- C ABI of libsipgpu.so: the B200 (sm_100a) block-tensor backend for the Aces4 SIP runtime.
 *
 * Every entry point is extern "C", takes plain pointers and sizes, and NEVER calls exit(): failures are
 * reported through `ierr` out-parameters (boundary 1) or int return codes (boundaries 2-4, 0 = success,
 * see SIPGPU_E_*).  There is no CPU fallback anywhere behind this header: a call made without a usable
 * CUDA device fails with SIPGPU_E_NODEVICE.
 *
 * Conventions (identical to the reference): FP64 elements, 32-bit extents, dense
 * column-major blocks (first index fastest), rank <= SIPGPU_MAX_RANK, segment numbers 1-based.
 *
 * Threading (the reference's contract for libtensordil, kept): the library is called from ONE thread per process --
 * the SIP interpreter is single-threaded and runs one process per worker / per GPU -- and keeps process-global state
 * (device, streams, block pool, the open recording).  It is not re-entrant; callers that add threads serialise the calls.
 *
 * Paths cited below are relative to the reference checkout (UFParLab/aces4).
 */
#ifndef SIPGPU_H_
#define SIPGPU_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIPGPU_MAX_RANK 6 /* src/sip/core/aces_defs.h:13 MAX_RANK */

enum {
    SIPGPU_OK = 0,
    SIPGPU_E_NODEVICE = 100, /* no CUDA device / driver: the product path has no CPU fallback */
    SIPGPU_E_CUDA = 101,     /* a CUDA runtime call failed; sipgpu_last_error() has the text */
    SIPGPU_E_ARG = 102,      /* bad rank / extents / null pointer */
    SIPGPU_E_PATTERN = 103,  /* illegal label pattern (label not exactly twice, trace, ...) */
    SIPGPU_E_NOMEM = 104,    /* device pool exhausted */
    SIPGPU_E_STATE = 105     /* library not initialised / wrong distributed-array section mode */
};

/* ---------------------------------------------------------------------------------------------
 * Boundary 1 -- the libtensordil ABI (HOST pointers).  Same names, argument order and by-reference
 * scalars as src/sip/tensor_algebra/tensor_ops_c_prototypes.h:41-178, so an unmodified aces4 links
 * this library in place of tensor_dil_omp.F90.  Operands are staged to the device, the sm_100a
 * kernels run, results are staged back; the call returns when the host buffers are valid.
 * `nthreads` is accepted and ignored.  ierr: 0 = success, otherwise the reference's own codes where it
 * defines them, else SIPGPU_E_*.
 * A translation unit of aces4 that already includes tensor_ops_c_prototypes.h (block.cpp, interpreter.cpp) has these ten
 * names declared with C++ reference parameters (`int&` -- the same ABI, another C++ type): it defines
 * SIPGPU_NO_TENSORDIL_PROTOTYPES before including this header (INTEGRATION.md level 1).
 * --------------------------------------------------------------------------------------------- */
#ifndef SIPGPU_NO_TENSORDIL_PROTOTYPES
long long tensor_size_by_shape_(int* num_dim, int* dims, int* ierr);           /* tensor_dil_omp.F90:65-85   */
void get_contraction_ptrn_(int* drank, int* lrank, int* rrank, int* aces_ptrn, /* tensor_dil_omp.F90:87-142  */
                           int* my_ptrn, int* ierr);
void tensor_block_init__(int* nthreads, double* tens, int* rank, int* ext, double* val, int* ierr);   /* F90:144-191 */
void tensor_block_scale__(int* nthreads, double* tens, int* rank, int* ext, double* fac, int* ierr);  /* F90:193-228 */
double tensor_block_norm2__(int* nthreads, double* tens, int* rank, int* ext, int* ierr);             /* F90:230-269 */
void tensor_block_slice__(int* nthreads, int* rank, double* tens, int* tens_ext, double* slice,       /* F90:271-330 */
                          int* slice_ext, int* ext_beg, int* ierr);
void tensor_block_insert__(int* nthreads, int* rank, double* tens, int* tens_ext, double* slice,      /* F90:332-392 */
                           int* slice_ext, int* ext_beg, int* ierr);
void tensor_block_add__(const int* nthreads, const int* rank, int* ext, double* tens0, double* tens1, /* F90:394-436 */
                        const double* fac, int* ierr);
void tensor_block_copy__(int* nthreads, int* rank, int* ext, int* dim_transp, double* tens_in,        /* F90:438-660 */
                         double* tens_out, int* ierr);
void tensor_block_contract__(int* nthreads, int* contr_ptrn, double* ltens, int* lrank, int* lext,    /* F90:662-796 */
                             double* rtens, int* rrank, int* rext, double* dtens, int* drank, int* dext,
                             int* ierr);
#endif /* SIPGPU_NO_TENSORDIL_PROTOTYPES */

/* ---------------------------------------------------------------------------------------------
 * Boundary 2 -- the device-block ABI (DEVICE pointers, resident blocks).  Same names and arguments as
 * src/sip/cuda/gpu_super_instructions.h:26-126 (+ _finalize_gpu, gpu_super_instructions.cu:64-70), but
 * extern "C", returning a status instead of void, backed by a slab pool instead of cudaMalloc per block,
 * and asynchronous on the library's compute stream (ordering between calls is preserved; only
 * _gpu_device_to_host blocks until the host buffer is valid).  _gpu_allocate returns NULL on failure.
 * --------------------------------------------------------------------------------------------- */
int _init_gpu(int* devid, int* my_rank);   /* device = *my_rank % device_count; *devid receives it */
int _finalize_gpu(void);
double* _gpu_allocate(const int num_elems);                               /* zero-filled, like the reference */
int _gpu_free(double* g_addr);
int _gpu_host_to_device(double* c_addr, double* g_addr, const int num_elems);
int _gpu_device_to_host(double* c_addr, double* g_addr, const int num_elems);
int _gpu_device_to_device(double* dst, double* src, const int num_elems);
int _gpu_double_memset(double* g_addr, double value, const int num_elems);
int _gpu_selfmultiply(double* x, const double alpha, const int num_elems); /* x *= alpha  */
int _gpu_axpy(double* y, double* x, const double alpha, const int num_elems); /* y += alpha*x */
int _gpu_permute(double* y, const int ny, const int* y_dims, const int* y_inds, /* y[y_inds] = x[x_inds] */
                 double* x, const int nx, const int* x_dims, const int* x_inds);
int _gpu_contract(double* y, const int ny, const int* y_dims, const int* y_inds, /* y = x1 * x2 (assign) */
                  double* x1, const int n1, const int* x1_dims, const int* x1_inds,
                  double* x2, const int n2, const int* x2_dims, const int* x2_inds);

/* ---------------------------------------------------------------------------------------------
 * Boundary 3 -- resident / batched extension (what a device-aware SialOps or super-instruction calls).
 * Blocks live in the device pool; ops are asynchronous on the compute stream.
 * --------------------------------------------------------------------------------------------- */
int sipgpu_init(int device);               /* idempotent; selects the device, creates streams + pool */
int sipgpu_finalize(void);
int sipgpu_device(void);                   /* current device ordinal or -1 */
const char* sipgpu_last_error(void);
int sipgpu_sync(void);                     /* wait for the compute stream */
void* sipgpu_stream(void);                 /* cudaStream_t of the compute stream (for event timing) */
long long sipgpu_kernel_launches(void);    /* number of sm_100a kernels this library has launched */

/* pool (replaces _gpu_allocate's cudaMalloc per block and Block's `new double[]`, block.cpp:28-79) */
int sipgpu_pool_reserve(size_t bytes);     /* pre-grow the arena */
double* sipgpu_block_alloc(long long num_elems, int zero);
int sipgpu_block_free(double* g_addr);
int sipgpu_pool_stats(size_t* bytes_reserved, size_t* bytes_in_use, size_t* n_live_blocks);
int sipgpu_h2d(double* g_dst, const double* h_src, long long n);   /* async when h_src is pinned */
int sipgpu_d2h(double* h_dst, const double* g_src, long long n);   /* blocking */
void* sipgpu_host_alloc(size_t bytes);     /* pinned host memory */
int sipgpu_host_free(void* p);

/* Host/device coherence of one block: the lazy_gpu_{read,write,update}_on_{device,host} transitions of
 * src/sip/dynamic_data/block_manager.cpp:340-441 over the status bits of sip::Block (block.h:198-204).  For callers
 * that do not carry the reference's Block class (Fortran super-instructions, other runtimes): the mirror owns the
 * device copy (pool block) and, if it was created without a host buffer, a pinned host copy. */
typedef struct sipgpu_mirror sipgpu_mirror;
enum { SIPGPU_ON_HOST = 1, SIPGPU_ON_GPU = 2, SIPGPU_DIRTY_ON_HOST = 4, SIPGPU_DIRTY_ON_GPU = 8 };
enum { SIPGPU_READ_ON_DEVICE = 0, SIPGPU_WRITE_ON_DEVICE, SIPGPU_UPDATE_ON_DEVICE,
       SIPGPU_READ_ON_HOST, SIPGPU_WRITE_ON_HOST, SIPGPU_UPDATE_ON_HOST };
/* pure state transition (host-only): action 0 none, 1 h2d, 2 d2h, 3 allocate device + h2d, 4 allocate host + d2h,
 * 5 fresh zeroed device block, 6 fresh zeroed host block; SIPGPU_E_STATE where the reference fails */
int sipgpu_mirror_transition(int status_bits, int op, int* new_status_bits, int* action);
int sipgpu_mirror_create(double* host_or_null, long long num_elems, sipgpu_mirror** out);
int sipgpu_mirror_destroy(sipgpu_mirror* m);
int sipgpu_mirror_status(const sipgpu_mirror* m);
double* sipgpu_mirror_access(sipgpu_mirror* m, int op);   /* pointer of the side that may now be used; NULL on error */
double* sipgpu_mirror_host_ptr(sipgpu_mirror* m);
double* sipgpu_mirror_device_ptr(sipgpu_mirror* m);

/* elementwise block ops of src/sip/dynamic_data/block.cpp:132-268 and interpreter.cpp:1874-1997 */
int sipgpu_block_fill(double* d, long long n, double v);                          /* block.cpp:153-176 */
int sipgpu_block_scale(double* d, long long n, double f);                         /* block.cpp:179-186 */
int sipgpu_block_scale_and_copy(double* d, const double* s, long long n, double f); /* block.cpp:189-202 */
int sipgpu_block_increment(double* d, long long n, double delta);                 /* block.cpp:206-213 */
int sipgpu_block_accumulate(double* d, const double* s, long long n);             /* block.cpp:259-268 */
int sipgpu_block_axpy(double* d, const double* s, long long n, double f);         /* F90:394-436      */
int sipgpu_block_add_sub(double* d, const double* l, const double* r, long long n, double sign); /* interpreter.cpp:1929,1992 */
/* synthetic inputs: d[i] = scale * uniform(-1,1) from splitmix64((seed ^ tag) + (i+1)*0x9E3779B97F4A7C15) */
int sipgpu_block_fill_hash(double* d, long long n, unsigned long long seed, unsigned long long tag, double scale);
int sipgpu_block_norm2(const double* t, long long n, double* result_host);
/* d_scalar[0] += sum l*r, asynchronous, result stays on the device (energy accumulation over many blocks) */
int sipgpu_block_dot_accumulate(const double* l, const double* r, long long n, double* d_scalar);        /* F90:230-269 (blocking) */
int sipgpu_block_dot(const double* l, const double* r, long long n, double* result_host); /* F90:910-938 (blocking) */
int sipgpu_block_slice(int rank, const double* t, const int* t_ext, double* s, const int* s_ext, const int* beg);
int sipgpu_block_insert(int rank, double* t, const int* t_ext, const double* s, const int* s_ext, const int* beg);

/* permute: out[new position of idx] = in[idx]; transp[0] = sign slot (ignored), transp[i] = NEW position
 * (1-based) of OLD dimension i -- the "dmitry_permute" vector of block.cpp:227-231 / F90:438-660. */
int sipgpu_block_permute(int rank, const int* ext, const int* transp, const double* in, double* out);
/* lhs[lhs_labels] = rhs[rhs_labels]  (block_permute_op, interpreter.cpp:656-679, 2049-2084) */
int sipgpu_block_permute_labels(int rank, const int* rhs_ext, const int* lhs_labels, const int* rhs_labels,
                                const double* rhs, double* lhs);

/* n blocks of identical extents and permutation in ONE launch: out_i = alpha * permuted(in_i) + beta * out_i.
 * alpha = 1, beta = 0 is n transposes; beta != 0 is the fused permute-accumulate that handle_block_add
 * (interpreter.cpp:1874-1997) performs as "permute into a temp block, then add" (e.g. the
 * `T2new[b,j,a,i] += R[a,i,b,j]` bodies of rlccd_rhf.sialx:558-603).  in/out are HOST arrays of n device pointers. */
int sipgpu_permute_batched(int n, int rank, const int* ext, const int* transp, const double* const* in,
                           double* const* out, double alpha, double beta);

/* contraction with the Lyakh pattern of get_contraction_ptrn_ (F90:662-796).  D = alpha*L*R + beta*D;
 * the reference op is alpha = 1, beta = 0 (assign, F90:752).  beta != 0 is the fused-accumulate
 * extension (contract followed by block_add / put +=). */
int sipgpu_block_contract(const int* ptrn, const double* L, int lrank, const int* lext, const double* R, int rrank,
                          const int* rext, double* D, int drank, const int* dext, double alpha, double beta);
/* D[dlab] = L[llab] * R[rlab] by labels (handle_contraction, interpreter.cpp:1210-1262) */
int sipgpu_block_contract_labels(int drank, const int* dext, const int* dlab, double* D, int lrank, const int* lext,
                                 const int* llab, const double* L, int rrank, const int* rext, const int* rlab,
                                 const double* R, double alpha, double beta);

/* Contraction whose operands may be SLICES of larger dense (static / contiguous) arrays, read and written in place:
 * X points to the parent array, Xparent_ext are its extents, Xbeg the 0-based first element of the block inside it
 * (both NULL: X is a dense block).  Replaces "extract_slice into a temp block, contract, free" for every use of a
 * static array block such as ca[mu,p] (contiguous_array_manager.cpp:162-230, block.cpp:272-323, F90:271-392). */
int sipgpu_block_contract_sliced(const int* ptrn, const double* L, int lrank, const int* lext, const int* lparent_ext,
                                 const int* lbeg, const double* R, int rrank, const int* rext, const int* rparent_ext,
                                 const int* rbeg, double* D, int drank, const int* dext, const int* dparent_ext,
                                 const int* dbeg, double alpha, double beta);

/* Batched contraction: n blocks in ONE launch per kernel variant.  All problems share rank/pattern
 * (`ptrn`), extents may differ per problem: lext/rext/dext are [n][rank] row-major arrays; L/R/D are
 * arrays of n device pointers; alpha/beta apply to all.  This is the pardo-body work-list entry point
 * (block-sparse batching of many small blocks per launch). */
int sipgpu_contract_batched(int n, const int* ptrn, int lrank, int rrank, int drank, const int* lext, const int* rext,
                            const int* dext, const double* const* L, const double* const* R, double* const* D,
                            double alpha, double beta);

/* Chained (block-sparse) contraction: destination i receives the SUM over the operand pairs
 * chain_start[i] .. chain_start[i+1]-1 of L[]/R[] (all pairs of a destination share its extents):
 *   D_i = alpha * sum_c L_c * R_c + beta * D_i.
 * This is a pardo body of the form `do k: T = A[..k..]*B[..k..]; put D += T` (e.g. rlccd_rhf.sialx:342-355,
 * 482-556) executed destination-stationary: the sum over contracted segments stays in the tile accumulators and
 * every destination block is written once. */
int sipgpu_contract_chained(int n, const int* ptrn, int lrank, int rrank, int drank, const int* lext, const int* rext,
                            const int* dext, const int* chain_start, const double* const* L, const double* const* R,
                            double* const* D, double alpha, double beta);

/* Prepared work-list: the arguments of sipgpu_contract_chained (chain_start may be NULL: n independent blocks) marshalled
 * ONCE -- shapes, problem and pair descriptors, tile prefix sums resident on the device -- and launched any number of
 * times.  A CC iteration replays the same work-lists (same blocks, same chains; rlccd_rhf.sialx:342-556 every
 * iteration), and for thousands of small blocks the per-call host marshalling is longer than the kernels.  The pointer
 * arrays are copied; the blocks they name must stay alive.  A launch with a new (alpha, beta) pair prepares that variant
 * on first use. */
typedef struct sipgpu_plan sipgpu_plan;
int sipgpu_plan_contract_chained(int n, const int* ptrn, int lrank, int rrank, int drank, const int* lext, const int* rext,
                                 const int* dext, const int* chain_start, const double* const* L, const double* const* R,
                                 double* const* D, sipgpu_plan** out);
int sipgpu_plan_launch(sipgpu_plan* plan, double alpha, double beta);
int sipgpu_plan_destroy(sipgpu_plan* plan);

/* Raw strided-batched GEMM view of the contraction core, for micro-benchmarks of the DMMA kernel:
 * C(m x n, col-major, ldc) = alpha * A^T * B + beta*C with A = [k x m] (lda), B = [k x n] (ldb): exactly the
 * dgemm('T','N',...) of F90:762. */
int sipgpu_dgemm_tn(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
                    double beta, double* C, int ldc);
/* register-resident DMMA issue-rate probe: returns achieved TFLOP/s of mma.sync.m8n8k4.f64 on all SMs */
int sipgpu_dmma_peak_probe(int iters, double* tflops_out);
/* device copy bandwidth probe (GB/s, read+write) over a buffer of `bytes` */
int sipgpu_copy_bw_probe(size_t bytes, int reps, double* gbs_out);
/* Per-entry-point trace: calls and host wall time of every C-ABI entry point since the last reset -- the table the
 * reference's Tracer keeps per interpreter opcode (src/sip/worker/tracer.h:41-50: pc_histogram_, opcode_timer_), so that
 * the two can be laid side by side.  Host time only: the device work of an entry point is asynchronous (device time per
 * kernel is what ncu / the CUDA-event brackets of bench.py measure).  Off by default.  sipgpu_trace_report fills up to `cap`
 * rows, most host time first (names are static strings), and returns the number of rows that have calls. */
int sipgpu_trace_enable(int on);
int sipgpu_trace_reset(void);
int sipgpu_trace_report(int cap, const char** names, long long* calls, double* host_seconds);
/* Launch-policy knobs for A/B measurements and for tests that must reach a particular kernel (defaults are the product
 * path).  "lowint_max_intensity": contractions with N <= 64 and at most this many flops per algorithmic byte run on the
 * bandwidth-shaped kernel (lowint.cu; default 7.0 = just above the roofline ridge of 5.7; negative: never).
 * "lowint_scope": 0 never, 1 (default) dot products + single-tile destinations where that kernel measured faster than the
 * 128-wide tiles, 2 every eligible shape (tests, A/B).  "lowint_slab": 1 (default) small results of long contractions whose
 * chunks are contiguous runs are fed by TMA bulk copies (slab_kernel), 0 never.  "permute_bulk": which permute kernel --
 * 0 register-staged, 1 TMA bulk runs, 2 cp.async ring, 3 (default) chosen per launch.  "permute_vec": -1 (default) / 0 / 1
 * 16-byte accesses in the ring kernel.  "copy_bulk": bit 0 whole-block copies (get / put), bit 1 put += travel as TMA bulk
 * transfers (default 3), 0 LDG.128 / red.global.add.f64; -1 back to the environment / default.  Returns SIPGPU_E_ARG for an
 * unknown key.  Host-only. */
int sipgpu_set_tuning(const char* key, double value);

/* host-only views of the planner (no device needed; used by the CPU tests of the host logic) */
int sipgpu_debug_contract_shape(const int* ptrn, int lrank, const int* lext, int rrank, const int* rext, int drank,
                                const int* dext, int* shape_ints);
int sipgpu_debug_shape_ints(void);
int sipgpu_debug_permute_plan(int rank, const int* ext, const int* transp, long long* meta, int* rtab, int* wtab,
                              int cap);

/* ---------------------------------------------------------------------------------------------
 * Boundary 3b -- elementwise CC super-instructions on DEVICE blocks, with the reference's super-instruction
 * calling convention (src/sip/worker/special_instructions.h:27-97): per argument the 6-tuple
 * (array_slot, rank, index_values, size, extents, data) passed by reference, then ierr.  `data` are DEVICE
 * pointers (resident blocks); everything else is host memory.  ierr: 0 = ok, else SIPGPU_E_*; the int return
 * value repeats it.  The routines are asynchronous on the compute stream.  They replace the Fortran routines
 * named in each comment; the maintainer registers them under the same SIAL names.
 * The predefined int array "moa_seg_ranges" that the Fortran code fetches through the sip_interface upcall
 * (sip_interface.h:17-35) is registered once per run with sipgpu_set_predefined_int_array.
 * --------------------------------------------------------------------------------------------- */
int sipgpu_set_predefined_int_array(const char* name, int n, const int* values);
/* qm-generic/energy_denominator_rhf.F:15-460  (special energy_denominator_rhf ur): array_0 /= eps(Fock diagonal);
 * array_0 rank 2, 4 or 6; array_1 = Fock array, rank 1 (diagonal) or 2 (full matrix, device) */
int sipgpu_si_energy_denominator_rhf(int* array_0, int* rank_0, int* index_values_0, int* size_0, int* extents_0, double* data_0,
                                     int* array_1, int* rank_1, int* index_values_1, int* size_1, int* extents_1, double* data_1,
                                     int* ierr);
/* qm-generic/energy_ty_denominator_rhf.F (special energy_ty_denominator_rhf urr): rank-4 array_0 /= eps + shift, array_1 = Fock
 * matrix (rank 2, device), array_2 = the scalar `shift` as a one-element DEVICE block */
int sipgpu_si_energy_ty_denominator_rhf(int* array_0, int* rank_0, int* index_values_0, int* size_0, int* extents_0, double* data_0,
                                        int* array_1, int* rank_1, int* index_values_1, int* size_1, int* extents_1, double* data_1,
                                        int* array_2, int* rank_2, int* index_values_2, int* size_2, int* extents_2, double* data_2,
                                        int* ierr);
/* qm/utility/stripi.F (special stripi ru): array_1 (one index thick in its stripped dimensions) = the matching
 * strip of array_0; ranks 2-4 */
int sipgpu_si_stripi(int* array_0, int* rank_0, int* index_values_0, int* size_0, int* extents_0, double* data_0,
                     int* array_1, int* rank_1, int* index_values_1, int* size_1, int* extents_1, double* data_1, int* ierr);
/* qm/utility/anti_symm_o.F / anti_symm_v.F: antisymmetrise a rank-4 block [a,i,b,j] in (i,j) / (a,b) in place */
int sipgpu_si_anti_symm_o(int* array_0, int* rank_0, int* index_values_0, int* size_0, int* extents_0, double* data_0, int* ierr);
int sipgpu_si_anti_symm_v(int* array_0, int* rank_0, int* index_values_0, int* size_0, int* extents_0, double* data_0, int* ierr);
/* qm/utility/return_sval.F (special return_sval rw): device scalar data_1[0] = last element of the rank-1/2 block */
int sipgpu_si_return_sval(int* array_0, int* rank_0, int* index_values_0, int* size_0, int* extents_0, double* data_0,
                          int* array_1, int* rank_1, int* index_values_1, int* size_1, int* extents_1, double* data_1, int* ierr);
/* qm/utility/invert_diagonal.F (special invert_diagonal ur): array_0 /= array_1 where array_1 != 0; ranks 3, 5 */
int sipgpu_si_invert_diagonal(int* array_0, int* rank_0, int* index_values_0, int* size_0, int* extents_0, double* data_0,
                              int* array_1, int* rank_1, int* index_values_1, int* size_1, int* extents_1, double* data_1,
                              int* ierr);
/* qm/utility/invert_diagonal_asym.F (special invert_diagonal_asym ur): rank-5 blocks [k,a,i,a1,i1]: array_0 /= array_1 where
 * a != a1 and i != i1 (orbital numbers) and array_1 != 0, array_0 = 0 on the a == a1 or i == i1 planes */
int sipgpu_si_invert_diagonal_asym(int* array_0, int* rank_0, int* index_values_0, int* size_0, int* extents_0, double* data_0,
                                   int* array_1, int* rank_1, int* index_values_1, int* size_1, int* extents_1, double* data_1,
                                   int* ierr);
/* qm/utility/return_diagonal_elements.F (special return_diagonal_elements u): keep x(p,p) (rank 2) / x(p,p,r,r) (rank 4) of a
 * block that is square per index pair, zero everything else */
int sipgpu_si_return_diagonal_elements(int* array_0, int* rank_0, int* index_values_0, int* size_0, int* extents_0, double* data_0,
                                       int* ierr);

/* ---------------------------------------------------------------------------------------------
 * Boundary 3c -- deferred op stream: the batching front-end for pardo bodies.
 * The reference interpreter dispatches one block operation per opcode (interpreter.cpp:98-910, e.g. the
 * `do k: T = A*B; D += T` bodies of rlccd_rhf.sialx:342-355).  Between sipgpu_wl_begin() and sipgpu_wl_end()
 * the asynchronous entry points of boundaries 2-4 (block alloc/free, fill/scale/copy/axpy/add, permute,
 * contract, slice/insert, array get/put/put_accumulate, _gpu_* equivalents) are RECORDED with their
 * read/write sets instead of launched.  A flush schedules the stream: temps that only feed an accumulate
 * are forwarded (T = L*R; D += f*T  ->  D += f*L*R, the temp is never written), accumulating contractions
 * into one destination are chained, hazards on device address intervals give levels, and each level is
 * emitted as ONE launch per kernel family (worklist.cu).  Program-order semantics are preserved; blocking
 * calls (d2h, dot, norm2, sync, host-pointer ABI) flush implicitly.  The natural call sites in the
 * reference are pardo entry and SialOpsParallel::sip_barrier / endpardo (sial_ops_parallel.cpp:39-99).
 * --------------------------------------------------------------------------------------------- */
#define SIPGPU_WL_DRY 1 /* record and schedule on the host only (no device): planner tests */
int sipgpu_wl_begin(int flags);
int sipgpu_wl_flush(void);      /* schedule + launch what is recorded; recording continues */
int sipgpu_wl_end(void);        /* flush and stop recording */
int sipgpu_wl_recording(void);  /* 0 = off, 1 = recording, 2 = dry recording */
/* automatic flush thresholds (0 = keep): recorded ops, bytes of temps whose free is deferred.  The three setters
 * change the open recording when called inside one, the defaults of later recordings otherwise. */
int sipgpu_wl_set_limits(long long max_ops, long long max_deferred_bytes);
/* A recording is also flushed early whenever the compute stream has drained and at least min_ops ops are pending, so
 * that the device works on batch k while the host records batch k+1 (default 4096; 0 = only flush at the limits). */
int sipgpu_wl_set_idle_flush(long long min_ops);
/* out9 = {ops recorded, ops scheduled, levels, kernel launches, temp->accumulate fusions, chains,
 *         operand pairs in chains, temps never written, flushes} since sipgpu_wl_begin */
int sipgpu_wl_stats(long long* out9);
/* plan of the most recent flush, per recorded op in program order: the level it ran in and the index of the
 * op it was fused into (itself if not fused).  Returns the number of ops of that flush. */
/* A stream that was seen before -- the same ops on the same blocks in the same order, which is what every iteration of a CC
 * program records (interpreter.cpp:98-910 walks the same pardo bodies) -- is not scheduled again: its launches were captured
 * (descriptors resident on the device) and are replayed.  Streams are recognised by a 128-bit content hash folded while
 * recording; streams with opaque ops (super-instructions, slices) are always scheduled.  Returns the number of flushes of
 * the open / last recording that were served by a replay.  SIPGPU_WL_REPLAY=0 in the environment disables the cache. */
long long sipgpu_wl_replays(void);
int sipgpu_wl_last_plan(int cap, int* level_of_op, int* unit_of_op);

/* ---------------------------------------------------------------------------------------------
 * Boundary 4 -- distributed / served arrays (SialOpsParallel method set, src/sip/worker/
 * sial_ops_parallel.cpp:39-99,132-171,232-284,332-408,549-565; owner rule data_distribution.cpp:19-82 and
 * array_table.cpp:50-97).  One process per GPU; every GPU is worker AND owner (no server ranks).  Each
 * rank exports its slab of the array through CUDA IPC; peers map it and get / put / put+= go directly
 * over NVLink (peer loads, peer stores, red.global.add.f64).  The bootstrap exchange of the IPC handles
 * is done by the caller (torch.distributed / MPI): the library only produces and consumes opaque bytes.
 * --------------------------------------------------------------------------------------------- */
typedef struct sipgpu_array sipgpu_array; /* opaque */

#define SIPGPU_IPC_HANDLE_BYTES 64

/* nseg[rank]: number of segments per index; seg_ext: concatenated extents per index (sum nseg entries).
 * Blocks are numbered as array_table.cpp:75-81 (LAST index fastest) and owned by block_number % world. */
int sipgpu_array_create(int rank, const int* nseg, const int* seg_ext, int my_rank, int world, sipgpu_array** out);
int sipgpu_array_destroy(sipgpu_array* a);
int sipgpu_array_export(sipgpu_array* a, void* handle_bytes /* SIPGPU_IPC_HANDLE_BYTES */);
int sipgpu_array_attach(sipgpu_array* a, int peer_rank, const void* handle_bytes, int peer_device);
long long sipgpu_array_block_number(const sipgpu_array* a, const int* idx /* 1-based segment numbers */);
int sipgpu_array_block_owner(const sipgpu_array* a, long long block_number);
long long sipgpu_array_block_size(const sipgpu_array* a, const int* idx);
/* device address of block `idx` in its OWNER's slab as mapped into this process (peer memory when remote) */
double* sipgpu_array_block_ptr(sipgpu_array* a, const int* idx);
int sipgpu_array_get(sipgpu_array* a, const int* idx, double* g_dst);            /* SialOpsParallel::get            */
int sipgpu_array_put(sipgpu_array* a, const int* idx, const double* g_src);      /* ::put_replace                   */
int sipgpu_array_put_accumulate(sipgpu_array* a, const int* idx, const double* g_src); /* ::put_accumulate (atomic)  */
/* a barrier section's worth of get / put += as ONE launch: idx = n x rank segment numbers, one block pointer per entry */
int sipgpu_array_get_many(sipgpu_array* a, int n, const int* idx, double* const* g_dst);
int sipgpu_array_put_accumulate_many(sipgpu_array* a, int n, const int* idx, const double* const* g_src);
int sipgpu_array_put_initialize(sipgpu_array* a, const int* idx, double value);  /* ::put_initialize :412-446, block = v */
int sipgpu_array_put_increment(sipgpu_array* a, const int* idx, double delta);   /* ::put_increment  :448-487, block += d */
int sipgpu_array_put_scale(sipgpu_array* a, const int* idx, double factor);      /* ::put_scale      :489-528, block *= f */
int sipgpu_array_fill_local(sipgpu_array* a, double v);                          /* all owned blocks = v (array creation / `T2new = 0`) */
size_t sipgpu_array_local_bytes(const sipgpu_array* a);
double* sipgpu_array_local_base(sipgpu_array* a);   /* this rank's slab (its owned blocks, contiguous) */
/* ---- race detection between barriers (src/sip/dynamic_data/distributed_block_consistency.cpp:25-175) ----
 * The reference's server checks every get / put / put += against a per-block state table; read as a whole the table
 * accepts the accesses of one barrier section to one block iff one worker made all of them, or all are GETs, or all
 * are PUT_ACCUMULATEs.  Here every rank records what it touched (when tracking is on), the caller's sip_barrier
 * exchanges the summaries, and each rank validates the union (host-only).  SIPGPU_E_STATE = inconsistent block. */
#define SIPGPU_ACCESS_GET 1
#define SIPGPU_ACCESS_PUT 2
#define SIPGPU_ACCESS_PUT_ACCUMULATE 4
int sipgpu_array_track_accesses(sipgpu_array* a, int on);
/* blocks this rank touched since the last reset, ascending block number; returns their count (call with cap 0 first) */
long long sipgpu_array_section_accesses(sipgpu_array* a, long long cap, long long* block_numbers, int* access_bits);
int sipgpu_array_section_reset(sipgpu_array* a);   /* at sip_barrier (reset_consistency_status) */
int sipgpu_consistency_validate(long long n, const long long* block_numbers, const int* access_bits, const int* workers,
                                long long* bad_block);

/* ---- persistence ----
 * Label registry = WorkerPersistentArrayManager::set_persistent / restore_persistent
 * (src/sip/dynamic_data/worker_persistent_array_manager.cpp:34-155): objects marked persistent at the end of one SIAL
 * program are handed to the next by label; scalars by value, contiguous arrays (a pool block) and distributed arrays by
 * ownership transfer, so they stay resident in HBM.  restore removes the label; a repeated label overwrites (the
 * replaced object is released).  SIPGPU_E_STATE when the label is unknown. */
int sipgpu_persist_scalar(const char* label, double value);
int sipgpu_restore_scalar(const char* label, double* value);
int sipgpu_persist_contiguous(const char* label, double* dev_block, int rank, const int* ext);
int sipgpu_restore_contiguous(const char* label, double** dev_block, int* rank, int* ext6 /* MAX_RANK entries */);
int sipgpu_persist_array(const char* label, sipgpu_array* a);
int sipgpu_restore_array(const char* label, sipgpu_array** a);
int sipgpu_persist_count(int* nscalars, int* ncontiguous, int* narrays);
/* checkpoint_persistent / init_from_checkpoint (:157-260): scalars + contiguous arrays of the registry in the
 * reference's own byte format (little-endian stream of setup/io_utils.cpp:44-70,114-155), so checkpoints are
 * interchangeable with the reference worker's.  init requires an empty registry (the reference CHECKs the same). */
int sipgpu_persist_checkpoint(const char* filename);
int sipgpu_persist_init_from_checkpoint(const char* filename);
/* server-side array files (src/sip/mpi/array_file.h:53-70): this rank's slab as <int chunk_size><int num_servers>
 * <double>* plus a dense index file <77><nblocks><byte offset per block number, -1 for blocks of other ranks>.
 * load validates header and index against the array's layout and world size. */
int sipgpu_array_save(sipgpu_array* a, const char* data_path, const char* index_path);
int sipgpu_array_load(sipgpu_array* a, const char* data_path, const char* index_path);

/* host-only layout arithmetic (no device needed) */
long long sipgpu_layout_block_number(int rank, const int* nseg, const int* idx);
int sipgpu_layout_block_owner(long long block_number, int world);

#ifdef __cplusplus
}
#endif
#endif /* SIPGPU_H_ */

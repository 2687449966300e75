// plan.h -- host-side planner: label/pattern analysis of the SIAL `contract` and `transpose`
// super-instructions, turned into stride tables the fused sm_100a kernels consume directly.
//
// Reference behaviour followed (not code): tensor_dil_omp.F90:87-142 (get_contraction_ptrn),
// :799-859 (determine_index_permutations), :861-896 (contr_ptrn_ok); interpreter.cpp:1210-1262,
// 2049-2084; block.cpp:216-255.
#pragma once
#include "common.h"

namespace sipgpu {

// One contraction  D'(m,n) = sum_k L'(k,m) R'(k,n)  expressed on the ORIGINAL (unpermuted) blocks: the
// three index groups with, per (collapsed) dimension, its extent and its element stride in each block it
// appears in.  Address of L'(k,m) in L = sum_i kidx_i*ksL[i] + sum_j midx_j*msL[j], etc.  The reference
// materialises L', R', D' with three permute passes (F90:731-785); here the permutes are folded into the
// kernel's global->shared loads and its epilogue stores.
struct Shape {
    int M, N, K;     // lld, lrd, lcd of F90:799-859
    int nm, nn, nk;  // dims per group after collapsing
    int mext[kMaxRank], msL[kMaxRank], msD[kMaxRank];
    int next[kMaxRank], nsR[kMaxRank], nsD[kMaxRank];
    int kext[kMaxRank], ksL[kMaxRank], ksR[kMaxRank];
    int a_kc, b_kc;  // operand's stride-1 dimension is contracted (K-contiguous) vs free (M/N-contiguous)
    int vec;         // both operands may be fetched as 16-byte pairs along their contiguous direction
    int tile;        // CTA tile of the kernel menu (contract_pick_tile), set by the launcher
    int swapped;     // the launcher exchanged the roles of L and R (small free dimension becomes n): pairs are {R, L}
};

// One destination block and the chain of (L, R) operand pairs summed into it:
//   D = alpha * sum_p L_p * R_p + beta * D        (all pairs share the Shape)
// A chain is how a block-sparse contraction over a segmented contracted index reaches the kernel: the sum over
// contracted SEGMENTS (the `do`-loop + `put +=` of a pardo body) runs inside one tile's accumulators, so the
// destination is written once instead of read-modify-written once per segment.
struct Pair {
    const double* L;
    const double* R;
};
struct Problem {
    double* D;
    int shape;        // index into the launch's Shape array
    int chain_begin;  // first Pair of the chain
    int chain_len;
    int kwin;  // split-K: k windows [kwin & 0xffff, kwin >> 16) of contract.cu's KWIN-element windows; 0 = the whole K
};

// F90:87-142.  Returns the reference's ierr (0 ok, 1..6).
int get_contraction_ptrn(int drank, int lrank, int rrank, const int* aces_ptrn, int* my_ptrn);

// F90:861-896 + :799-859: validate the pattern against the extents and build the fused-kernel Shape.
// Returns 0, or 1 (the reference's "invalid contraction pattern" ierr), or SIPGPU_E_ARG.
// Requires lrank,rrank,drank >= 1 (rank-0 operands are routed to the dot/axpy kernels by the caller).
int build_shape(const int* ptrn, int lrank, const int* lext, int rrank, const int* rext, int drank, const int* dext,
                Shape* out);
int build_shape_strided(const int* ptrn, int lrank, const int* lext, const int* lpar, int rrank, const int* rext,
                        const int* rpar, int drank, const int* dext, const int* dpar, Shape* out);
// The same contraction with the roles of the operands exchanged: D'(n,m) = sum_k R'(k,n) L'(k,m).  Used to put the
// SMALL free dimension on the n side, where the narrow tiles of the kernel menu live.
Shape swap_operands(const Shape& s);
bool contr_ptrn_ok(const int* ptrn, int lrank, int rrank, int drank, const int* lext, const int* rext,
                   const int* dext);

// interpreter.cpp:2049-2084 + block.cpp:227-231: transp[0]=+1, transp[j+1] = 1-based NEW position of rhs dim j.
int permutation_from_labels(int rank, const int* lhs_labels, const int* rhs_labels, int* transp);

// Permute plan: collapsed dims, in input order.
struct PermShape {
    int rank;               // after collapsing
    int ext[kMaxRank];      // input-order extents
    int in_stride[kMaxRank];
    int out_stride[kMaxRank];  // stride in the output of input dim i
    long long total;
};
int build_perm_shape(int rank, const int* ext, const int* transp, PermShape* out);

}  // namespace sipgpu

// abi.cu -- the extern "C" entry points of include/sipgpu.h (boundaries 1-3) and the host-side
// orchestration of the contraction: pattern analysis -> Shape -> one fused-kernel launch.
#include <map>
#include <vector>

#include "contract.h"
#include "elementwise.h"
#include "plan.h"

#include <array>
#include <vector>
#include "worklist.h"

namespace sipgpu {
int run_worklist(int n, std::vector<Shape>& shapes, const std::vector<int>& pshape, const int* chain_start,
                 const double* const* L, const double* const* R, double* const* D, double alpha, double beta);

namespace {

long long volume(int rank, const int* ext) {
    long long n = 1;
    for (int i = 0; i < rank; ++i) n *= ext[i];
    return n;
}

// d[i] = alpha * s[i] * r[0] + beta * d[i]   (tensor x rank-0 block: F90:771-774)
__global__ void __launch_bounds__(256) scal_dev_kernel(double* __restrict__ d, const double* __restrict__ s, long long n,
                                                       const double* __restrict__ r, double alpha, double beta) {
    const double f = alpha * r[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        d[i] = (beta != 0.0) ? f * s[i] + beta * d[i] : f * s[i];
}

int scaled_by_device_scalar(double* d, const double* s, long long n, const double* r, double alpha, double beta) {
    long long b = (n + 255) / 256;
    if (b > ctx().num_sms * 8) b = ctx().num_sms * 8;
    scal_dev_kernel<<<(int)b, 256, 0, ctx().stream>>>(d, s, n, r, alpha, beta);
    SIP_CUDA(cudaGetLastError());
    count_launch();
    return SIPGPU_OK;
}

struct TempBlock {  // pool block released at scope exit (stream order keeps it alive for queued kernels)
    double* p = nullptr;
    ~TempBlock() { if (p) pool_free(p); }
    int get(long long n) {
        p = pool_alloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        return p ? SIPGPU_OK : SIPGPU_E_NOMEM;
    }
};

int tiles_of(const Shape& s) {
    int bm, bn;
    contract_tile_dims(s.tile, &bm, &bn);
    return ((s.M + bm - 1) / bm) * ((s.N + bn - 1) / bn);
}

// rank-0 special cases of tensor_block_contract_ (F90:764-780)
int contract_degenerate(const int* ptrn, const double* L, int lrank, const int* lext, const double* R, int rrank,
                        const int* rext, double* D, int drank, const int* dext, double alpha, double beta) {
    if (!contr_ptrn_ok(ptrn, lrank, rrank, drank, lext, rext, dext)) return 1;
    if (drank == 0 && lrank == 0 && rrank == 0) {  // scalar * scalar
        return scaled_by_device_scalar(D, L, 1, R, alpha, beta);
    }
    if (drank == 0) {  // full contraction to a scalar: permute L into R's index order, then a dot product
        int transp[33];
        transp[0] = 1;
        bool trivial = true;
        for (int j = 0; j < lrank; ++j) {
            transp[j + 1] = -ptrn[j];
            trivial = trivial && transp[j + 1] == j + 1;
        }
        const long long n = volume(lrank, lext);
        TempBlock tmp;
        const double* lp = L;
        if (!trivial) {
            SIP_TRY(tmp.get(n));
            SIP_TRY(permute_block(lrank, lext, transp, L, tmp.p));
            lp = tmp.p;
        }
        if (alpha != 1.0) {  // D = beta*D + alpha*dot: fold alpha through a scaled scratch result
            SIP_TRY(ew_dot_device(lp, R, n, ctx().d_reduce + 1, 0.0));
            return scaled_by_device_scalar(D, ctx().d_reduce + 1, 1, ctx().d_reduce + 2, alpha, beta);
        }
        return ew_dot_device(lp, R, n, D, beta);
    }
    // tensor * rank-0 block: D[ptrn[j]] <- T dim j, times the scalar
    const bool l_is_tensor = lrank > 0;
    const double* T = l_is_tensor ? L : R;
    const double* S = l_is_tensor ? R : L;
    const int trank = l_is_tensor ? lrank : rrank;
    const int* text = l_is_tensor ? lext : rext;
    const int* tp = l_is_tensor ? ptrn : ptrn + lrank;
    int transp[33];
    transp[0] = 1;
    bool trivial = true;
    for (int j = 0; j < trank; ++j) {
        transp[j + 1] = tp[j];
        trivial = trivial && tp[j] == j + 1;
    }
    const long long n = volume(trank, text);
    if (trivial) return scaled_by_device_scalar(D, T, n, S, alpha, beta);
    TempBlock tmp;
    SIP_TRY(tmp.get(n));
    SIP_TRY(permute_block(trank, text, transp, T, tmp.p));
    return scaled_by_device_scalar(D, tmp.p, n, S, alpha, beta);
}

// Every operand may be a slice of a larger dense array: (*par = parent extents, *beg = 0-based first element of the
// slice in the parent; nullptr = dense block).
int contract_device_sliced(const int* ptrn, const double* L, int lrank, const int* lext, const int* lpar, const int* lbeg,
                           const double* R, int rrank, const int* rext, const int* rpar, const int* rbeg, double* D, int drank,
                           const int* dext, const int* dpar, const int* dbeg, double alpha, double beta) {
    SIP_TRY(ensure_init());
    if (!L || !R || !D || lrank < 0 || rrank < 0 || drank < 0 || lrank > 32 || rrank > 32 || drank > 32) return SIPGPU_E_ARG;
    if (lrank == 0 || rrank == 0 || drank == 0) {
        if (lpar || rpar || dpar) return SIPGPU_E_ARG;  // scalar-operand forms take dense blocks only
        return contract_degenerate(ptrn, L, lrank, lext, R, rrank, rext, D, drank, dext, alpha, beta);
    }
    auto base_offset = [](int rank, const int* ext, const int* par, const int* beg, long long* off) {
        *off = 0;
        if (!par) return beg ? SIPGPU_E_ARG : SIPGPU_OK;
        long long s = 1;
        for (int i = 0; i < rank; ++i) {
            const int b = beg ? beg[i] : 0;
            if (b < 0 || b + ext[i] > par[i]) return SIPGPU_E_ARG;
            *off += s * b;
            s *= par[i];
        }
        return SIPGPU_OK;
    };
    long long ol, orr, od;
    SIP_TRY(base_offset(lrank, lext, lpar, lbeg, &ol));
    SIP_TRY(base_offset(rrank, rext, rpar, rbeg, &orr));
    SIP_TRY(base_offset(drank, dext, dpar, dbeg, &od));
    ContractArgs a;
    memset(&a, 0, sizeof(a));
    SIP_TRY(build_shape_strided(ptrn, lrank, lext, lpar, rrank, rext, rpar, drank, dext, dpar, &a.s0));
    // D is a single element (all free indices have extent 1 -- e.g. the DIIS matrix element
    // BB[k1,k2] = E[a,i,b,j,k1]*E[a,i,b,j,k2] with simple indices k1, k2) and the contracted elements are one contiguous
    // run in both operands: that is a dot product for the reduction kernels, not a 1x1 tile for the tensor cores
    if (a.s0.M == 1 && a.s0.N == 1 && a.s0.nk == 1 && a.s0.ksL[0] == 1 && a.s0.ksR[0] == 1 && !lpar && !rpar) {
        if (alpha != 1.0) {
            SIP_TRY(ew_dot_device(L + ol, R + orr, a.s0.K, ctx().d_reduce + 1, 0.0));
            return scaled_by_device_scalar(D + od, ctx().d_reduce + 1, 1, ctx().d_reduce + 2, alpha, beta);
        }
        return ew_dot_device(L + ol, R + orr, a.s0.K, D + od, beta);
    }
    if (a.s0.M < a.s0.N && a.s0.M <= 64) a.s0 = swap_operands(a.s0);  // the small free dimension goes to the n side
    if (lowint_eligible(a.s0)) {  // bandwidth-shaped kernel (strided operands and a strided destination included: no split then)
        std::vector<Pair> pairs(1, a.s0.swapped ? Pair{R + orr, L + ol} : Pair{L + ol, R + orr});
        std::vector<int> chain = {0, 1};
        double* Dp = D + od;
        return lowint_launch(a.s0, 1, pairs, chain, &Dp, alpha, beta, dpar == nullptr);
    }
    a.s0.tile = contract_pick_tile(a.s0.M, a.s0.N);
    if (contract_tile_count(a.s0.M, a.s0.N, a.s0.tile) < ctx().num_sms) a.s0.tile = kSmallTile;  // spread a small block
    // one small destination with a long contracted range: split-K through the work-list path (dense D only: the
    // beta pre-pass of the split treats D as one contiguous run)
    if (!dpar && contract_tile_count(a.s0.M, a.s0.N, a.s0.tile) * 2 <= ctx().num_sms * 2 && a.s0.K >= 2 * 4096) {
        std::vector<Shape> shapes(1, a.s0);
        std::vector<int> pshape(1, 0);
        const double* Lp = L + ol;
        const double* Rp = R + orr;
        double* Dp = D + od;
        return run_worklist(1, shapes, pshape, nullptr, &Lp, &Rp, &Dp, alpha, beta);
    }
    a.nprob = 1;
    a.total_tiles = tiles_of(a.s0);
    a.alpha = alpha;
    a.beta = beta;
    a.pair0.L = a.s0.swapped ? R + orr : L + ol;
    a.pair0.R = a.s0.swapped ? L + ol : R + orr;
    a.p0.D = D + od;
    a.p0.chain_len = 1;
    const bool vec = a.s0.vec && (((uintptr_t)a.pair0.L | (uintptr_t)a.pair0.R) & 15) == 0;
    return launch_contract(a, a.s0.a_kc, a.s0.b_kc, vec, a.s0.tile);
}

int contract_device(const int* ptrn, const double* L, int lrank, const int* lext, const double* R, int rrank,
                    const int* rext, double* D, int drank, const int* dext, double alpha, double beta) {
    return contract_device_sliced(ptrn, L, lrank, lext, nullptr, nullptr, R, rrank, rext, nullptr, nullptr, D, drank, dext,
                                  nullptr, nullptr, alpha, beta);
}

}  // namespace

// n destination blocks; destination i sums the operand pairs chain_start[i] .. chain_start[i+1]-1 of L[]/R[].
int contract_chained(int n, const int* ptrn, int lrank, int rrank, int drank, const int* lext, const int* rext,
                     const int* dext, const int* chain_start, const double* const* L, const double* const* R,
                     double* const* D, double alpha, double beta) {
    SIP_TRY(ensure_init());
    if (n < 0 || lrank < 1 || rrank < 1 || drank < 1 || lrank > kMaxRank || rrank > kMaxRank || drank > kMaxRank)
        return SIPGPU_E_ARG;
    if (n == 0) return SIPGPU_OK;
    // distinct extent tuples -> Shapes (a pardo body has a handful of them: segment sizes are few)
    std::map<std::vector<int>, int> shape_ids;
    std::vector<Shape> shapes;
    std::vector<int> pshape(n);
    for (int i = 0; i < n; ++i) {
        // work-lists are long runs of identical extents: compare with the previous problem before touching the map
        if (i > 0 && !memcmp(lext + (size_t)i * lrank, lext + (size_t)(i - 1) * lrank, sizeof(int) * lrank) &&
            !memcmp(rext + (size_t)i * rrank, rext + (size_t)(i - 1) * rrank, sizeof(int) * rrank) &&
            !memcmp(dext + (size_t)i * drank, dext + (size_t)(i - 1) * drank, sizeof(int) * drank)) {
            pshape[i] = pshape[i - 1];
            continue;
        }
        std::vector<int> key(lext + (size_t)i * lrank, lext + (size_t)(i + 1) * lrank);
        key.insert(key.end(), rext + (size_t)i * rrank, rext + (size_t)(i + 1) * rrank);
        key.insert(key.end(), dext + (size_t)i * drank, dext + (size_t)(i + 1) * drank);
        auto it = shape_ids.find(key);
        if (it == shape_ids.end()) {
            Shape s;
            SIP_TRY(build_shape(ptrn, lrank, lext + (size_t)i * lrank, rrank, rext + (size_t)i * rrank, drank,
                                dext + (size_t)i * drank, &s));
            if (s.M < s.N && s.M <= 64) s = swap_operands(s);
            s.tile = contract_pick_tile(s.M, s.N);
            it = shape_ids.emplace(key, (int)shapes.size()).first;
            shapes.push_back(s);
        }
        pshape[i] = it->second;
    }
    return run_worklist(n, shapes, pshape, chain_start, L, R, D, alpha, beta);
}

// The launch half of a work-list: tile choice for under-filled launches, split-K, dot routing, one launch per kernel
// variant.  shapes[pshape[i]] is the Shape of problem i (operand roles already swapped where Shape::swapped).
int run_worklist(int n, std::vector<Shape>& shapes, const std::vector<int>& pshape, const int* chain_start,
                 const double* const* L, const double* const* R, double* const* D, double alpha, double beta) {
    // ---- low arithmetic intensity shapes (rank-2 results of rank-4 blocks, one segment-sized free + contracted index,
    // matrix-vector shapes, tiny matrices, dot products) go to the bandwidth-shaped kernel, one launch per shape ----
    {
        std::vector<char> low(shapes.size(), 0);
        bool any = false, all = true;
        for (size_t k = 0; k < shapes.size(); ++k) low[k] = lowint_eligible(shapes[k]) ? 1 : 0;
        for (int i = 0; i < n; ++i) { any = any || low[pshape[i]]; all = all && low[pshape[i]]; }  // over PROBLEMS, not shapes
        if (any) {
            std::vector<int> rest_idx;
            for (size_t k = 0; k < shapes.size(); ++k) {
                if (!low[k]) continue;
                std::vector<Pair> pairs;
                std::vector<int> chain(1, 0);
                std::vector<double*> dd;
                for (int i = 0; i < n; ++i) {
                    if (pshape[i] != (int)k) continue;
                    const int c0 = chain_start ? chain_start[i] : i, c1 = chain_start ? chain_start[i + 1] : i + 1;
                    if (!D[i] || c1 <= c0) return SIPGPU_E_ARG;
                    for (int c = c0; c < c1; ++c) {
                        if (!L[c] || !R[c]) return SIPGPU_E_ARG;
                        pairs.push_back(shapes[k].swapped ? Pair{R[c], L[c]} : Pair{L[c], R[c]});
                    }
                    chain.push_back((int)pairs.size());
                    dd.push_back(D[i]);
                }
                if (!dd.empty()) SIP_TRY(lowint_launch(shapes[k], (int)dd.size(), pairs, chain, dd.data(), alpha, beta, true));
            }
            if (all) return SIPGPU_OK;
            // the remaining problems run below: compact the work-list
            std::vector<int> pshape2, chain2(1, 0);
            std::vector<const double*> L2, R2;
            std::vector<double*> D2;
            for (int i = 0; i < n; ++i) {
                if (low[pshape[i]]) continue;
                const int c0 = chain_start ? chain_start[i] : i, c1 = chain_start ? chain_start[i + 1] : i + 1;
                for (int c = c0; c < c1; ++c) { L2.push_back(L[c]); R2.push_back(R[c]); }
                chain2.push_back((int)L2.size());
                D2.push_back(D[i]);
                pshape2.push_back(pshape[i]);
            }
            return run_worklist((int)D2.size(), shapes, pshape2, chain2.data(), L2.data(), R2.data(), D2.data(), alpha, beta);
        }
    }
    long long total = 0;
    for (int i = 0; i < n; ++i) total += contract_tile_count(shapes[pshape[i]].M, shapes[pshape[i]].N, shapes[pshape[i]].tile);
    if (total < ctx().num_sms) {  // a work-list that cannot fill the SMs with large tiles runs on 64x64 tiles (two CTAs per SM)
        for (Shape& sh : shapes) sh.tile = kSmallTile;
        total = 0;
        for (int i = 0; i < n; ++i) total += contract_tile_count(shapes[pshape[i]].M, shapes[pshape[i]].N, kSmallTile);
    }
    // single-element destinations with contiguous contracted runs are dot products (see contract_device_sliced)
    // (short dots stay in the batched tensor-core launch: two reduction launches per block would be launch-bound)
    auto is_dot = [&](const Shape& s) { return alpha == 1.0 && s.M == 1 && s.N == 1 && s.K >= 16384 && s.nk == 1 && s.ksL[0] == 1 && s.ksR[0] == 1; };
    for (int i = 0; i < n; ++i) {
        const Shape& s = shapes[pshape[i]];
        if (!is_dot(s)) continue;
        const int c0 = chain_start ? chain_start[i] : i, c1 = chain_start ? chain_start[i + 1] : i + 1;
        if (!D[i] || c1 <= c0) return SIPGPU_E_ARG;
        if (capture()) { capture()->failed = true; return SIPGPU_E_STATE; }  // not a capturable launch site
        for (int c = c0; c < c1; ++c) {
            if (!L[c] || !R[c]) return SIPGPU_E_ARG;
            SIP_TRY(ew_dot_device(L[c], R[c], s.K, D[i], c == c0 ? beta : 1.0));
        }
    }
    // ---- split-K: a launch that leaves most CTA slots idle while every tile walks a long contracted range (rank-2
    // results of rank-4 blocks: D[a,b] = L[a,i,c,j]*R[b,i,c,j], K = o*v*o per pair and a chain of pairs on top) is cut
    // along the chain and along the k windows into `split` partial problems per destination; the partial sums meet in
    // D through red.global.add.f64 (beta is applied to D by one batched pre-pass) ----
    const int slots = ctx().num_sms * 2;  // small / narrow tiles run two CTAs per SM
    std::vector<int> split(n, 1);
    bool any_split = false;
    if (total > 0 && total * 2 <= slots) {
        const int want = (int)std::min<long long>(16, slots / total);
        for (int i = 0; i < n; ++i) {
            const Shape& s = shapes[pshape[i]];
            if (is_dot(s)) continue;
            const int c0 = chain_start ? chain_start[i] : i, c1 = chain_start ? chain_start[i + 1] : i + 1;
            const long long nwin = (s.K + kContractKWin - 1) / kContractKWin;
            // at least 4096 contracted elements per part, so that a part is worth a tile's prologue and epilogue
            const long long parts = std::min<long long>((long long)(c1 - c0) * nwin, (long long)(c1 - c0) * s.K / 4096);
            split[i] = (int)std::max<long long>(1, std::min<long long>(want, parts));
            any_split = any_split || split[i] > 1;
        }
    }
    if (any_split) {
        std::vector<double*> dd;
        std::vector<long long> cnt;
        for (int i = 0; i < n; ++i) {
            const Shape& s = shapes[pshape[i]];
            if (is_dot(s) || !D[i]) continue;
            dd.push_back(D[i]);
            cnt.push_back((long long)s.M * s.N);
        }
        // dense destinations only (a sliced destination never comes here: see contract_device_sliced)
        if (beta != 1.0) {
            if (Capture* cap = capture())
                cap->steps.push_back([dd, cnt, beta]() mutable { return ew_scale_many((int)dd.size(), dd.data(), cnt.data(), beta); });
            else
                SIP_TRY(ew_scale_many((int)dd.size(), dd.data(), cnt.data(), beta));
        }
    }
    // kernel variant of every problem: (a_kc, b_kc, 16-byte loads, tile)
    std::vector<int> pvariant(n, -1);
    bool used[40] = {false};
    for (int i = 0; i < n; ++i) {
        const Shape& s = shapes[pshape[i]];
        if (is_dot(s)) continue;
        const int c0 = chain_start ? chain_start[i] : i, c1 = chain_start ? chain_start[i + 1] : i + 1;
        if (!D[i] || c1 <= c0) return SIPGPU_E_ARG;
        bool pvec = s.vec != 0;
        for (int c = c0; c < c1 && pvec; ++c) pvec = (((uintptr_t)L[c] | (uintptr_t)R[c]) & 15) == 0;
        pvariant[i] = (s.a_kc ? 1 : 0) | (s.b_kc ? 2 : 0) | (pvec ? 4 : 0) | (s.tile << 3);
        used[pvariant[i]] = true;
    }
    for (int variant = 0; variant < 40; ++variant) {
        if (!used[variant]) continue;
        const bool a_kc = variant & 1, b_kc = variant & 2, vec = variant & 4;
        const int tile = variant >> 3;
        std::vector<Problem> probs;
        std::vector<Pair> pairs;
        std::vector<int> prefix(1, 0);
        for (int i = 0; i < n; ++i) {
            if (pvariant[i] != variant) continue;
            const Shape& s = shapes[pshape[i]];
            const int c0 = chain_start ? chain_start[i] : i, c1 = chain_start ? chain_start[i + 1] : i + 1;
            const int first_pair = (int)pairs.size(), len = c1 - c0;
            for (int c = c0; c < c1; ++c) {
                if (!L[c] || !R[c]) return SIPGPU_E_ARG;
                pairs.push_back(s.swapped ? Pair{R[c], L[c]} : Pair{L[c], R[c]});
            }
            const int S = split[i];
            if (S <= 1) {
                probs.push_back(Problem{D[i], pshape[i], first_pair, len, 0});
                prefix.push_back(prefix.back() + tiles_of(s));
            } else if (len >= S) {  // cut the chain: part q takes pairs [q*len/S, (q+1)*len/S), every k window
                for (int q = 0; q < S; ++q) {
                    const int p0 = (int)((long long)q * len / S), p1 = (int)((long long)(q + 1) * len / S);
                    probs.push_back(Problem{D[i], pshape[i], first_pair + p0, p1 - p0, 0});
                    prefix.push_back(prefix.back() + tiles_of(s));
                }
            } else {  // fewer pairs than parts: every pair is cut along its k windows
                const int nwin = (s.K + kContractKWin - 1) / kContractKWin;
                const int per_pair = std::min(nwin, std::max(1, S / len));
                for (int c = 0; c < len; ++c)
                    for (int q = 0; q < per_pair; ++q) {
                        const int w0 = (int)((long long)q * nwin / per_pair), w1 = (int)((long long)(q + 1) * nwin / per_pair);
                        probs.push_back(Problem{D[i], pshape[i], first_pair + c, 1, per_pair > 1 ? (w0 | (w1 << 16)) : 0});
                        prefix.push_back(prefix.back() + tiles_of(s));
                    }
            }
        }
        if (probs.empty()) continue;
        auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t b_probs = sizeof(Problem) * probs.size(), b_pairs = sizeof(Pair) * pairs.size(),
                     b_shapes = sizeof(Shape) * shapes.size(), b_prefix = sizeof(int) * prefix.size();
        const size_t off_pairs = up(b_probs), off_shapes = off_pairs + up(b_pairs), off_prefix = off_shapes + up(b_shapes);
        void *h, *d;
        SIP_TRY(desc_alloc(off_prefix + b_prefix, &h, &d));
        memcpy(h, probs.data(), b_probs);
        memcpy((char*)h + off_pairs, pairs.data(), b_pairs);
        memcpy((char*)h + off_shapes, shapes.data(), b_shapes);
        memcpy((char*)h + off_prefix, prefix.data(), b_prefix);
        SIP_TRY(desc_commit(h, d, off_prefix + b_prefix));
        ContractArgs a;
        memset(&a, 0, sizeof(a));
        a.probs = (const Problem*)d;
        a.pairs = (const Pair*)((char*)d + off_pairs);
        a.shapes = (const Shape*)((char*)d + off_shapes);
        a.tile_prefix = (const int*)((char*)d + off_prefix);
        a.nprob = (int)probs.size();
        a.total_tiles = prefix.back();
        a.alpha = alpha;
        a.beta = beta;
        a.atomic = any_split ? 1 : 0;
        {   // lockstep throttle: only for launches whose tiles all walk the same number of ring stages
            long long steps0 = -1;
            bool uniform = true;
            for (const Problem& pr : probs) {
                const Shape& sh = shapes[pr.shape];
                const int nwin = (sh.K + kContractKWin - 1) / kContractKWin;
                const long long per_pair = (long long)(nwin - 1) * (kContractKWin / 16) + (sh.K - (nwin - 1) * kContractKWin + 15) / 16;
                const long long st = pr.kwin ? -2 : per_pair * pr.chain_len;
                if (steps0 == -1) steps0 = st;
                if (st != steps0 || st < 0) { uniform = false; break; }
            }
            if (uniform && steps0 >= 4 * kSyncEvery && steps0 < (1LL << 30) && a.total_tiles >= ctx().num_sms)
                a.sync_q = (int)((steps0 + kSyncEvery - 1) / kSyncEvery);
        }
        SIP_TRY(launch_contract(a, a_kc, b_kc, vec, tile));
    }
    return SIPGPU_OK;
}

namespace {

int labels_to_ptrn(int drank, const int* dlab, int lrank, const int* llab, int rrank, const int* rlab, int* ptrn) {
    int aces[3 + 96], k = 3;
    if (drank < 0 || lrank < 0 || rrank < 0 || drank > 32 || lrank > 32 || rrank > 32) return SIPGPU_E_ARG;
    aces[0] = drank; aces[1] = lrank; aces[2] = rrank;
    for (int i = 0; i < drank; ++i) aces[k++] = dlab[i];
    for (int i = 0; i < lrank; ++i) aces[k++] = llab[i];
    for (int i = 0; i < rrank; ++i) aces[k++] = rlab[i];
    // a pardo body repeats a handful of label triples thousands of times: remember the last few conversions
    struct Memo {
        int n = 0;
        int key[3 + 96], ptrn[96];
    };
    static Memo memo[4];
    static int victim = 0;
    for (const Memo& m : memo)
        if (m.n == k && !memcmp(m.key, aces, sizeof(int) * k)) {
            memcpy(ptrn, m.ptrn, sizeof(int) * (lrank + rrank));
            return SIPGPU_OK;
        }
    const int e = get_contraction_ptrn(drank, lrank, rrank, aces + 3, ptrn);
    if (e) {
        set_error("illegal contraction label pattern (get_contraction_ptrn ierr=%d)", e);
        return SIPGPU_E_PATTERN;
    }
    Memo& m = memo[victim];
    victim = (victim + 1) & 3;
    m.n = k;
    memcpy(m.key, aces, sizeof(int) * k);
    memcpy(m.ptrn, ptrn, sizeof(int) * (lrank + rrank));
    return SIPGPU_OK;
}

// host-pointer staging for boundary 1
struct Staged {
    double* d = nullptr;
    long long n = 0;
    ~Staged() { if (d) pool_free(d); }
    int up(const double* h, long long n_) {
        SIP_TRY(wl_flush());  // host-pointer calls are blocking: drain a pending recording first
        n = n_;
        d = pool_alloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        if (!d) return SIPGPU_E_NOMEM;
        if (h && n > 0) SIP_CUDA(cudaMemcpyAsync(d, h, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx().stream));
        return SIPGPU_OK;
    }
    int down(double* h) {
        SIP_CUDA(cudaMemcpyAsync(h, d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx().stream));
        SIP_CUDA(cudaStreamSynchronize(ctx().stream));
        return SIPGPU_OK;
    }
};

}  // namespace
}  // namespace sipgpu

using namespace sipgpu;

#define B1_TRY(call)                       \
    do {                                   \
        int rc__ = (call);                 \
        if (rc__ != 0) { *ierr = rc__; return; } \
    } while (0)

extern "C" {

// ------------------------------------------------ boundary 1 ------------------------------------------------
long long tensor_size_by_shape_(int* num_dim, int* dims, int* ierr) {
    SIP_TRACE("tensor_size_by_shape_");
    long long sz = 1;
    *ierr = 0;
    if (*num_dim > 0) {
        for (int i = 0; i < *num_dim; ++i) sz *= dims[i];
        if (sz <= 0) *ierr = 2;
    } else if (*num_dim < 0) {
        *ierr = 1;
        sz = 0;
    }
    return sz;
}

void get_contraction_ptrn_(int* drank, int* lrank, int* rrank, int* aces_ptrn, int* my_ptrn, int* ierr) {
    SIP_TRACE("get_contraction_ptrn_");
    *ierr = get_contraction_ptrn(*drank, *lrank, *rrank, aces_ptrn, my_ptrn);
}

void tensor_block_init__(int* nthreads, double* tens, int* rank, int* ext, double* val, int* ierr) {
    SIP_TRACE("tensor_block_init__");
    (void)nthreads;
    *ierr = 0;
    if (*rank < 0) { *ierr = -1; return; }
    B1_TRY(ensure_init());
    const long long n = volume(*rank, ext);
    Staged t;
    B1_TRY(t.up(nullptr, n));
    B1_TRY(ew_fill(t.d, n, *val));
    B1_TRY(t.down(tens));
}

void tensor_block_scale__(int* nthreads, double* tens, int* rank, int* ext, double* fac, int* ierr) {
    SIP_TRACE("tensor_block_scale__");
    (void)nthreads;
    *ierr = 0;
    if (*rank < 0) { *ierr = -1; return; }
    B1_TRY(ensure_init());
    const long long n = volume(*rank, ext);
    Staged t;
    B1_TRY(t.up(tens, n));
    B1_TRY(ew_scale(t.d, n, *fac));
    B1_TRY(t.down(tens));
}

double tensor_block_norm2__(int* nthreads, double* tens, int* rank, int* ext, int* ierr) {
    SIP_TRACE("tensor_block_norm2__");
    (void)nthreads;
    *ierr = 0;
    if (*rank < 0) { *ierr = -1; return 0.0; }
    if ((*ierr = ensure_init()) != 0) return 0.0;
    const long long n = volume(*rank, ext);
    Staged t;
    if ((*ierr = t.up(tens, n)) != 0) return 0.0;
    double r = 0.0;
    *ierr = ew_dot(t.d, t.d, n, &r);
    return r;
}

void tensor_block_slice__(int* nthreads, int* rank, double* tens, int* tens_ext, double* slice, int* slice_ext,
                          int* ext_beg, int* ierr) {
    SIP_TRACE("tensor_block_slice__");
    (void)nthreads;
    *ierr = 0;
    if (*rank < 0) { *ierr = 1; return; }
    B1_TRY(ensure_init());
    Staged t, s;
    B1_TRY(t.up(tens, volume(*rank, tens_ext)));
    B1_TRY(s.up(nullptr, volume(*rank, slice_ext)));
    B1_TRY(ew_slice(*rank, t.d, tens_ext, s.d, slice_ext, ext_beg));
    B1_TRY(s.down(slice));
}

void tensor_block_insert__(int* nthreads, int* rank, double* tens, int* tens_ext, double* slice, int* slice_ext,
                           int* ext_beg, int* ierr) {
    SIP_TRACE("tensor_block_insert__");
    (void)nthreads;
    *ierr = 0;
    if (*rank < 0) { *ierr = 1; return; }
    B1_TRY(ensure_init());
    Staged t, s;
    B1_TRY(t.up(tens, volume(*rank, tens_ext)));
    B1_TRY(s.up(slice, volume(*rank, slice_ext)));
    B1_TRY(ew_insert(*rank, t.d, tens_ext, s.d, slice_ext, ext_beg));
    B1_TRY(t.down(tens));
}

void tensor_block_add__(const int* nthreads, const int* rank, int* ext, double* tens0, double* tens1, const double* fac,
                        int* ierr) {
    SIP_TRACE("tensor_block_add__");
    (void)nthreads;
    *ierr = 0;
    if (*rank < 0) { *ierr = -1; return; }
    B1_TRY(ensure_init());
    const long long n = volume(*rank, ext);
    Staged t0, t1;
    B1_TRY(t0.up(tens0, n));
    B1_TRY(t1.up(tens1, n));
    B1_TRY(ew_axpy(t0.d, t1.d, n, *fac));
    B1_TRY(t0.down(tens0));
}

void tensor_block_copy__(int* nthreads, int* rank, int* ext, int* dim_transp, double* tens_in, double* tens_out,
                         int* ierr) {
    SIP_TRACE("tensor_block_copy__");
    (void)nthreads;
    *ierr = 0;
    if (*rank < 0) { *ierr = *rank; return; }
    B1_TRY(ensure_init());
    const long long n = volume(*rank, ext);
    Staged in, out;
    B1_TRY(in.up(tens_in, n));
    B1_TRY(out.up(nullptr, n));
    B1_TRY(permute_block(*rank, ext, dim_transp, in.d, out.d));
    B1_TRY(out.down(tens_out));
}

void tensor_block_contract__(int* nthreads, int* contr_ptrn, double* ltens, int* lrank, int* lext, double* rtens,
                             int* rrank, int* rext, double* dtens, int* drank, int* dext, int* ierr) {
    SIP_TRACE("tensor_block_contract__");
    (void)nthreads;
    *ierr = 0;
    if (*lrank < 0 || *rrank < 0 || *drank < 0 || *lrank > 32 || *rrank > 32 || *drank > 32) { *ierr = -1; return; }
    B1_TRY(ensure_init());
    if (!contr_ptrn_ok(contr_ptrn, *lrank, *rrank, *drank, lext, rext, dext)) { *ierr = 1; return; }
    Staged l, r, d;
    B1_TRY(l.up(ltens, volume(*lrank, lext)));
    B1_TRY(r.up(rtens, volume(*rrank, rext)));
    B1_TRY(d.up(nullptr, volume(*drank, dext)));
    B1_TRY(contract_device(contr_ptrn, l.d, *lrank, lext, r.d, *rrank, rext, d.d, *drank, dext, 1.0, 0.0));
    B1_TRY(d.down(dtens));
}

// ------------------------------------------------ boundary 2 ------------------------------------------------
int _init_gpu(int* devid, int* my_rank) {
    SIP_TRACE("_init_gpu");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_error("no CUDA device available; libsipgpu has no CPU fallback");
        return SIPGPU_E_NODEVICE;
    }
    const int dev = (my_rank ? *my_rank : 0) % ndev;  // the reference hard-codes rank 0 (interpreter.cpp:84-88)
    SIP_TRY(sipgpu_init(dev));
    if (devid) *devid = dev;
    return SIPGPU_OK;
}
int _finalize_gpu(void) {
    SIP_TRACE("_finalize_gpu"); return sipgpu_finalize(); }
double* _gpu_allocate(const int n) {
    SIP_TRACE("_gpu_allocate"); return sipgpu_block_alloc(n, 1); }
int _gpu_free(double* g) {
    SIP_TRACE("_gpu_free"); return sipgpu_block_free(g); }
int _gpu_host_to_device(double* c_addr, double* g_addr, const int n) {
    SIP_TRACE("_gpu_host_to_device"); return sipgpu_h2d(g_addr, c_addr, n); }
int _gpu_device_to_host(double* c_addr, double* g_addr, const int n) {
    SIP_TRACE("_gpu_device_to_host"); return sipgpu_d2h(c_addr, g_addr, n); }
int _gpu_device_to_device(double* dst, double* src, const int n) {
    SIP_TRACE("_gpu_device_to_device");
    if (n < 0) return SIPGPU_E_ARG;
    if (wl_active()) return wl_rec_ew(WL_SCALE_COPY, dst, src, nullptr, n, 1.0);
    SIP_TRY(ensure_init());
    SIP_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, ctx().stream));
    return SIPGPU_OK;
}
int _gpu_double_memset(double* g, double value, const int n) {
    SIP_TRACE("_gpu_double_memset"); return sipgpu_block_fill(g, n, value); }
int _gpu_selfmultiply(double* x, const double alpha, const int n) {
    SIP_TRACE("_gpu_selfmultiply"); return sipgpu_block_scale(x, n, alpha); }
int _gpu_axpy(double* y, double* x, const double alpha, const int n) {
    SIP_TRACE("_gpu_axpy"); return sipgpu_block_axpy(y, x, n, alpha); }
int _gpu_permute(double* y, const int ny, const int* y_dims, const int* y_inds, double* x, const int nx, const int* x_dims,
                 const int* x_inds) {
    SIP_TRACE("_gpu_permute");
    (void)y_dims;
    if (ny != nx) return SIPGPU_E_ARG;
    return sipgpu_block_permute_labels(nx, x_dims, y_inds, x_inds, x, y);
}
int _gpu_contract(double* y, const int ny, const int* y_dims, const int* y_inds, double* x1, const int n1, const int* x1_dims,
                  const int* x1_inds, double* x2, const int n2, const int* x2_dims, const int* x2_inds) {
    SIP_TRACE("_gpu_contract");
    return sipgpu_block_contract_labels(ny, y_dims, y_inds, y, n1, x1_dims, x1_inds, x1, n2, x2_dims, x2_inds, x2, 1.0, 0.0);
}

// ------------------------------------------------ boundary 3 ------------------------------------------------
// While a work-list recording is open (sipgpu_wl_begin) the asynchronous ops below are recorded instead of launched.
int sipgpu_block_fill(double* d, long long n, double v) {
    SIP_TRACE("sipgpu_block_fill");
    return wl_active() ? wl_rec_ew(WL_FILL, d, nullptr, nullptr, n, v) : ew_fill(d, n, v);
}
int sipgpu_block_scale(double* d, long long n, double f) {
    SIP_TRACE("sipgpu_block_scale");
    return wl_active() ? wl_rec_ew(WL_SCALE, d, nullptr, nullptr, n, f) : ew_scale(d, n, f);
}
int sipgpu_block_scale_and_copy(double* d, const double* s, long long n, double f) {
    SIP_TRACE("sipgpu_block_scale_and_copy");
    if (wl_active()) return s ? wl_rec_ew(WL_SCALE_COPY, d, s, nullptr, n, f) : SIPGPU_E_ARG;
    return ew_scale_copy(d, s, n, f);
}
int sipgpu_block_increment(double* d, long long n, double delta) {
    SIP_TRACE("sipgpu_block_increment");
    return wl_active() ? wl_rec_ew(WL_INCR, d, nullptr, nullptr, n, delta) : ew_increment(d, n, delta);
}
int sipgpu_block_axpy(double* d, const double* s, long long n, double f) {
    SIP_TRACE("sipgpu_block_axpy");
    if (wl_active()) return s ? wl_rec_ew(WL_AXPY, d, s, nullptr, n, f) : SIPGPU_E_ARG;
    return ew_axpy(d, s, n, f);
}
int sipgpu_block_accumulate(double* d, const double* s, long long n) {
    SIP_TRACE("sipgpu_block_accumulate"); return sipgpu_block_axpy(d, s, n, 1.0); }
int sipgpu_block_add_sub(double* d, const double* l, const double* r, long long n, double sign) {
    SIP_TRACE("sipgpu_block_add_sub");
    if (wl_active()) return (l && r) ? wl_rec_ew(WL_ADDSUB, d, l, r, n, sign) : SIPGPU_E_ARG;
    return ew_add_sub(d, l, r, n, sign);
}
int sipgpu_block_fill_hash(double* d, long long n, unsigned long long seed, unsigned long long tag, double scale) {
    SIP_TRACE("sipgpu_block_fill_hash");
    if (wl_active())
        return wl_rec_opaque([=] { return ew_fill_hash(d, n, seed, tag, scale); }, {{d, sizeof(double) * (size_t)n, WL_W}});
    return ew_fill_hash(d, n, seed, tag, scale);
}
int sipgpu_block_dot_accumulate(const double* l, const double* r, long long n, double* d_scalar) {
    SIP_TRACE("sipgpu_block_dot_accumulate");
    if (wl_active())
        return wl_rec_opaque([=] { return ew_dot_device(l, r, n, d_scalar, 1.0); },
                             {{l, sizeof(double) * (size_t)n, WL_R}, {r, sizeof(double) * (size_t)n, WL_R}, {d_scalar, 8, WL_RW}});
    return ew_dot_device(l, r, n, d_scalar, 1.0);
}
int sipgpu_block_norm2(const double* t, long long n, double* out) {
    SIP_TRACE("sipgpu_block_norm2");
    SIP_TRY(wl_flush());
    return ew_dot(t, t, n, out);
}
int sipgpu_block_dot(const double* l, const double* r, long long n, double* out) {
    SIP_TRACE("sipgpu_block_dot");
    SIP_TRY(wl_flush());
    return ew_dot(l, r, n, out);
}
int sipgpu_block_slice(int rank, const double* t, const int* t_ext, double* s, const int* s_ext, const int* beg) {
    SIP_TRACE("sipgpu_block_slice");
    if (wl_active()) {
        if (rank < 0 || rank > kMaxRank || !t || !s || !t_ext || !s_ext || !beg) return SIPGPU_E_ARG;
        std::array<int, kMaxRank> te{}, se{}, bg{};
        for (int i = 0; i < rank; ++i) te[i] = t_ext[i], se[i] = s_ext[i], bg[i] = beg[i];
        return wl_rec_opaque([=] { return ew_slice(rank, t, te.data(), s, se.data(), bg.data()); },
                             {{t, sizeof(double) * (size_t)volume(rank, t_ext), WL_R}, {s, sizeof(double) * (size_t)volume(rank, s_ext), WL_W}});
    }
    return ew_slice(rank, t, t_ext, s, s_ext, beg);
}
int sipgpu_block_insert(int rank, double* t, const int* t_ext, const double* s, const int* s_ext, const int* beg) {
    SIP_TRACE("sipgpu_block_insert");
    if (wl_active()) {
        if (rank < 0 || rank > kMaxRank || !t || !s || !t_ext || !s_ext || !beg) return SIPGPU_E_ARG;
        std::array<int, kMaxRank> te{}, se{}, bg{};
        for (int i = 0; i < rank; ++i) te[i] = t_ext[i], se[i] = s_ext[i], bg[i] = beg[i];
        return wl_rec_opaque([=] { return ew_insert(rank, t, te.data(), s, se.data(), bg.data()); },
                             {{t, sizeof(double) * (size_t)volume(rank, t_ext), WL_RW}, {s, sizeof(double) * (size_t)volume(rank, s_ext), WL_R}});
    }
    return ew_insert(rank, t, t_ext, s, s_ext, beg);
}
int sipgpu_block_permute(int rank, const int* ext, const int* transp, const double* in, double* out) {
    SIP_TRACE("sipgpu_block_permute");
    if (wl_active() && rank >= 1 && rank <= kMaxRank) return wl_rec_permute(rank, ext, transp, in, out, 1.0, 0.0);
    SIP_TRY(wl_flush());
    return permute_block(rank, ext, transp, in, out);
}
int sipgpu_permute_batched(int n, int rank, const int* ext, const int* transp, const double* const* in,
                           double* const* out, double alpha, double beta) {
    SIP_TRACE("sipgpu_permute_batched");
    if (wl_active() && rank >= 1 && rank <= kMaxRank) {
        if (n < 0 || (n && (!in || !out))) return SIPGPU_E_ARG;
        for (int i = 0; i < n; ++i) SIP_TRY(wl_rec_permute(rank, ext, transp, in[i], out[i], alpha, beta));
        return SIPGPU_OK;
    }
    SIP_TRY(wl_flush());
    return permute_batched(n, rank, ext, transp, in, out, alpha, beta);
}
int sipgpu_block_permute_labels(int rank, const int* rhs_ext, const int* lhs_labels, const int* rhs_labels, const double* rhs,
                                double* lhs) {
    SIP_TRACE("sipgpu_block_permute_labels");
    if (rank < 0 || rank > 32) return SIPGPU_E_ARG;
    int transp[33];
    SIP_TRY(permutation_from_labels(rank, lhs_labels, rhs_labels, transp));
    return sipgpu_block_permute(rank, rhs_ext, transp, rhs, lhs);
}
int sipgpu_block_contract(const int* ptrn, const double* L, int lrank, const int* lext, const double* R, int rrank,
                          const int* rext, double* D, int drank, const int* dext, double alpha, double beta) {
    SIP_TRACE("sipgpu_block_contract");
    if (wl_active()) {
        int rc;
        if (lrank >= 1 && rrank >= 1 && drank >= 1 && lrank <= kMaxRank && rrank <= kMaxRank && drank <= kMaxRank) {
            rc = wl_rec_contract(ptrn, L, lrank, lext, R, rrank, rext, D, drank, dext, alpha, beta);
        } else {  // scalar-operand forms (F90:764-780): not batchable, scheduled as they are
            if (!L || !R || !D || lrank < 0 || rrank < 0 || drank < 0 || lrank > kMaxRank || rrank > kMaxRank || drank > kMaxRank)
                return SIPGPU_E_ARG;
            if (!contr_ptrn_ok(ptrn, lrank, rrank, drank, lext, rext, dext)) {
                rc = 1;
            } else {
                std::array<int, 2 * kMaxRank> pt{};
                std::array<int, kMaxRank> le{}, re{}, de{};
                for (int i = 0; i < lrank + rrank; ++i) pt[i] = ptrn[i];
                for (int i = 0; i < lrank; ++i) le[i] = lext[i];
                for (int i = 0; i < rrank; ++i) re[i] = rext[i];
                for (int i = 0; i < drank; ++i) de[i] = dext[i];
                rc = wl_rec_opaque(
                    [=] { return contract_device(pt.data(), L, lrank, le.data(), R, rrank, re.data(), D, drank, de.data(), alpha, beta); },
                    {{L, sizeof(double) * (size_t)volume(lrank, lext), WL_R},
                     {R, sizeof(double) * (size_t)volume(rrank, rext), WL_R},
                     {D, sizeof(double) * (size_t)volume(drank, dext), beta != 0.0 ? WL_RW : WL_W}});
            }
        }
        if (rc == 1) {
            set_error("invalid contraction pattern for these extents (contr_ptrn_ok)");
            return SIPGPU_E_PATTERN;
        }
        return rc;
    }
    const int rc = contract_device(ptrn, L, lrank, lext, R, rrank, rext, D, drank, dext, alpha, beta);
    if (rc == 1) {
        set_error("invalid contraction pattern for these extents (contr_ptrn_ok)");
        return SIPGPU_E_PATTERN;
    }
    return rc;
}
int sipgpu_block_contract_labels(int drank, const int* dext, const int* dlab, double* D, int lrank, const int* lext,
                                 const int* llab, const double* L, int rrank, const int* rext, const int* rlab,
                                 const double* R, double alpha, double beta) {
    SIP_TRACE("sipgpu_block_contract_labels");
    // ranks are validated BEFORE the label conversion writes lrank + rrank pattern entries
    if (drank < 0 || lrank < 0 || rrank < 0 || drank > kMaxRank || lrank > kMaxRank || rrank > kMaxRank) {
        set_error("contract: ranks (%d, %d, %d) outside 0..%d", drank, lrank, rrank, kMaxRank);
        return SIPGPU_E_ARG;
    }
    int ptrn[2 * kMaxRank];
    SIP_TRY(labels_to_ptrn(drank, dlab, lrank, llab, rrank, rlab, ptrn));
    return sipgpu_block_contract(ptrn, L, lrank, lext, R, rrank, rext, D, drank, dext, alpha, beta);
}
int sipgpu_block_contract_sliced(const int* ptrn, const double* L, int lrank, const int* lext, const int* lparent_ext,
                                 const int* lbeg, const double* R, int rrank, const int* rext, const int* rparent_ext,
                                 const int* rbeg, double* D, int drank, const int* dext, const int* dparent_ext,
                                 const int* dbeg, double alpha, double beta) {
    SIP_TRACE("sipgpu_block_contract_sliced");
    if (wl_active()) {
        if (!L || !R || !D || lrank < 1 || rrank < 1 || drank < 1 || lrank > kMaxRank || rrank > kMaxRank || drank > kMaxRank)
            return SIPGPU_E_ARG;
        struct Cap {
            int ptrn[2 * kMaxRank], ext[3][kMaxRank], par[3][kMaxRank], beg[3][kMaxRank];
            bool has_par[3], has_beg[3];
        } c;
        memset(&c, 0, sizeof(c));
        const int ranks[3] = {lrank, rrank, drank};
        const int* exts[3] = {lext, rext, dext};
        const int* pars[3] = {lparent_ext, rparent_ext, dparent_ext};
        const int* begs[3] = {lbeg, rbeg, dbeg};
        size_t bytes[3];
        for (int i = 0; i < lrank + rrank; ++i) c.ptrn[i] = ptrn[i];
        for (int t = 0; t < 3; ++t) {
            c.has_par[t] = pars[t] != nullptr;
            c.has_beg[t] = begs[t] != nullptr;
            for (int i = 0; i < ranks[t]; ++i) {
                c.ext[t][i] = exts[t][i];
                if (pars[t]) c.par[t][i] = pars[t][i];
                if (begs[t]) c.beg[t][i] = begs[t][i];
            }
            bytes[t] = sizeof(double) * (size_t)volume(ranks[t], pars[t] ? pars[t] : exts[t]);  // the whole parent array
        }
        return wl_rec_opaque(
            [=] {
                const int rc = contract_device_sliced(c.ptrn, L, lrank, c.ext[0], c.has_par[0] ? c.par[0] : nullptr,
                                                      c.has_beg[0] ? c.beg[0] : nullptr, R, rrank, c.ext[1],
                                                      c.has_par[1] ? c.par[1] : nullptr, c.has_beg[1] ? c.beg[1] : nullptr, D, drank,
                                                      c.ext[2], c.has_par[2] ? c.par[2] : nullptr, c.has_beg[2] ? c.beg[2] : nullptr,
                                                      alpha, beta);
                return rc == 1 ? SIPGPU_E_PATTERN : rc;
            },
            {{L, bytes[0], WL_R}, {R, bytes[1], WL_R}, {D, bytes[2], (beta != 0.0 || dparent_ext) ? WL_RW : WL_W}});
    }
    const int rc = contract_device_sliced(ptrn, L, lrank, lext, lparent_ext, lbeg, R, rrank, rext, rparent_ext, rbeg, D, drank,
                                          dext, dparent_ext, dbeg, alpha, beta);
    return rc == 1 ? SIPGPU_E_PATTERN : rc;
}
int sipgpu_contract_batched(int n, const int* ptrn, int lrank, int rrank, int drank, const int* lext, const int* rext,
                            const int* dext, const double* const* L, const double* const* R, double* const* D,
                            double alpha, double beta) {
    SIP_TRACE("sipgpu_contract_batched");
    SIP_TRY(wl_flush());  // already a work-list: runs as it is, after whatever was recorded before it
    const int rc = contract_chained(n, ptrn, lrank, rrank, drank, lext, rext, dext, nullptr, L, R, D, alpha, beta);
    return rc == 1 ? SIPGPU_E_PATTERN : rc;
}
int sipgpu_contract_chained(int n, const int* ptrn, int lrank, int rrank, int drank, const int* lext, const int* rext,
                            const int* dext, const int* chain_start, const double* const* L, const double* const* R,
                            double* const* D, double alpha, double beta) {
    SIP_TRACE("sipgpu_contract_chained");
    if (!chain_start) return SIPGPU_E_ARG;
    SIP_TRY(wl_flush());
    const int rc = contract_chained(n, ptrn, lrank, rrank, drank, lext, rext, dext, chain_start, L, R, D, alpha, beta);
    return rc == 1 ? SIPGPU_E_PATTERN : rc;
}

// ---- prepared work-lists: marshal once, launch many times ----
// A CC iteration launches the same work-list every time (same blocks, same chains); sipgpu_contract_chained rebuilds
// shapes, problems, tile prefix sums and uploads them on every call -- for thousands of small blocks that host work is
// longer than the kernel.  A plan keeps the descriptors resident and replays the launches.
}  // extern "C"
struct sipgpu_plan {
    int n = 0, lrank = 0, rrank = 0, drank = 0;
    std::vector<int> ptrn, lext, rext, dext, chain;
    std::vector<const double*> L, R;
    std::vector<double*> D;
    struct Built {
        double alpha, beta;
        sipgpu::Capture cap;
    };
    std::vector<Built*> built;  // one captured launch sequence per (alpha, beta) used so far
    ~sipgpu_plan() {
        for (Built* b : built) {
            for (void* p : b->cap.device) sipgpu::pool_free(p);
            delete b;
        }
    }
};
extern "C" {
int sipgpu_plan_contract_chained(int n, const int* ptrn, int lrank, int rrank, int drank, const int* lext, const int* rext,
                                 const int* dext, const int* chain_start, const double* const* L, const double* const* R,
                                 double* const* D, sipgpu_plan** out) {
    SIP_TRACE("sipgpu_plan_contract_chained");
    if (!out || n < 0 || !ptrn || lrank < 1 || rrank < 1 || drank < 1 || lrank > kMaxRank || rrank > kMaxRank || drank > kMaxRank ||
        (n && (!lext || !rext || !dext || !L || !R || !D)))
        return SIPGPU_E_ARG;
    sipgpu_plan* p = new sipgpu_plan();
    p->n = n; p->lrank = lrank; p->rrank = rrank; p->drank = drank;
    p->ptrn.assign(ptrn, ptrn + lrank + rrank);
    p->lext.assign(lext, lext + (size_t)n * lrank);
    p->rext.assign(rext, rext + (size_t)n * rrank);
    p->dext.assign(dext, dext + (size_t)n * drank);
    const int npairs = chain_start ? chain_start[n] : n;
    if (chain_start) p->chain.assign(chain_start, chain_start + n + 1);
    p->L.assign(L, L + npairs);
    p->R.assign(R, R + npairs);
    p->D.assign(D, D + n);
    *out = p;
    return SIPGPU_OK;
}
int sipgpu_plan_launch(sipgpu_plan* p, double alpha, double beta) {
    SIP_TRACE("sipgpu_plan_launch");
    if (!p) return SIPGPU_E_ARG;
    SIP_TRY(wl_flush());  // runs after whatever was recorded before it, like sipgpu_contract_chained
    if (p->n == 0) return SIPGPU_OK;
    SIP_TRY(ensure_init());
    sipgpu_plan::Built* b = nullptr;
    for (sipgpu_plan::Built* x : p->built)
        if (x->alpha == alpha && x->beta == beta) b = x;
    if (!b) {
        b = new sipgpu_plan::Built{alpha, beta, {}};
        capture() = &b->cap;
        const int rc = contract_chained(p->n, p->ptrn.data(), p->lrank, p->rrank, p->drank, p->lext.data(), p->rext.data(),
                                        p->dext.data(), p->chain.empty() ? nullptr : p->chain.data(), p->L.data(), p->R.data(),
                                        p->D.data(), alpha, beta);
        capture() = nullptr;
        if (rc != SIPGPU_OK && !b->cap.failed) {
            for (void* q : b->cap.device) pool_free(q);
            delete b;
            return rc == 1 ? SIPGPU_E_PATTERN : rc;
        }
        if (b->cap.failed) b->cap.steps.clear();
        p->built.push_back(b);
    }
    if (b->cap.failed) {  // a launch site that cannot be replayed: the eager path, every time
        const int rc = contract_chained(p->n, p->ptrn.data(), p->lrank, p->rrank, p->drank, p->lext.data(), p->rext.data(),
                                        p->dext.data(), p->chain.empty() ? nullptr : p->chain.data(), p->L.data(), p->R.data(),
                                        p->D.data(), alpha, beta);
        return rc == 1 ? SIPGPU_E_PATTERN : rc;
    }
    for (auto& step : b->cap.steps) SIP_TRY(step());
    return SIPGPU_OK;
}
int sipgpu_plan_destroy(sipgpu_plan* p) {
    SIP_TRACE("sipgpu_plan_destroy");
    if (!p) return SIPGPU_OK;
    if (ctx().inited) cudaStreamSynchronize(ctx().stream);  // descriptors may still be read by queued launches
    delete p;
    return SIPGPU_OK;
}

// ---- boundary 3b: elementwise CC super-instructions (superinstr.cu), reference calling convention ----
int sipgpu_set_predefined_int_array(const char* name, int n, const int* values) {
    SIP_TRACE("sipgpu_set_predefined_int_array"); return si_set_int_array(name, n, values); }
#define SI_RETURN(expr)                   \
    do {                                  \
        const int rc__ = (expr);          \
        if (ierr) *ierr = rc__;           \
        return rc__;                      \
    } while (0)
int sipgpu_si_energy_denominator_rhf(int*, int* rank_0, int* index_values_0, int*, int* extents_0, double* data_0, int*, int* rank_1,
                                     int*, int*, int* extents_1, double* data_1, int* ierr) {
    SIP_TRACE("sipgpu_si_energy_denominator_rhf");
    if (!rank_0 || !rank_1) SI_RETURN(SIPGPU_E_ARG);
    if (wl_active()) {  // recorded with its read/write sets; runs at its scheduled position
        const int r0 = *rank_0, r1 = *rank_1;
        if (r0 < 0 || r0 > kMaxRank || r1 < 1 || r1 > 2 || !index_values_0 || !extents_0 || !extents_1 || !data_0 || !data_1)
            SI_RETURN(SIPGPU_E_ARG);
        std::array<int, kMaxRank> iv{}, e0{}, e1{};
        for (int i = 0; i < r0; ++i) iv[i] = index_values_0[i], e0[i] = extents_0[i];
        for (int i = 0; i < r1; ++i) e1[i] = extents_1[i];
        SI_RETURN(wl_rec_opaque([=] { return si_energy_denominator_rhf(r0, iv.data(), e0.data(), data_0, r1, e1.data(), data_1); },
                                {{data_0, sizeof(double) * (size_t)volume(r0, extents_0), WL_RW},
                                 {data_1, sizeof(double) * (size_t)volume(r1, extents_1), WL_R}}));
    }
    SI_RETURN(si_energy_denominator_rhf(*rank_0, index_values_0, extents_0, data_0, *rank_1, extents_1, data_1));
}
int sipgpu_si_energy_ty_denominator_rhf(int*, int* rank_0, int* index_values_0, int*, int* extents_0, double* data_0, int*, int* rank_1,
                                        int*, int*, int* extents_1, double* data_1, int*, int*, int*, int*, int*, double* data_2, int* ierr) {
    SIP_TRACE("sipgpu_si_energy_ty_denominator_rhf");
    if (!rank_0 || !rank_1 || !data_2) SI_RETURN(SIPGPU_E_ARG);
    if (wl_active()) {
        const int r0 = *rank_0, r1 = *rank_1;
        if (r0 != 4 || r1 != 2 || !index_values_0 || !extents_0 || !extents_1 || !data_0 || !data_1) SI_RETURN(SIPGPU_E_ARG);
        std::array<int, kMaxRank> iv{}, e0{}, e1{};
        for (int i = 0; i < r0; ++i) iv[i] = index_values_0[i], e0[i] = extents_0[i];
        for (int i = 0; i < r1; ++i) e1[i] = extents_1[i];
        SI_RETURN(wl_rec_opaque([=] { return si_energy_denominator_rhf(r0, iv.data(), e0.data(), data_0, r1, e1.data(), data_1, data_2); },
                                {{data_0, sizeof(double) * (size_t)volume(r0, extents_0), WL_RW},
                                 {data_1, sizeof(double) * (size_t)volume(r1, extents_1), WL_R}, {data_2, sizeof(double), WL_R}}));
    }
    SI_RETURN(si_energy_denominator_rhf(*rank_0, index_values_0, extents_0, data_0, *rank_1, extents_1, data_1, data_2));
}
int sipgpu_si_stripi(int*, int* rank_0, int* index_values_0, int*, int* extents_0, double* data_0, int*, int* rank_1,
                     int* index_values_1, int*, int* extents_1, double* data_1, int* ierr) {
    SIP_TRACE("sipgpu_si_stripi");
    if (!rank_0 || !rank_1 || *rank_0 != *rank_1) SI_RETURN(SIPGPU_E_ARG);
    if (wl_active()) {
        const int r = *rank_0;
        if (r < 1 || r > kMaxRank || !index_values_0 || !index_values_1 || !extents_0 || !extents_1 || !data_0 || !data_1)
            SI_RETURN(SIPGPU_E_ARG);
        std::array<int, kMaxRank> iv0{}, iv1{}, e0{}, e1{};
        for (int i = 0; i < r; ++i) iv0[i] = index_values_0[i], iv1[i] = index_values_1[i], e0[i] = extents_0[i], e1[i] = extents_1[i];
        SI_RETURN(wl_rec_opaque([=] { return si_stripi(r, iv0.data(), e0.data(), data_0, iv1.data(), e1.data(), data_1); },
                                {{data_0, sizeof(double) * (size_t)volume(r, extents_0), WL_R},
                                 {data_1, sizeof(double) * (size_t)volume(r, extents_1), WL_RW}}));
    }
    SI_RETURN(si_stripi(*rank_0, index_values_0, extents_0, data_0, index_values_1, extents_1, data_1));
}
int sipgpu_si_anti_symm_o(int*, int* rank_0, int* index_values_0, int*, int* extents_0, double* data_0, int* ierr) {
    SIP_TRACE("sipgpu_si_anti_symm_o");
    if (!rank_0) SI_RETURN(SIPGPU_E_ARG);
    if (wl_active()) {
        const int r = *rank_0;
        if (r < 1 || r > kMaxRank || !index_values_0 || !extents_0 || !data_0) SI_RETURN(SIPGPU_E_ARG);
        std::array<int, kMaxRank> iv{}, e0{};
        for (int i = 0; i < r; ++i) iv[i] = index_values_0[i], e0[i] = extents_0[i];
        SI_RETURN(wl_rec_opaque([=] { return si_anti_symm_o(r, iv.data(), e0.data(), data_0); },
                                {{data_0, sizeof(double) * (size_t)volume(r, extents_0), WL_RW}}));
    }
    SI_RETURN(si_anti_symm_o(*rank_0, index_values_0, extents_0, data_0));
}
int sipgpu_si_anti_symm_v(int*, int* rank_0, int* index_values_0, int*, int* extents_0, double* data_0, int* ierr) {
    SIP_TRACE("sipgpu_si_anti_symm_v");
    if (!rank_0) SI_RETURN(SIPGPU_E_ARG);
    if (wl_active()) {
        const int r = *rank_0;
        if (r < 1 || r > kMaxRank || !index_values_0 || !extents_0 || !data_0) SI_RETURN(SIPGPU_E_ARG);
        std::array<int, kMaxRank> iv{}, e0{};
        for (int i = 0; i < r; ++i) iv[i] = index_values_0[i], e0[i] = extents_0[i];
        SI_RETURN(wl_rec_opaque([=] { return si_anti_symm_v(r, iv.data(), e0.data(), data_0); },
                                {{data_0, sizeof(double) * (size_t)volume(r, extents_0), WL_RW}}));
    }
    SI_RETURN(si_anti_symm_v(*rank_0, index_values_0, extents_0, data_0));
}
int sipgpu_si_return_sval(int*, int* rank_0, int*, int*, int* extents_0, double* data_0, int*, int* rank_1, int*, int*, int*,
                          double* data_1, int* ierr) {
    SIP_TRACE("sipgpu_si_return_sval");
    if (!rank_0 || !rank_1 || *rank_1 != 0) SI_RETURN(SIPGPU_E_ARG);
    if (wl_active()) {
        const int r = *rank_0;
        if (r < 1 || r > kMaxRank || !extents_0 || !data_0 || !data_1) SI_RETURN(SIPGPU_E_ARG);
        std::array<int, kMaxRank> e0{};
        for (int i = 0; i < r; ++i) e0[i] = extents_0[i];
        SI_RETURN(wl_rec_opaque([=] { return si_return_sval(r, e0.data(), data_0, data_1); },
                                {{data_0, sizeof(double) * (size_t)volume(r, extents_0), WL_R}, {data_1, sizeof(double), WL_W}}));
    }
    SI_RETURN(si_return_sval(*rank_0, extents_0, data_0, data_1));
}
int sipgpu_si_invert_diagonal(int*, int* rank_0, int*, int*, int* extents_0, double* data_0, int*, int* rank_1, int*, int*, int*,
                              double* data_1, int* ierr) {
    SIP_TRACE("sipgpu_si_invert_diagonal");
    if (!rank_0 || !rank_1) SI_RETURN(SIPGPU_E_ARG);
    if (wl_active()) {
        const int r0 = *rank_0, r1 = *rank_1;
        if (r0 < 1 || r0 > kMaxRank || !extents_0 || !data_0 || !data_1) SI_RETURN(SIPGPU_E_ARG);
        std::array<int, kMaxRank> e0{};
        for (int i = 0; i < r0; ++i) e0[i] = extents_0[i];
        const size_t nb = sizeof(double) * (size_t)volume(r0, extents_0);
        SI_RETURN(wl_rec_opaque([=] { return si_invert_diagonal(r0, r1, e0.data(), data_0, data_1); },
                                {{data_0, nb, WL_RW}, {data_1, nb, WL_R}}));
    }
    SI_RETURN(si_invert_diagonal(*rank_0, *rank_1, extents_0, data_0, data_1));
}

int sipgpu_si_invert_diagonal_asym(int*, int* rank_0, int* index_values_0, int*, int* extents_0, double* data_0, int*, int* rank_1, int*,
                                   int*, int*, double* data_1, int* ierr) {
    SIP_TRACE("sipgpu_si_invert_diagonal_asym");
    if (!rank_0 || !rank_1) SI_RETURN(SIPGPU_E_ARG);
    if (wl_active()) {
        const int r0 = *rank_0, r1 = *rank_1;
        if (r0 < 1 || r0 > kMaxRank || !index_values_0 || !extents_0 || !data_0 || !data_1) SI_RETURN(SIPGPU_E_ARG);
        std::array<int, kMaxRank> iv{}, e0{};
        for (int i = 0; i < r0; ++i) iv[i] = index_values_0[i], e0[i] = extents_0[i];
        const size_t nb = sizeof(double) * (size_t)volume(r0, extents_0);
        SI_RETURN(wl_rec_opaque([=] { return si_invert_diagonal_asym(r0, r1, iv.data(), e0.data(), data_0, data_1); },
                                {{data_0, nb, WL_RW}, {data_1, nb, WL_R}}));
    }
    SI_RETURN(si_invert_diagonal_asym(*rank_0, *rank_1, index_values_0, extents_0, data_0, data_1));
}
int sipgpu_si_return_diagonal_elements(int*, int* rank_0, int*, int*, int* extents_0, double* data_0, int* ierr) {
    SIP_TRACE("sipgpu_si_return_diagonal_elements");
    if (!rank_0) SI_RETURN(SIPGPU_E_ARG);
    if (wl_active()) {
        const int r = *rank_0;
        if (r < 1 || r > kMaxRank || !extents_0 || !data_0) SI_RETURN(SIPGPU_E_ARG);
        std::array<int, kMaxRank> e0{};
        for (int i = 0; i < r; ++i) e0[i] = extents_0[i];
        SI_RETURN(wl_rec_opaque([=] { return si_return_diagonal_elements(r, e0.data(), data_0); },
                                {{data_0, sizeof(double) * (size_t)volume(r, extents_0), WL_RW}}));
    }
    SI_RETURN(si_return_diagonal_elements(*rank_0, extents_0, data_0));
}

int sipgpu_dgemm_tn(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
                    double* C, int ldc) {
    SIP_TRACE("sipgpu_dgemm_tn");
    SIP_TRY(wl_flush());
    SIP_TRY(ensure_init());
    if (m < 1 || n < 1 || k < 1 || lda < k || ldb < k || ldc < m || !A || !B || !C) return SIPGPU_E_ARG;
    ContractArgs a;
    memset(&a, 0, sizeof(a));
    Shape& s = a.s0;
    s.M = m; s.N = n; s.K = k;
    s.nm = s.nn = s.nk = 1;
    s.mext[0] = m; s.msL[0] = lda; s.msD[0] = 1;
    s.next[0] = n; s.nsR[0] = ldb; s.nsD[0] = ldc;
    s.kext[0] = k; s.ksL[0] = 1; s.ksR[0] = 1;
    s.a_kc = s.b_kc = 1;
    a.nprob = 1;
    a.total_tiles = tiles_of(s);
    a.alpha = alpha;
    a.beta = beta;
    a.pair0.L = A; a.pair0.R = B; a.p0.D = C; a.p0.chain_len = 1;
    s.vec = (k % 2 == 0 && lda % 2 == 0 && ldb % 2 == 0 && (((uintptr_t)A | (uintptr_t)B) & 15) == 0) ? 1 : 0;
    s.tile = contract_pick_tile(m, n);
    if (contract_tile_count(m, n, s.tile) < ctx().num_sms) s.tile = kSmallTile;
    a.total_tiles = tiles_of(s);
    return launch_contract(a, true, true, s.vec != 0, s.tile);
}

// ---- host-only planner views (no device needed): used by the CPU tests of the host logic ----
int sipgpu_debug_contract_shape(const int* ptrn, int lrank, const int* lext, int rrank, const int* rext, int drank,
                                const int* dext, int* shape_ints /* sizeof(Shape)/4 */) {
    SIP_TRACE("sipgpu_debug_contract_shape");
    Shape s;
    const int rc = build_shape(ptrn, lrank, lext, rrank, rext, drank, dext, &s);
    if (rc == 0) memcpy(shape_ints, &s, sizeof(s));
    return rc;
}
int sipgpu_debug_shape_ints(void) {
    SIP_TRACE("sipgpu_debug_shape_ints"); return (int)(sizeof(Shape) / sizeof(int)); }
int sipgpu_debug_permute_plan(int rank, const int* ext, const int* transp, long long* meta, int* rtab, int* wtab, int cap) {
    SIP_TRACE("sipgpu_debug_permute_plan");
    return sipgpu::permute_plan_debug(rank, ext, transp, meta, rtab, wtab, cap);
}

int sipgpu_set_tuning(const char* key, double value) {
    SIP_TRACE("sipgpu_set_tuning");
    if (!key) return SIPGPU_E_ARG;
    if (!strcmp(key, "lowint_max_intensity")) { lowint_set_max_intensity(value); wl_tuning_changed(); return SIPGPU_OK; }
    if (!strcmp(key, "permute_bulk")) { permute_set_bulk((int)value); wl_tuning_changed(); return SIPGPU_OK; }
    if (!strcmp(key, "copy_bulk")) { wl_set_copy_bulk((int)value); wl_tuning_changed(); return SIPGPU_OK; }
    if (!strcmp(key, "permute_vec")) { permute_set_vec((int)value); wl_tuning_changed(); return SIPGPU_OK; }
    if (!strcmp(key, "lowint_slab")) { lowint_set_slab((int)value); wl_tuning_changed(); return SIPGPU_OK; }
    if (!strcmp(key, "lowint_scope")) { lowint_set_scope((int)value); wl_tuning_changed(); return SIPGPU_OK; }
    set_error("sipgpu_set_tuning: unknown key '%s'", key);
    return SIPGPU_E_ARG;
}
int sipgpu_dmma_peak_probe(int iters, double* tflops_out) {
    SIP_TRACE("sipgpu_dmma_peak_probe");
    SIP_TRY(ensure_init());
    return dmma_probe(iters, tflops_out);
}

int sipgpu_copy_bw_probe(size_t bytes, int reps, double* gbs_out) {
    SIP_TRACE("sipgpu_copy_bw_probe");
    SIP_TRY(ensure_init());
    Ctx& c = ctx();
    const long long n = (long long)(bytes / 16) * 2;
    double *a = nullptr, *b = nullptr;
    SIP_CUDA(cudaMalloc(&a, n * 8));
    SIP_CUDA(cudaMalloc(&b, n * 8));
    SIP_CUDA(cudaMemsetAsync(a, 0, n * 8, c.stream));
    cudaEvent_t e0, e1;
    SIP_CUDA(cudaEventCreate(&e0));
    SIP_CUDA(cudaEventCreate(&e1));
    ew_copy_probe(b, a, n);
    SIP_CUDA(cudaEventRecord(e0, c.stream));
    for (int r = 0; r < reps; ++r) ew_copy_probe(b, a, n);
    SIP_CUDA(cudaEventRecord(e1, c.stream));
    SIP_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    SIP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *gbs_out = 2.0 * n * 8 * reps / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    return SIPGPU_OK;
}

}  // extern "C"

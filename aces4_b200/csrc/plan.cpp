// plan.cpp -- host planner (see plan.h).  Pure integer work on <= 18 labels; runs on the calling thread.
#include "plan.h"

#include <stdlib.h>

#include <algorithm>

namespace sipgpu {

int get_contraction_ptrn(int drank, int lrank, int rrank, const int* aces, int* my_ptrn) {
    if (drank < 0 || lrank < 0 || rrank < 0) return 1;
    const int m = lrank + rrank, n = drank + m;
    if (n % 2 != 0) return 2;
    if (m <= 0) return drank > 0 ? 3 : 0;
    if (n > 3 * 32) return 1;
    int pos[96];
    for (int i = 0; i < n; ++i) pos[i] = i;
    // stable order by label: equal labels keep D < L < R position order (what the reference's merge sort yields)
    std::stable_sort(pos, pos + n, [&](int a, int b) { return aces[a] < aces[b]; });
    for (int i = 0; i + 1 < n; i += 2) {  // every label exactly twice (F90:106-116)
        if (aces[pos[i]] != aces[pos[i + 1]]) return 4;
        if (i > 0 && aces[pos[i]] == aces[pos[i - 1]]) return 5;
    }
    for (int i = 0; i + 1 < n; i += 2) {  // F90:118-128
        const int j = pos[i] + 1, k = pos[i + 1] + 1;  // 1-based, j < k
        if (j <= drank && k > drank) {
            my_ptrn[k - drank - 1] = j;  // free index: operand position -> destination position
        } else if (j > drank && j <= drank + lrank && k > drank + lrank) {
            my_ptrn[j - drank - 1] = -(k - (drank + lrank));  // contracted: L position <-> R position
            my_ptrn[k - drank - 1] = -(j - drank);
        } else {
            return 6;
        }
    }
    return 0;
}

bool contr_ptrn_ok(const int* p, int lr, int rr, int dr, const int* lext, const int* rext, const int* dext) {
    int seen[3 * 32] = {0};
    if (dr + lr + rr > 96) return false;
    for (int j = 1; j <= lr; ++j) {
        const int q = p[j - 1];
        if (q > 0 && q <= dr) {
            seen[q - 1]++, seen[dr + j - 1]++;
            if (lext[j - 1] != dext[q - 1]) return false;
        } else if (q < 0 && -q <= rr) {
            seen[dr + lr - q - 1]++;
            if (lext[j - 1] != rext[-q - 1]) return false;
        } else {
            return false;
        }
    }
    for (int j = 1; j <= rr; ++j) {
        const int q = p[lr + j - 1];
        if (q > 0 && q <= dr) {
            seen[q - 1]++, seen[dr + lr + j - 1]++;
            if (rext[j - 1] != dext[q - 1]) return false;
        } else if (q < 0 && -q <= lr) {
            seen[dr - q - 1]++;
            if (rext[j - 1] != lext[-q - 1]) return false;
            if (p[-q - 1] != -j) return false;  // the two halves of a contracted pair must name each other
        } else {
            return false;
        }
    }
    for (int i = 0; i < dr + lr + rr; ++i)
        if (seen[i] != 1) return false;
    return true;
}

namespace {
struct Dim {
    int ext, s0, s1;
};

// drop extent-1 dims, then merge neighbours that are contiguous in both tensors
int tidy(Dim* d, int n) {
    int w = 0;
    for (int i = 0; i < n; ++i)
        if (d[i].ext != 1) d[w++] = d[i];
    n = w;
    w = 0;
    for (int i = 0; i < n; ++i) {
        if (w > 0 && (long long)d[w - 1].s0 * d[w - 1].ext == d[i].s0 && (long long)d[w - 1].s1 * d[w - 1].ext == d[i].s1)
            d[w - 1].ext *= d[i].ext;
        else
            d[w++] = d[i];
    }
    return w;
}
}  // namespace

int build_shape(const int* ptrn, int lrank, const int* lext, int rrank, const int* rext, int drank, const int* dext,
                Shape* out) {
    return build_shape_strided(ptrn, lrank, lext, nullptr, rrank, rext, nullptr, drank, dext, nullptr, out);
}

// As build_shape, but every operand may be a SLICE of a larger dense array: *par are the extents of the parent array
// (nullptr: the block is dense).  Strides come from the parent, extents from the slice; the caller offsets the base
// pointer to the slice's first element.  This is how `T[a,i,mu,j] = T2[a,i,b,j] * ca[mu,b]` reads the block of the
// static array `ca` in place instead of extracting it first (contiguous_array_manager.cpp:202-230, F90:271-330).
int build_shape_strided(const int* ptrn, int lrank, const int* lext, const int* lpar, int rrank, const int* rext,
                        const int* rpar, int drank, const int* dext, const int* dpar, Shape* out) {
    if (lrank < 1 || rrank < 1 || drank < 1 || lrank > 32 || rrank > 32 || drank > 32) return SIPGPU_E_ARG;
    if (!contr_ptrn_ok(ptrn, lrank, rrank, drank, lext, rext, dext)) return 1;
    long long sL[32], sR[32], sD[32], s = 1;
    bool even_l = true, even_r = true;  // all non-unit strides even (16-byte fetches stay aligned row to row)
    for (int i = 0; i < lrank; ++i) {
        if (lpar && lpar[i] < lext[i]) return SIPGPU_E_ARG;
        sL[i] = s; s *= lpar ? lpar[i] : lext[i];
        if (sL[i] != 1 && (sL[i] & 1)) even_l = false;
    }
    if (s <= 0 || s >= (1LL << 31)) return SIPGPU_E_ARG;
    s = 1;
    for (int i = 0; i < rrank; ++i) {
        if (rpar && rpar[i] < rext[i]) return SIPGPU_E_ARG;
        sR[i] = s; s *= rpar ? rpar[i] : rext[i];
        if (sR[i] != 1 && (sR[i] & 1)) even_r = false;
    }
    if (s <= 0 || s >= (1LL << 31)) return SIPGPU_E_ARG;
    s = 1;
    for (int i = 0; i < drank; ++i) {
        if (dpar && dpar[i] < dext[i]) return SIPGPU_E_ARG;
        sD[i] = s; s *= dpar ? dpar[i] : dext[i];
    }
    if (s <= 0 || s >= (1LL << 31)) return SIPGPU_E_ARG;

    Dim md[32], nd[32], kd[32];
    int nm = 0, nn = 0, nk = 0;
    long long M = 1, N = 1, K = 1;
    for (int j = 0; j < lrank; ++j)
        if (ptrn[j] > 0) { md[nm++] = {lext[j], (int)sL[j], (int)sD[ptrn[j] - 1]}; M *= lext[j]; }
    for (int j = 0; j < rrank; ++j) {
        const int q = ptrn[lrank + j];
        if (q > 0) { nd[nn++] = {rext[j], (int)sR[j], (int)sD[q - 1]}; N *= rext[j]; }
        else { kd[nk++] = {rext[j], (int)sL[-q - 1], (int)sR[j]}; K *= rext[j]; }
    }
    // Order of the contracted group is free (any order consistent between L and R gives the same sums up to
    // rounding).  Put the stride-1 dimension of an operand first so that its K-runs are contiguous: R's if R's
    // fastest (non-unit) dimension is contracted, else L's; when BOTH operands have a contracted fastest dimension and
    // their orders disagree (D[a,b] = L[c,d,b,a]*R[d,c]: a matrix-vector product whose big operand would be read with a
    // stride of one segment), the order of the larger operand wins.  m and n groups already ascend in L / R stride.
    auto min_stride = [](const Dim* d, int n, bool second) {
        int best = INT32_MAX;
        for (int i = 0; i < n; ++i)
            if (d[i].ext != 1) best = std::min(best, second ? d[i].s1 : d[i].s0);
        return best;
    };
    const int kminL = min_stride(kd, nk, false), kminR = min_stride(kd, nk, true);
    const int mminL = min_stride(md, nm, false), nminR = min_stride(nd, nn, false);
    const bool r_k_fast = kminR < nminR, l_k_fast = kminL < mminL;
    static const bool by_size = [] { const char* e = getenv("SIPGPU_KORDER_BY_SIZE"); return !e || atoi(e) != 0; }();
    if (l_k_fast && (!r_k_fast || (by_size && M > N)))
        std::stable_sort(kd, kd + nk, [](const Dim& a, const Dim& b) { return a.s0 < b.s0; });
    nm = tidy(md, nm);
    nn = tidy(nd, nn);
    nk = tidy(kd, nk);
    if (nm > kMaxRank || nn > kMaxRank || nk > kMaxRank) return SIPGPU_E_ARG;
    // Destination-friendly order of the free group that holds D's fastest index: if that index is not the group's
    // first (i.e. the operand's order and the destination's order differ -- a transposing contraction), move it to
    // the front.  Consecutive tile rows then fall on consecutive destination elements, so the epilogue writes whole
    // 32-byte sectors instead of scattering single doubles over sectors that other tiles complete later (measured:
    // D[a,b,c,d] = L[b,a,d,e]*R[c,e], o = 20, v = 50: 0.72 -> 2.7 TB/s algorithmic); the price is that this operand is gathered with 8-byte
    // instead of 16-byte cp.async items (its sectors are still fully used inside the tile).
    {
        static const bool enabled = [] { const char* e = getenv("SIPGPU_DFAST_ORDER"); return !e || atoi(e) != 0; }();
        int best = INT32_MAX, grp = -1, idx = -1;
        for (int i = 0; i < nm; ++i) if (md[i].s1 < best) { best = md[i].s1; grp = 0; idx = i; }
        for (int i = 0; i < nn; ++i) if (nd[i].s1 < best) { best = nd[i].s1; grp = 1; idx = i; }
        // only when the destination is at least as large as the operand that gets gathered (N >= K for L, M >= K for
        // R): for a small rank-2 result of a big block (D[a,i] = L[a,i,b,j]*R[b,j]) the reads are what matters
        const long long other = grp == 0 ? N : M;
        if (enabled && idx > 0 && other >= K) {
            Dim* g = grp == 0 ? md : nd;
            if (g[idx].ext >= 4) std::rotate(g, g + idx, g + idx + 1);
        }
    }

    Shape sh;
    memset(&sh, 0, sizeof(sh));
    sh.M = (int)M; sh.N = (int)N; sh.K = (int)K;
    sh.nm = nm; sh.nn = nn; sh.nk = nk;
    for (int i = 0; i < nm; ++i) { sh.mext[i] = md[i].ext; sh.msL[i] = md[i].s0; sh.msD[i] = md[i].s1; }
    for (int i = 0; i < nn; ++i) { sh.next[i] = nd[i].ext; sh.nsR[i] = nd[i].s0; sh.nsD[i] = nd[i].s1; }
    for (int i = 0; i < nk; ++i) { sh.kext[i] = kd[i].ext; sh.ksL[i] = kd[i].s0; sh.ksR[i] = kd[i].s1; }
    sh.a_kc = (nk > 0 && sh.ksL[0] == 1) ? 1 : 0;
    sh.b_kc = (nk > 0 && sh.ksR[0] == 1) ? 1 : 0;
    // 16-byte fetches need the contiguous direction of each operand to be a unit-stride dimension of even extent:
    // every other stride of that operand is then a multiple of it (even), so element pairs never straddle a
    // dimension boundary and stay 16-byte aligned (given a 16-byte aligned block).
    const bool a_vec = sh.a_kc ? (sh.kext[0] % 2 == 0) : (nm > 0 && sh.msL[0] == 1 && sh.mext[0] % 2 == 0);
    const bool b_vec = sh.b_kc ? (sh.kext[0] % 2 == 0) : (nn > 0 && sh.nsR[0] == 1 && sh.next[0] % 2 == 0);
    sh.vec = (a_vec && b_vec && (!lpar || even_l) && (!rpar || even_r)) ? 1 : 0;
    *out = sh;
    return 0;
}

Shape swap_operands(const Shape& s) {
    Shape o = s;
    o.M = s.N; o.N = s.M;
    o.nm = s.nn; o.nn = s.nm;
    for (int i = 0; i < kMaxRank; ++i) {
        o.mext[i] = s.next[i]; o.msL[i] = s.nsR[i]; o.msD[i] = s.nsD[i];
        o.next[i] = s.mext[i]; o.nsR[i] = s.msL[i]; o.nsD[i] = s.msD[i];
        o.ksL[i] = s.ksR[i];   o.ksR[i] = s.ksL[i];
    }
    o.a_kc = s.b_kc; o.b_kc = s.a_kc;
    o.swapped = !s.swapped;
    return o;
}

int permutation_from_labels(int rank, const int* lhs_labels, const int* rhs_labels, int* transp) {
    transp[0] = 1;
    for (int i = 0; i < rank; ++i) {
        int j = 0;
        while (j < rank && rhs_labels[j] != lhs_labels[i]) ++j;
        if (j >= rank) return SIPGPU_E_PATTERN;  // "illegal transpose" (interpreter.cpp:2049-2084)
        transp[j + 1] = i + 1;
    }
    // each rhs dim must have received exactly one destination
    int cnt[32] = {0};
    for (int j = 0; j < rank; ++j) {
        if (transp[j + 1] < 1 || transp[j + 1] > rank) return SIPGPU_E_PATTERN;
        if (cnt[transp[j + 1] - 1]++) return SIPGPU_E_PATTERN;
    }
    return 0;
}

int build_perm_shape(int rank, const int* ext, const int* transp, PermShape* out) {
    if (rank < 1 || rank > 32) return SIPGPU_E_ARG;
    int oext[32];
    bool used[32] = {false};
    for (int i = 0; i < rank; ++i) {
        const int np = transp[i + 1];
        if (np < 1 || np > rank || used[np - 1] || ext[i] <= 0) return SIPGPU_E_ARG;
        used[np - 1] = true;
        oext[np - 1] = ext[i];
    }
    long long ostr[32], s = 1;
    for (int i = 0; i < rank; ++i) { ostr[i] = s; s *= oext[i]; }
    if (s >= (1LL << 31)) return SIPGPU_E_ARG;
    Dim d[32];
    long long is = 1;
    for (int i = 0; i < rank; ++i) { d[i] = {ext[i], (int)is, (int)ostr[transp[i + 1] - 1]}; is *= ext[i]; }
    int n = tidy(d, rank);
    if (n > kMaxRank) return SIPGPU_E_ARG;
    PermShape p;
    memset(&p, 0, sizeof(p));
    p.rank = n;
    p.total = s;
    for (int i = 0; i < n; ++i) { p.ext[i] = d[i].ext; p.in_stride[i] = d[i].s0; p.out_stride[i] = d[i].s1; }
    *out = p;
    return 0;
}

}  // namespace sipgpu

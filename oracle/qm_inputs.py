"""TEST INFRASTRUCTURE (like everything under oracle/): inputs for an energy-level parity check of the LCCD hot path.

The reference's energy goldens (test/test_qm.cpp:388-468: water / 3-21G, drop_mo=1-1: scf -75.58432674274046,
lccd_correlation -0.12610179886435) need integrals, an SCF and an integral transformation that the reference computes
with Fortran engines (ERD/OED, src/sip/super_instructions/qm/**) that cannot be built in this image.
Those stages are NOT on the hot path and are not rebuilt in the product.  This file computes the same quantities with
numpy, from the geometry and basis tables of the reference's own `.dat` setup file (decoded into
tests/golden/water_321g_setup.json by scripts/make_water_setup_golden.py), so that the block-contraction path can be
driven with the real amplitudes equations of src/sialx/qm/cc/rlccd_rhf.sialx and its result compared with the
reference's golden energies:

  * McMurchie-Davidson integrals over contracted s/p/d Gaussians (cartesian, or real solid harmonics for the d shells of a
    setup with intspherical = 1 -- test/ccsdpt_test.dat; overlap, kinetic, nuclear attraction,
    electron repulsion), Boys function by series / asymptotic expansion + recursion;
  * closed-shell RHF with DIIS from the core-Hamiltonian guess (what scf_rhf_coreh.sialx converges to);
  * the Mulliken-ordered MO integral classes of src/sialx/qm/utility/tran_rhf_no4v.sialx:494-523
    (Vpiqj[p,i,q,j] = (pi|qj), VSpipi = Vpiqj - its i<->j exchange, Vaaii[a,a1,i,i1] = (a a1|i i1),
    Viaai[i,a,a1,i1] = (ia|a1 i1), Vaaai[a,a1,a2,i] = (a a1|a2 i)).

Pinned by: nn_repulsion of the .dat (9.361611480180377) and the reference's scf_energy golden to 1e-10
(tests/test_lccd_water_energy_cpu.py).  Nothing in aces4_b200/ imports this module.
"""
import itertools
import math

import numpy as np


# ------------------------------------------------------------------------------------------------------------------
# basis: the shell tables of a .dat (setup_reader.cpp:486-612 names) -> primitive cartesian functions + contraction
# ------------------------------------------------------------------------------------------------------------------
_CARTESIAN = {0: [(0, 0, 0)], 1: [(1, 0, 0), (0, 1, 0), (0, 0, 1)],
              2: [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)]}
_SOLID_D = [{(1, 1, 0): 1.0}, {(0, 1, 1): 1.0}, {(1, 0, 1): 1.0}, {(2, 0, 0): 1.0, (0, 2, 0): -1.0},
            {(0, 0, 2): 2.0, (2, 0, 0): -1.0, (0, 2, 0): -1.0}]


def basis_from_setup(setup):
    """-> dict(center[nprim,3], alpha[nprim], lmn[nprim,3], W[nprim, nao]): AO = sum_p W[p, ao] * g_p, the g_p being
    UNNORMALISED cartesian Gaussians x^l y^m z^n exp(-alpha r^2); W carries coefficient x primitive norm, and every
    contracted function is scaled to unit self-overlap afterwards (the energies do not depend on that scale)."""
    ia, fa = setup["int_arrays"], setup["arrays"]
    spherical = bool(setup["ints"].get("intspherical", 0))
    coords = np.array(fa["coords"][1]).reshape(-1, 3)          # coords(xyz, atom), column-major
    alphas, pcoeffs = np.array(fa["alphas"][1]), np.array(fa["pcoeffs"][1])
    nshells = len(ia["ivangmom"][1])
    if "atom" not in ia and len(coords) != 1:
        raise ValueError("setup without a shell -> atom table and more than one centre")
    prim, cols = [], []                                        # prim: (center, alpha, (l,m,n)); cols: list of {prim index: w}
    for s in range(nshells):
        L, ncf, npf = ia["ivangmom"][1][s], ia["ncfps"][1][s], ia["npfps"][1][s]
        if L > 2:
            raise ValueError("only s, p and d shells are implemented")
        A = coords[ia["atom"][1][s] - 1] if "atom" in ia else coords[0]      # older setups of one atom carry no table
        a0, c0 = ia["ixalphas"][1][s] - 1, ia["ixpcoeffs"][1][s] - 1
        comps = _CARTESIAN[L]
        # functions of the shell as combinations of its cartesian components: the components themselves, or (d shells of a
        # setup with intspherical = 1) the five real solid harmonics.  Only the SPAN matters for the energies.
        funcs = _SOLID_D if (L == 2 and spherical) else [{lmn: 1.0} for lmn in comps]
        base = {}
        for p in range(npf):
            for lmn in comps:
                base[p, lmn] = len(prim)
                prim.append((A, alphas[a0 + p], lmn))
        for c in range(ncf):
            for f in funcs:
                col = {}
                for p in range(npf):
                    coef = pcoeffs[c0 + c * npf + p]           # pcoeffs(prim, contracted) per shell, column-major
                    if coef != 0.0:
                        al = alphas[a0 + p]
                        norm = (2.0 * al / math.pi) ** 0.75 * (2.0 * math.sqrt(al)) ** L
                        for lmn, w in f.items():
                            col[base[p, lmn]] = coef * norm * w
                cols.append(col)
    W = np.zeros((len(prim), len(cols)))
    for j, col in enumerate(cols):
        for p, w in col.items():
            W[p, j] = w
    return {"center": np.array([p[0] for p in prim]), "alpha": np.array([p[1] for p in prim]),
            "lmn": np.array([p[2] for p in prim], dtype=int), "W": W,
            "charge": np.array(fa["charge"][1]), "coords": coords}


# ------------------------------------------------------------------------------------------------------------------
# Boys function F_n(T), n = 0..nmax
# ------------------------------------------------------------------------------------------------------------------
def boys(nmax, T):
    T = np.asarray(T, dtype=float)
    out = np.empty((nmax + 1,) + T.shape)
    small = T < 35.0
    Ts = np.where(small, T, 0.0)
    # series for the highest order, downward recursion for the rest (stable for every T)
    term = np.full(T.shape, 1.0 / (2 * nmax + 1))
    acc = term.copy()
    for k in range(1, 400):
        term = term * (2.0 * Ts) / (2 * nmax + 2 * k + 1)
        acc += term
        if np.all(term <= 1e-18 * acc):
            break
    e = np.exp(-Ts)
    f = acc * e
    lo = np.empty_like(out)
    lo[nmax] = f
    for n in range(nmax, 0, -1):
        lo[n - 1] = (2.0 * Ts * lo[n] + e) / (2 * n - 1)
    # large T: F_0 = sqrt(pi/T)/2 (the erf is 1 to 1e-16), upward recursion (stable for large T)
    Tl = np.where(small, 1.0, T)
    hi = np.empty_like(out)
    hi[0] = 0.5 * np.sqrt(math.pi / Tl)
    el = np.exp(-Tl)
    for n in range(nmax):
        hi[n + 1] = ((2 * n + 1) * hi[n] - el) / (2.0 * Tl)
    for n in range(nmax + 1):
        out[n] = np.where(small, lo[n], hi[n])
    return out


# ------------------------------------------------------------------------------------------------------------------
# McMurchie-Davidson machinery on grids of primitive pairs
# ------------------------------------------------------------------------------------------------------------------
def _hermite_E(imax, jmax, a, b, XAB):
    """E[i, j] = list over t (0..i+j) of the 1-D Hermite expansion coefficients of x_A^i x_B^j exp(-a x_A^2 - b x_B^2)"""
    p = a + b
    XPA, XPB = -b / p * XAB, a / p * XAB
    E = {(0, 0): [np.exp(-a * b / p * XAB * XAB)]}
    for i in range(imax + 1):
        for j in range(jmax + 1):
            if i == 0 and j == 0:
                continue
            prev, X = (E[i - 1, 0], XPA) if j == 0 else (E[i, j - 1], XPB)
            n = i + j
            new = []
            for t in range(n + 1):
                v = np.zeros_like(p)
                if t >= 1:
                    v = v + prev[t - 1] / (2.0 * p)
                if t <= n - 1:
                    v = v + X * prev[t]
                if t + 1 <= n - 1:
                    v = v + (t + 1) * prev[t + 1]
                new.append(v)
            E[i, j] = new
    return E


def _select(E, li, lj, dj, tmax):
    """per-pair selection: out[t][i, j] = E[li[i], lj[j] + dj][t][i, j] (zero where t exceeds the order or lj+dj < 0)"""
    out = [np.zeros_like(E[0, 0][0]) for _ in range(tmax + 1)]
    for i in range(int(li.max()) + 1):
        for j in range(int(lj.max()) + 1):
            if j + dj < 0:
                continue
            mask = (li[:, None] == i) & (lj[None, :] == j)
            if not mask.any():
                continue
            for t, e in enumerate(E[i, j + dj]):
                if t <= tmax:
                    out[t] = out[t] + np.where(mask, e, 0.0)
    return out


def _hermite_R(tmax, alpha, X, Y, Z):
    """R[t,u,v] (order-0 auxiliary Hermite Coulomb integrals) for t+u+v <= tmax on a grid"""
    F = boys(tmax, alpha * (X * X + Y * Y + Z * Z))
    base = [(-2.0 * alpha) ** n * F[n] for n in range(tmax + 1)]
    memo = {}

    def R(t, u, v, n):
        if t < 0 or u < 0 or v < 0:
            return 0.0
        key = (t, u, v, n)
        if key not in memo:
            if t == 0 and u == 0 and v == 0:
                memo[key] = base[n]
            elif t > 0:
                memo[key] = (t - 1) * R(t - 2, u, v, n + 1) + X * R(t - 1, u, v, n + 1)
            elif u > 0:
                memo[key] = (u - 1) * R(t, u - 2, v, n + 1) + Y * R(t, u - 1, v, n + 1)
            else:
                memo[key] = (v - 1) * R(t, u, v - 2, n + 1) + Z * R(t, u, v - 1, n + 1)
        return memo[key]

    return {(t, u, v): R(t, u, v, 0) for t in range(tmax + 1) for u in range(tmax + 1 - t)
            for v in range(tmax + 1 - t - u)}


def ao_integrals(basis):
    """-> S, T, V (nao x nao) and the AO repulsion integrals eri[mu,nu,la,si] = (mu nu|la si), Mulliken order"""
    A, al, lmn, W = basis["center"], basis["alpha"], basis["lmn"], basis["W"]
    n = len(al)
    a, b = al[:, None] * np.ones((1, n)), np.ones((n, 1)) * al[None, :]
    p = a + b
    P = (a[..., None] * A[:, None, :] + b[..., None] * A[None, :, :]) / p[..., None]
    lmax = int(lmn.max())                                                         # 1 for s/p sets, 2 with d shells
    E1 = [_hermite_E(lmax, lmax + 2, a, b, A[:, None, d] - A[None, :, d]) for d in range(3)]
    Es = [_select(E1[d], lmn[:, d], lmn[:, d], 0, 2 * lmax) for d in range(3)]    # [dir][t] -> (n, n)
    # ---- overlap and kinetic energy (1-D factors) ----
    s1 = [Es[d][0] * np.sqrt(math.pi / p) for d in range(3)]
    k1 = []
    for d in range(3):
        lj = lmn[:, d][None, :] * np.ones((n, 1))
        up = _select(E1[d], lmn[:, d], lmn[:, d], 2, 0)[0] * np.sqrt(math.pi / p)
        dn = _select(E1[d], lmn[:, d], lmn[:, d], -2, 0)[0] * np.sqrt(math.pi / p)   # zero where l <= 1
        k1.append(-0.5 * (lj * (lj - 1) * dn - 2.0 * b * (2 * lj + 1) * s1[d] + 4.0 * b * b * up))
    S = s1[0] * s1[1] * s1[2]
    T = k1[0] * s1[1] * s1[2] + s1[0] * k1[1] * s1[2] + s1[0] * s1[1] * k1[2]
    # ---- nuclear attraction ----
    o = 2 * lmax                                                                  # Hermite order of a primitive pair
    tuv = [(t, u, v) for t in range(o + 1) for u in range(o + 1 - t) for v in range(o + 1 - t - u)]
    Epair = {k: Es[0][k[0]] * Es[1][k[1]] * Es[2][k[2]] for k in tuv}
    V = np.zeros((n, n))
    for Z, C in zip(basis["charge"], basis["coords"]):
        R = _hermite_R(o, p, P[..., 0] - C[0], P[..., 1] - C[1], P[..., 2] - C[2])
        V -= Z * 2.0 * math.pi / p * sum(Epair[k] * R[k] for k in tuv)
    # ---- electron repulsion on the (pair, pair) grid ----
    pf, Pf = p.ravel(), P.reshape(-1, 3)
    Ef = {k: Epair[k].ravel() for k in tuv}
    pp, qq = pf[:, None], pf[None, :]
    alpha = pp * qq / (pp + qq)
    R = _hermite_R(2 * o, alpha, Pf[:, None, 0] - Pf[None, :, 0], Pf[:, None, 1] - Pf[None, :, 1],
                   Pf[:, None, 2] - Pf[None, :, 2])
    G = np.zeros((n * n, n * n))
    for k1_ in tuv:
        inner = np.zeros((n * n, n * n))
        for k2 in tuv:
            sign = -1.0 if (k2[0] + k2[1] + k2[2]) % 2 else 1.0
            inner += sign * Ef[k2][None, :] * R[k1_[0] + k2[0], k1_[1] + k2[1], k1_[2] + k2[2]]
        G += Ef[k1_][:, None] * inner
    G *= 2.0 * math.pi ** 2.5 / (pp * qq * np.sqrt(pp + qq))
    G = G.reshape(n, n, n, n)
    # ---- contraction; unit self-overlap of every contracted function ----
    S, T, V = (W.T @ M @ W for M in (S, T, V))
    eri = np.einsum("pqrs,pi,qj,rk,sl->ijkl", G, W, W, W, W, optimize=True)
    d = 1.0 / np.sqrt(np.diag(S))
    S, T, V = (M * d[:, None] * d[None, :] for M in (S, T, V))
    eri = eri * d[:, None, None, None] * d[None, :, None, None] * d[None, None, :, None] * d[None, None, None, :]
    return S, T, V, eri


def dipole_integrals(basis, origin=(0.0, 0.0, 0.0)):
    """-> D[3, nao, nao]: <mu| r_d - origin_d |nu> over the contracted, normalised functions of ao_integrals (the reference's
    `compute_dipole_integrals` super-instruction: OED package; used by the response / transition-moment parts of its programs).
    McMurchie-Davidson: the 1-D moment factor is (E_1 + (P_d - C_d) E_0) sqrt(pi / p)."""
    A, al, lmn, W = basis["center"], basis["alpha"], basis["lmn"], basis["W"]
    n = len(al)
    a, b = al[:, None] * np.ones((1, n)), np.ones((n, 1)) * al[None, :]
    p = a + b
    P = (a[..., None] * A[:, None, :] + b[..., None] * A[None, :, :]) / p[..., None]
    lmax = int(lmn.max())
    E1 = [_hermite_E(lmax, lmax, a, b, A[:, None, d] - A[None, :, d]) for d in range(3)]
    Es = [_select(E1[d], lmn[:, d], lmn[:, d], 0, 1) for d in range(3)]
    s1 = [Es[d][0] * np.sqrt(math.pi / p) for d in range(3)]
    m1 = [(Es[d][1] + (P[..., d] - origin[d]) * Es[d][0]) * np.sqrt(math.pi / p) for d in range(3)]
    S = W.T @ (s1[0] * s1[1] * s1[2]) @ W
    dn = 1.0 / np.sqrt(np.diag(S))
    out = []
    for d in range(3):
        f = [m1[k] if k == d else s1[k] for k in range(3)]
        out.append((W.T @ (f[0] * f[1] * f[2]) @ W) * dn[:, None] * dn[None, :])
    return np.array(out)


def nuclear_repulsion(basis):
    Z, X = basis["charge"], basis["coords"]
    return sum(Z[i] * Z[j] / np.linalg.norm(X[i] - X[j]) for i, j in itertools.combinations(range(len(Z)), 2))


# ------------------------------------------------------------------------------------------------------------------
# closed-shell RHF (the fixed point scf_rhf_coreh.sialx converges to)
# ------------------------------------------------------------------------------------------------------------------
def rhf(S, H, eri, nocc, e_nuc, conv=1e-12, max_iter=200, diis_space=8):
    s, U = np.linalg.eigh(S)
    X = U @ np.diag(s ** -0.5) @ U.T

    def fock_eig(F):
        e, Cp = np.linalg.eigh(X.T @ F @ X)
        return e, X @ Cp

    eps, C = fock_eig(H)
    errs, focks = [], []
    e_old = 0.0
    for it in range(max_iter):
        D = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        F = H + np.einsum("mnls,ls->mn", eri, D) - 0.5 * np.einsum("mlns,ls->mn", eri, D)
        e_el = 0.5 * np.sum(D * (H + F))
        err = X.T @ (F @ D @ S - S @ D @ F) @ X
        errs.append(err)
        focks.append(F)
        errs, focks = errs[-diis_space:], focks[-diis_space:]
        if np.max(np.abs(err)) < conv and abs(e_el - e_old) < conv:
            break
        e_old = e_el
        if len(errs) > 1:
            m = len(errs)
            B = -np.ones((m + 1, m + 1))
            B[m, m] = 0.0
            for i in range(m):
                for j in range(m):
                    B[i, j] = np.sum(errs[i] * errs[j])
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            c = np.linalg.lstsq(B, rhs, rcond=None)[0][:m]
            F = sum(ci * Fi for ci, Fi in zip(c, focks))
        eps, C = fock_eig(F)
    else:
        raise RuntimeError("SCF did not converge")
    eps, C = fock_eig(F)           # canonical orbitals of the converged Fock matrix
    return e_el + e_nuc, eps, C, it + 1


# ------------------------------------------------------------------------------------------------------------------
# MO integral classes with the index order of tran_rhf_no4v.sialx (all Mulliken: V[w,x,y,z] = (wx|yz))
# ------------------------------------------------------------------------------------------------------------------
def mo_classes(eri, C, occ, virt):
    """occ / virt: slices of the ACTIVE occupied and virtual orbitals.  p = active occupied followed by virtual."""
    Co, Cv = C[:, occ], C[:, virt]
    Cp = np.hstack([Co, Cv])
    tr = lambda c1, c2, c3, c4: np.einsum("mnls,mw,nx,ly,sz->wxyz", eri, c1, c2, c3, c4, optimize=True)  # noqa: E731
    vpiqj = tr(Cp, Co, Cp, Co)
    return {"vpiqj": vpiqj,                              # (p i|q j)              tran_rhf_no4v.sialx:494
            "vspipi": vpiqj - vpiqj.transpose(0, 3, 2, 1),   # (p i|q j) - (p j|q i)  :482-501
            "vaaii": tr(Cv, Cv, Co, Co),                 # (a a1|i i1)            :516
            "viaai": tr(Co, Cv, Cv, Co),                 # (i a|a1 i1)            :523
            "vaaai": tr(Cv, Cv, Cv, Co)}                 # (a a1|a2 i)            :444-470 + the last quarter transformation


def cis_singlets(eps_occ, eps_virt, vpiqj_vovo, vaaii, nroots):
    """Singlet CIS of a closed-shell reference (what the reference's rcis_rhf.sialx iterates to; its converged vectors C1_a are
    the starting guesses of the EOM-CCSD program): A[ai,bj] = d_ab d_ij (e_a - e_i) + 2 (ai|bj) - (ab|ij), lowest `nroots`
    eigenpairs.  vpiqj_vovo = (a i|b j) [a,i,b,j], vaaii = (a b|i j) [a,b,i,j].  -> (energies [nroots], vectors [nroots,a,i],
    each normalised to 1 over (a, i))"""
    nv, no = len(eps_virt), len(eps_occ)
    A = 2.0 * vpiqj_vovo - vaaii.transpose(0, 2, 1, 3)
    A = A.reshape(nv * no, nv * no).copy()
    A[np.diag_indices(nv * no)] += (eps_virt[:, None] - eps_occ[None, :]).reshape(-1)
    w, v = np.linalg.eigh(0.5 * (A + A.T))
    vec = v[:, :nroots].T.reshape(nroots, nv, no)
    for k in range(nroots):          # a definite sign: the largest component positive
        if vec[k].reshape(-1)[np.argmax(np.abs(vec[k]))] < 0:
            vec[k] = -vec[k]
    return w[:nroots], vec


def split_blocks(dense, seg_lists):
    """dense ndarray -> {1-based segment tuple: Fortran-ordered block}"""
    offs = [np.concatenate([[0], np.cumsum(s)]) for s in seg_lists]
    out = {}
    for idx in np.ndindex(*[len(s) for s in seg_lists]):
        sl = tuple(slice(offs[d][i], offs[d][i + 1]) for d, i in enumerate(idx))
        out[tuple(i + 1 for i in idx)] = np.asfortranarray(dense[sl])
    return out


def join_blocks(blocks, seg_lists):
    offs = [np.concatenate([[0], np.cumsum(s)]) for s in seg_lists]
    full = np.zeros([o[-1] for o in offs])
    for idx, b in blocks.items():
        sl = tuple(slice(offs[d][i - 1], offs[d][i]) for d, i in enumerate(idx))
        full[sl] = b
    return full

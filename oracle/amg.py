"""CPU restatement of the preconditioner path (rows a16/a17 + "next" row f1 of SURVEY.md section 8).
TEST INFRASTRUCTURE ONLY -- the product never imports this file.

Reference: keo_regularized::rebuild builds  P = K + blockdiag(2x2)  (src/keo_regularized.cpp:181-264)
and keo_regularized::apply applies ONE V-cycle of a MueLu smoothed-aggregation hierarchy built with
"number of equations" = 2 and "reuse: type" = "full" (src/keo_regularized.cpp:88-165,290-336).
MueLu is third-party code that is NOT in the reference tree (unpinned version, CMakeLists.txt:18) and
no reference test pins a V-cycle result or an iteration count  =>  PARITY UNPINNED for this row.  What
is restated here is the published smoothed-aggregation algorithm (Vanek/Mandel/Brezina 1996) with
MueLu's default choices where they are documented (2 dofs per node, tentative prolongator from the
per-dof constant null space, damping 4/3 / lambda_max(D^-1 A) from 10 power iterations, Chebyshev
smoothing with eigenvalue ratio 20, direct coarse solve), with the l1-Jacobi scaling S = absolute row sums
in the smoother, for which lambda_max(S^-1 A) <= 1 is a true bound (MueLu: point diagonal and 1.1 x the
power estimate, which undershoots lambda_max by ~10% on these meshes and lets the V-cycle go indefinite on
large jittered grids) and with a DETERMINISTIC aggregation (MIS-2 by hashed priority) so that the GPU build and this file produce the same
hierarchy.  The same algorithm is implemented on the GPU in nosh_b200/csrc/amg.cu; tests compare the two.

Everything works on the real 2N x 2N matrix (interleaved re/im, the reference's layout); a "node" is a
vertex = one 2x2 block row.
"""
import numpy as np
import scipy.sparse as sp

from .meshgen import splitmix64

IN, UNDECIDED, OUT = 2, 1, 0
POWER_ITS = 10
SA_DAMPING = 4.0 / 3.0
CHEB_RATIO = 20.0


def priority(n, level):
    """30-bit hashed priority of node v on `level` (ties broken by the node id)."""
    with np.errstate(over="ignore"):
        k = np.arange(n, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(level + 1)
    return (splitmix64(k) >> np.uint64(34)).astype(np.int64)


def aggregate_mis2(G, level):
    """Deterministic MIS-2 aggregation of the node graph G (scipy CSR pattern, diagonal ignored).

    Roots: the lexicographically first maximal independent set of G^2 with respect to the order
    (priority desc, id desc) -- what rounds of "local maxima among the undecided nodes become roots"
    produce on the GPU, computed here by the sequential greedy sweep.  Phase 2: the neighbours of a
    root join it.  Phase 3: a remaining node joins the aggregate of its assigned neighbour with the
    largest (priority, id).  Aggregates are numbered by ascending root id."""
    n = G.shape[0]
    indptr, indices = G.indptr, G.indices
    pr = priority(n, level)
    key = (pr << 32) | np.arange(n, dtype=np.int64)
    order = np.argsort(-key, kind="stable")
    state = np.full(n, UNDECIDED, np.int8)
    for v in order:
        if state[v] != UNDECIDED:
            continue
        state[v] = IN
        for u in indices[indptr[v]:indptr[v + 1]]:
            if state[u] == UNDECIDED:
                state[u] = OUT
            for w in indices[indptr[u]:indptr[u + 1]]:
                if state[w] == UNDECIDED:
                    state[w] = OUT
    roots = np.flatnonzero(state == IN)
    agg = np.full(n, -1, np.int64)
    agg[roots] = np.arange(roots.size)
    for r in roots:
        for u in indices[indptr[r]:indptr[r + 1]]:
            if u != r:
                agg[u] = agg[r]
    phase2 = agg.copy()
    for v in np.flatnonzero(agg < 0):
        best, bk = -1, -1
        for u in indices[indptr[v]:indptr[v + 1]]:
            if phase2[u] >= 0 and key[u] > bk:
                bk, best = key[u], phase2[u]
        assert best >= 0
        agg[v] = best
    return agg, roots.size


def aggregate_mis2_rounds(G, level):
    """The same aggregation computed the way the GPU does it (nosh_b200/csrc/amg.cu): synchronous rounds in
    which every undecided node that holds the maximum (state, priority, id) tuple of its distance-2
    neighbourhood becomes a root and every node that sees a root within distance 2 drops out.  Used by the
    CPU tests to show that the rounds and the sequential greedy sweep give identical aggregates."""
    n = G.shape[0]
    Gs = sp.csr_matrix(G)
    indptr, indices = Gs.indptr, Gs.indices
    rows = np.repeat(np.arange(n), np.diff(indptr))
    pr = priority(n, level)
    state = np.full(n, UNDECIDED, np.int64)

    def nbr_max(T):
        out = T.copy()
        np.maximum.at(out, rows, T[indices])
        return out

    ids = np.arange(n, dtype=np.int64)
    low = ((pr << 32) | ids).astype(np.uint64)
    while (state == UNDECIDED).any():
        T0 = (state.astype(np.uint64) << np.uint64(62)) | low
        T2 = nbr_max(nbr_max(T0))
        und = state == UNDECIDED
        new_in = und & (T2 == T0)
        new_out = und & ~new_in & ((T2 >> np.uint64(62)) == IN)
        state[new_in] = IN
        state[new_out] = OUT
    roots = np.flatnonzero(state == IN)
    rootnum = np.full(n, -1, np.int64)
    rootnum[roots] = np.arange(roots.size)
    agg2 = rootnum.copy()
    is_root_nbr = (state[indices] == IN)
    agg2[rows[is_root_nbr]] = np.where(state[rows[is_root_nbr]] == IN, agg2[rows[is_root_nbr]],
                                       rootnum[indices[is_root_nbr]])
    key = (pr << 32) | ids
    agg = agg2.copy()
    for v in np.flatnonzero(agg2 < 0):
        nb = indices[indptr[v]:indptr[v + 1]]
        nb = nb[agg2[nb] >= 0]
        agg[v] = agg2[nb[np.argmax(key[nb])]]
    return agg, roots.size


def start_vector(n2):
    """deterministic start vector of the power iteration: U(-1,1) hashed from the row index"""
    z = splitmix64(np.arange(n2, dtype=np.uint64) + np.uint64(0x51ED270B7A2F3C15))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def lambda_max(A, dinv, its=POWER_ITS):
    """Rayleigh-quotient estimate of lambda_max(D^-1 A) after `its` power iterations."""
    x = start_vector(A.shape[0])
    x /= np.sqrt(x @ x)
    lam = 0.0
    for _ in range(its):
        y = dinv * (A @ x)
        lam = x @ y
        x = y / np.sqrt(y @ y)
    return lam


def l1_scaling(A):
    """1 / absolute row sums: the l1-Jacobi smoother scaling S^-1; lambda_max(S^-1 A) <= 1 (Gershgorin)"""
    return 1.0 / np.asarray(abs(A).sum(axis=1)).ravel()


def node_pattern(A):
    """node-level (block) pattern of a real 2n x 2m matrix as an int CSR"""
    A = A.tocoo()
    n, m = A.shape[0] // 2, A.shape[1] // 2
    G = sp.csr_matrix((np.ones(A.nnz, np.int64), (A.row // 2, A.col // 2)), shape=(n, m))
    G.sum_duplicates()
    G.data[:] = 1
    return G


class Level:
    pass


class Hierarchy:
    """Smoothed-aggregation hierarchy for the real 2N x 2N SPD matrix A (scipy CSR)."""

    def __init__(self, A, G=None, coarse_max=512, max_levels=10, degree=1, coarse_degree=2):
        self.degree = degree                # Chebyshev degree on the finest level
        self.coarse_degree = coarse_degree  # ... on the coarse levels
        self.levels = []
        A = sp.csr_matrix(A)
        G = node_pattern(A) if G is None else G
        s = np.ones(A.shape[0] // 2)
        for lev in range(max_levels):
            L = Level()
            L.A, L.G, L.n = A, G, A.shape[0] // 2
            L.diag = A.diagonal()
            L.dinv = 1.0 / L.diag
            L.sinv = l1_scaling(A)
            self.levels.append(L)
            if L.n <= coarse_max or lev == max_levels - 1:
                break
            L.agg, nc = aggregate_mis2(G, lev)
            if nc >= L.n:
                break
            L.lam = lambda_max(A, L.dinv)
            ssum = np.zeros(nc)
            np.add.at(ssum, L.agg, s * s)
            sc = np.sqrt(ssum)
            p = s / sc[L.agg]
            L.p0 = p
            rows = np.arange(2 * L.n)
            cols = 2 * np.repeat(L.agg, 2) + np.tile([0, 1], L.n)
            P0 = sp.csr_matrix((np.repeat(p, 2), (rows, cols)), shape=(2 * L.n, 2 * nc))
            omega = SA_DAMPING / L.lam
            L.omega = omega
            L.P = sp.csr_matrix(P0 - sp.diags(omega * L.dinv) @ (A @ P0))
            Ac = sp.csr_matrix(L.P.T @ A @ L.P)
            P0n = sp.csr_matrix((np.ones(L.n, np.int64), (np.arange(L.n), L.agg)), shape=(L.n, nc))
            Pn = sp.csr_matrix(G @ P0n + P0n)
            Gc = sp.csr_matrix(Pn.T @ G @ Pn)
            Gc.data[:] = 1
            A, G, s = Ac, Gc, sc
        Lc = self.levels[-1]
        Lc.dense = A.toarray()
        Lc.inv = np.linalg.inv(Lc.dense)
        Lc.inv = 0.5 * (Lc.inv + Lc.inv.T)

    def update_fine(self, A):
        """"reuse: type" = "full" (src/keo_regularized.cpp:300): keep aggregates, prolongators and coarse
        operators; the finest level follows the new matrix (values and diagonal)."""
        L = self.levels[0]
        L.A = sp.csr_matrix(A)
        L.diag = L.A.diagonal()
        L.dinv = 1.0 / L.diag
        L.sinv = l1_scaling(L.A)
        if len(self.levels) == 1:
            L.dense = L.A.toarray()
            L.inv = np.linalg.inv(L.dense)
            L.inv = 0.5 * (L.inv + L.inv.T)

    # Chebyshev smoothing (Ifpack2-style three-term recurrence) on S^-1 A, spectrum in [1/ratio, 1]
    def _cheb(self, L, b, x, zero_start):
        lmax = 1.0
        lmin = lmax / CHEB_RATIO
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta
        rho = 1.0 / sigma
        r = b if zero_start else b - L.A @ x
        d = (L.sinv * r) / theta
        x = d.copy() if zero_start else x + d
        for _ in range(1, self.degree if L is self.levels[0] else self.coarse_degree):
            rho_new = 1.0 / (2.0 * sigma - rho)
            r = b - L.A @ x
            d = (rho_new * rho) * d + (2.0 * rho_new / delta) * (L.sinv * r)
            x = x + d
            rho = rho_new
        return x

    def vcycle(self, b, lev=0):
        L = self.levels[lev]
        if lev == len(self.levels) - 1:
            return L.inv @ b
        x = self._cheb(L, b, None, True)
        r = b - L.A @ x
        xc = self.vcycle(L.P.T @ r, lev + 1)
        x = x + L.P @ xc
        return self._cheb(L, b, x, False)

    def complexity(self):
        return sum(L.A.nnz for L in self.levels) / self.levels[0].A.nnz


def sym_ortho(a, b):
    absA, absB = abs(a), abs(b)
    sgn = lambda t: 1.0 if t >= 0.0 else -1.0
    if absB == 0.0:
        return (1.0 if absA == 0.0 else sgn(a)), 0.0, absA
    if absA == 0.0:
        return 0.0, sgn(b), absB
    if absB >= absA:
        tau = a / b
        s = sgn(b) / np.sqrt(1.0 + tau * tau)
        return s * tau, s, b / s
    tau = b / a
    c = sgn(a) / np.sqrt(1.0 + tau * tau)
    return c, c * tau, a / c


def pminres(apply_A, apply_M, b, tol, maxit):
    """Preconditioned MINRES organised as Belos::MinresIter::iterate() (see nosh_oracle.cpp:minres,
    which is the M = I case of this function).  Returns x, iterations, relres, history."""
    n = b.size
    x = np.zeros(n)
    r1 = b.copy()
    r2 = b.copy()
    y = apply_M(r2)
    beta1 = r1 @ y
    hist = [1.0]
    if beta1 <= 0.0:
        return x, 0, 0.0, hist
    beta1 = np.sqrt(beta1)
    oldBeta, beta, dbar, epsln, phibar, cs, sn = 0.0, beta1, 0.0, 0.0, beta1, -1.0, 0.0
    w = np.zeros(n)
    w2 = np.zeros(n)
    it = 0
    while it < maxit and not (phibar / beta1 <= tol):
        it += 1
        v = y * (1.0 / beta)
        y = apply_A(v)
        if it > 1:
            y = y - (beta / oldBeta) * r1
        alpha = v @ y
        y = y - (alpha / beta) * r2
        r1 = r2
        r2 = y
        y = apply_M(r2)
        oldBeta = beta
        beta = r2 @ y
        if beta < 0.0:
            break
        beta = np.sqrt(beta)
        oldeps = epsln
        delta = cs * dbar + sn * alpha
        gbar = sn * dbar - cs * alpha
        epsln = sn * beta
        dbar = -cs * beta
        cs, sn, gamma = sym_ortho(gbar, beta)
        phi = cs * phibar
        phibar = sn * phibar
        if gamma == 0.0:
            break
        w1 = w2
        w2 = w
        w = (v - oldeps * w1 - delta * w2) * (1.0 / gamma)
        x = x + phi * w
        hist.append(phibar / beta1)
    return x, it, phibar / beta1, hist


def pcg(apply_A, apply_M, b, tol, maxit):
    """Preconditioned CG (Belos::PseudoBlockCGIter with a left preconditioner): stops on
    ||r||_2 / ||r0||_2 <= tol."""
    x = np.zeros(b.size)
    r = b.copy()
    r0 = np.sqrt(r @ r)
    hist = [1.0]
    if r0 == 0.0:
        return x, 0, 0.0, hist
    z = apply_M(r)
    p = z.copy()
    rho = r @ z
    it = 0
    rn = r0
    while it < maxit and not (rn / r0 <= tol):
        it += 1
        Ap = apply_A(p)
        al = rho / (p @ Ap)
        x = x + al * p
        r = r - al * Ap
        rn = np.sqrt(r @ r)
        hist.append(rn / r0)
        z = apply_M(r)
        rho_new = r @ z
        p = z + (rho_new / rho) * p
        rho = rho_new
    return x, it, rn / r0, hist

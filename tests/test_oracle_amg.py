"""CPU checks of the preconditioner oracle (oracle/amg.py): the restated smoothed-aggregation V-cycle is a
symmetric positive definite operator, accelerates MINRES, and its two formulations of the aggregation (the
sequential greedy sweep and the synchronous rounds the GPU runs) agree.  MueLu itself is not in the reference
tree -- parity for this row is unpinned (see the header of oracle/amg.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from oracle import amg, meshgen


def regularised_problem(n, mu=0.1, g=1.0, state="ones"):
    coords, cells = meshgen.tetgrid(n)
    psi, A = meshgen.plain_gl_fields(coords)
    P = oracle.OracleProblem(coords, cells, ("explicit", A))
    x = psi if state == "ones" else meshgen.random_state(P.N)
    P.keo_fill(mu)
    P.jac_rebuild(g, x)
    N = P.N
    K = sp.csr_matrix((P.vals.copy(), P.cols, P.rowptr), shape=(2 * N, 2 * N))
    r = np.arange(N)
    D = sp.csr_matrix((np.concatenate([P.d0[0::2], P.d0[1::2], P.d1b, P.d1b]),
                       (np.concatenate([2 * r, 2 * r + 1, 2 * r, 2 * r + 1]),
                        np.concatenate([2 * r, 2 * r + 1, 2 * r + 1, 2 * r]))), shape=(2 * N, 2 * N))
    J = (K + D).tocsr()
    Pm = sp.csr_matrix((P.keoreg_fill(mu, g, x), P.cols, P.rowptr), shape=(2 * N, 2 * N))
    return P, J, Pm, x


@pytest.mark.parametrize("level", [0, 1])
def test_mis2_rounds_equal_greedy(level):
    P, J, Pm, _ = regularised_problem(9)
    G = amg.node_pattern(Pm)
    a, na = amg.aggregate_mis2(G, level)
    b, nb = amg.aggregate_mis2_rounds(G, level)
    assert na == nb and np.array_equal(a, b)
    assert a.min() == 0 and a.max() == na - 1 and np.unique(a).size == na
    # roots are pairwise more than two edges apart: no node sees two aggregates' roots
    sizes = np.bincount(a)
    assert sizes.min() >= 1 and sizes.mean() > 8


def test_vcycle_is_spd_and_accelerates_minres():
    P, J, Pm, x = regularised_problem(12, state="random")
    H = amg.Hierarchy(Pm, coarse_max=64)
    assert len(H.levels) >= 2
    n2 = Pm.shape[0]
    rng = np.random.default_rng(0)
    u, v = rng.standard_normal(n2), rng.standard_normal(n2)
    Mu, Mv = H.vcycle(u), H.vcycle(v)
    assert abs(v @ Mu - u @ Mv) <= 1e-12 * abs(v @ Mu)      # symmetric
    assert u @ Mu > 0 and v @ Mv > 0                         # positive
    # M approximates the inverse of the regularised KEO
    assert np.linalg.norm(Pm @ Mu - u) < 0.9 * np.linalg.norm(u)
    b = -P.compute_f(1.0, x)
    x0, it0, _ = P.krylov(b, 1e-10, 2000)
    x1, it1, rr, hist = amg.pminres(lambda t: J @ t, H.vcycle, b, 1e-10, 2000)
    assert it1 < it0 / 2
    assert np.linalg.norm(J @ x1 - b) <= 1e-8 * np.linalg.norm(b)
    # with M = I the Python restatement is the C++ oracle's MINRES
    x2, it2, _, _ = amg.pminres(lambda t: J @ t, lambda r: r.copy(), b, 1e-10, 2000)
    assert it2 == it0 and np.abs(x2 - x0).max() <= 1e-9 * np.abs(x0).max()


def test_pcg_on_the_regularised_keo():
    P, J, Pm, x = regularised_problem(10)
    H = amg.Hierarchy(Pm, coarse_max=64, degree=2)
    b = np.sin(np.arange(Pm.shape[0]) * 0.37)
    x1, it1, rr, _ = amg.pcg(lambda t: Pm @ t, H.vcycle, b, 1e-10, 500)
    assert rr <= 1e-10 and it1 < 40
    assert np.linalg.norm(Pm @ x1 - b) <= 1e-9 * np.linalg.norm(b)


def test_gmres_restatement():
    """oracle/gmres.py: on a symmetric matrix full GMRES and MINRES minimise the same residual, so their
    residual histories agree; restarted and right-preconditioned variants converge to the same solution."""
    from oracle import gmres as og
    P, J, Pm, x = regularised_problem(8, state="random")
    b = -P.compute_f(1.0, x)
    xm, itm, _, hm = amg.pminres(lambda t: J @ t, lambda r: r.copy(), b, 1e-10, 2000)
    xg, itg, rr, hg = og.gmres(lambda t: J @ t, None, b, 1e-10, 2000, restart=2000)
    # (MINRES' short recurrences lose orthogonality late in the run: a few more iterations at the end)
    assert itg <= itm <= itg + 10
    assert np.allclose(hg[:60], hm[:60], rtol=1e-6, atol=0)
    assert np.abs(xg - xm).max() <= 1e-8 * np.abs(xm).max()
    xr, itr, rr, _ = og.gmres(lambda t: J @ t, None, b, 1e-10, 5000, restart=25)
    assert itr > itg and np.linalg.norm(J @ xr - b) <= 1.0001e-10 * np.linalg.norm(b) * 10
    H = amg.Hierarchy(Pm, coarse_max=64, degree=1)
    xp, itp, rr, _ = og.gmres(lambda t: J @ t, H.vcycle, b, 1e-10, 500, restart=50)
    assert itp < itg / 2 and np.linalg.norm(J @ xp - b) <= 1.0001e-10 * np.linalg.norm(b)


def test_arclength_restatement_follows_the_branch_through_a_fold():
    """oracle/continuation.py: the arc-length stepper reproduces the points natural continuation finds on the
    way up, keeps going where natural continuation runs out of solutions (the fold), and every accepted point
    is a solution of F(x, mu) = 0."""
    from oracle import continuation
    coords, cells = meshgen.tetgrid(7)
    psi, A = meshgen.plain_gl_fields(coords)
    P = oracle.OracleProblem(coords, cells, ("explicit", A), nthreads=2)
    x, recs = continuation.arclength(P, 1.0, 0.0, psi, 0.05, 1e-7, 0.1, 2.0, 8)
    assert len(recs) == 9 and recs[0]["param"] == 0.0
    dp = [r["dparam_ds"] for r in recs[1:]]
    mus = [r["param"] for r in recs]
    assert dp[0] > 0.5 and min(dp) < 0                     # starts along +mu, turns around
    k = int(np.argmax(mus))
    assert 0 < k < len(mus) - 1                            # the largest mu is an interior point: a fold
    assert all(r["fnorm"] < 1e-8 for r in recs)
    P.keo_fill(mus[-1])
    assert np.linalg.norm(P.compute_f(1.0, x)) < 1e-8      # the last point solves the equations
    # ||psi|| decreases on the way up to the fold
    norms = [r["norm"] for r in recs]
    assert all(b < a for a, b in zip(norms[:k + 1], norms[1:k + 1]))
    # natural continuation to mu = 0.1 lands on the same branch: same norm as the arc-length curve there
    xn, nrecs = P.continuation(1.0, "mu", 0.0, 0.05, 2, psi)
    assert np.interp(nrecs[-1]["param"], mus[:k + 1], norms[:k + 1]) == pytest.approx(nrecs[-1]["norm"], rel=5e-3)  # linear interpolation of a curved branch


def vcycle_matrix(H, n2):
    return np.column_stack([H.vcycle(e) for e in np.eye(n2)])


def test_vcycle_is_a_convergent_mesh_independent_preconditioner():
    """Multigrid theory instead of another implementation: with a symmetric V(1,1) cycle B on an SPD matrix A the
    eigenvalues of B A lie in (0, 1] (the stationary iteration x += B (b - A x) converges monotonically in the energy
    norm), and for smoothed aggregation their spread does not grow with the mesh.  Dense eigenvalues, two meshes."""
    kappa = {}
    for n in (7, 11):
        P, J, Pm, x = regularised_problem(n, state="random")
        H = amg.Hierarchy(Pm, coarse_max=40)
        assert len(H.levels) >= 2
        n2 = Pm.shape[0]
        B = vcycle_matrix(H, n2)
        assert np.abs(B - B.T).max() <= 1e-12 * np.abs(B).max()
        A = Pm.toarray()
        L = np.linalg.cholesky(A)                        # B A is similar to the symmetric L^T B L
        ev = np.linalg.eigvalsh(L.T @ (0.5 * (B + B.T)) @ L)
        assert ev.min() > 0.0 and ev.max() <= 1.0 + 1e-10
        kappa[n] = ev.max() / ev.min()
        assert 1.0 - ev.min() < 0.9                      # energy-norm contraction number of the cycle
    assert kappa[11] <= 1.5 * kappa[7] and kappa[11] < 12.0


def test_coarse_grid_correction_is_the_energy_projection():
    """P (P^T A P)^-1 P^T A is the A-orthogonal projection onto range(P): idempotent, A-self-adjoint, and it leaves
    range(P) alone -- which holds iff the coarse operator really is the Galerkin product of the stored P."""
    P_, J, Pm, x = regularised_problem(7, state="random")
    H = amg.Hierarchy(Pm, coarse_max=40)
    L0, L1 = H.levels[0], H.levels[1]
    A, Pr, Ac = L0.A.toarray(), L0.P.toarray(), L1.A.toarray()
    assert np.abs(Ac - Pr.T @ A @ Pr).max() <= 1e-12 * np.abs(Ac).max()
    Pi = Pr @ np.linalg.solve(Ac, Pr.T @ A)
    assert np.abs(Pi @ Pi - Pi).max() <= 1e-9
    assert np.abs(A @ Pi - (A @ Pi).T).max() <= 1e-9 * np.abs(A).max()
    assert np.abs(Pi @ Pr - Pr).max() <= 1e-9
    # the tentative prolongator has orthonormal columns, one aggregate per fine node; smoothing keeps the pattern
    # of A P0 and damps with 4/3 / lambda_max(D^-1 A)
    assert L0.agg.min() == 0 and np.unique(L0.agg).size == L1.n
    lam_true = np.linalg.eigvals(np.diag(L0.dinv) @ A).real.max()
    assert 0.7 * lam_true <= L0.lam <= lam_true * (1 + 1e-12)       # power iteration: a lower estimate
    assert L0.omega == pytest.approx(amg.SA_DAMPING / L0.lam)

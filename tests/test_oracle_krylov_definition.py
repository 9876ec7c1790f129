"""The Krylov oracles against the DEFINITION of the methods, by dense linear algebra (no Krylov code involved):
MINRES's k-th iterate minimises ||b - A x||_M over x in span{M b, (M A) M b, ..., (M A)^(k-1) M b}; restarted GMRES
does the same in the 2-norm with a right preconditioner inside every cycle; CG minimises the A-norm of the error.
Belos is absent (parity unpinned); this pins the restatements -- which the device solvers are compared with
iteration by iteration in tests/test_gpu_*.py -- to the mathematics instead of to another implementation."""
import numpy as np
import pytest

import oracle
from oracle import amg as oamg
from oracle import gmres as ogmres


def krylov_basis(op, start, k):
    """orthonormal basis of span{start, op start, ..., op^(k-1) start}: Arnoldi with two re-orthogonalisations"""
    Q = np.zeros((start.size, k))
    v = start / np.linalg.norm(start)
    for j in range(k):
        Q[:, j] = v
        w = op(v)
        for _ in range(3):
            w = w - Q[:, :j + 1] @ (Q[:, :j + 1].T @ w)
        nw = np.linalg.norm(w)
        if nw < 1e-13:
            return Q[:, :j + 1]
        v = w / nw
    return Q


@pytest.fixture(scope="module")
def system():
    coords, cells = oracle.meshgen.tetgrid(4)                     # 64 vertices -> 128 real unknowns
    psi, A = oracle.meshgen.plain_gl_fields(coords)
    P = oracle.OracleProblem(coords, cells, ("explicit", A), V=-1.0, thickness=1.0)
    x = oracle.meshgen.random_state(P.N, 3)
    P.keo_fill(0.4)
    P.jac_rebuild(1.0, x)                                         # symmetric, indefinite
    n = 2 * P.N
    J = np.column_stack([P.jac_apply(e) for e in np.eye(n)])
    assert np.abs(J - J.T).max() <= 1e-12 * np.abs(J).max()
    ev = np.linalg.eigvalsh(0.5 * (J + J.T))
    assert ev.min() < 0 < ev.max()
    b = np.random.default_rng(0).standard_normal(n)
    return P, J, b


def test_minres_iterates_minimise_the_residual(system):
    P, J, b = system
    K = 25
    _, it, _, hist = P.krylov(b, 0.0, K, history=True)
    assert it == K
    Q = krylov_basis(lambda v: J @ v, b, K)
    for k in range(1, K + 1):
        y = np.linalg.lstsq(J @ Q[:, :k], b, rcond=None)[0]
        best = np.linalg.norm(b - J @ (Q[:, :k] @ y)) / np.linalg.norm(b)
        assert hist[k] == pytest.approx(best, rel=1e-8), k
    # ... and the python restatement (the M = I case of the preconditioned one) is the same recurrence
    _, it2, _, hist2 = oamg.pminres(lambda v: J @ v, lambda r: r.copy(), b, 0.0, K)
    assert it2 == K and np.allclose(hist2, hist, rtol=1e-9, atol=0)


def test_preconditioned_minres_minimises_the_m_norm_of_the_residual(system):
    P, J, b = system
    n = b.size
    rng = np.random.default_rng(1)
    G = rng.standard_normal((n, n))
    M = np.eye(n) + 0.05 * (G @ G.T) / n                           # symmetric positive definite
    L = np.linalg.cholesky(M)                                       # M = L L^T
    K = 20
    x, it, _, hist = oamg.pminres(lambda v: J @ v, lambda r: M @ r, b, 0.0, K)
    assert it == K
    # x_k = L y_k with y_k = argmin || L^T b - (L^T J L) y ||_2 over the Krylov space of (L^T J L, L^T b)
    Jh, bh = L.T @ J @ L, L.T @ b
    Q = krylov_basis(lambda v: Jh @ v, bh, K)
    for k in range(1, K + 1):
        y = np.linalg.lstsq(Jh @ Q[:, :k], bh, rcond=None)[0]
        best = np.linalg.norm(bh - Jh @ (Q[:, :k] @ y)) / np.linalg.norm(bh)
        assert hist[k] == pytest.approx(best, rel=1e-8), k
    yK = np.linalg.lstsq(Jh @ Q, bh, rcond=None)[0]
    assert np.linalg.norm(x - L @ (Q @ yK)) <= 1e-8 * np.linalg.norm(x)
    r = b - J @ x
    assert np.sqrt(r @ M @ r) / np.sqrt(b @ M @ b) == pytest.approx(hist[K], rel=1e-8)


def test_restarted_gmres_minimises_inside_every_cycle(system):
    P, J, b = system
    m, cycles = 8, 3
    x, it, _, hist = ogmres.gmres(lambda v: J @ v, None, b, 0.0, m * cycles, restart=m)
    assert it == m * cycles
    x0 = np.zeros_like(b)
    nb = np.linalg.norm(b)
    for c in range(cycles):
        r0 = b - J @ x0
        Q = krylov_basis(lambda v: J @ v, r0, m)
        for k in range(1, m + 1):
            y = np.linalg.lstsq(J @ Q[:, :k], r0, rcond=None)[0]
            best = np.linalg.norm(r0 - J @ (Q[:, :k] @ y)) / nb
            assert hist[c * m + k] == pytest.approx(best, rel=1e-8), (c, k)
        x0 = x0 + Q @ np.linalg.lstsq(J @ Q, r0, rcond=None)[0]
    assert np.linalg.norm(x - x0) <= 1e-8 * np.linalg.norm(x0)


def test_cg_minimises_the_energy_norm_of_the_error():
    coords, cells = oracle.meshgen.tetgrid(4)
    psi, A = oracle.meshgen.plain_gl_fields(coords)
    P = oracle.OracleProblem(coords, cells, ("explicit", A), V=1.0, thickness=1.0)     # V > 0: positive definite
    x = oracle.meshgen.random_state(P.N, 5)
    P.keo_fill(0.2)
    P.jac_rebuild(1.0, x)
    n = 2 * P.N
    J = np.column_stack([P.jac_apply(e) for e in np.eye(n)])
    assert np.linalg.eigvalsh(0.5 * (J + J.T)).min() > 0
    b = np.random.default_rng(2).standard_normal(n)
    xs = np.linalg.solve(J, b)
    K = 12
    Q = krylov_basis(lambda v: J @ v, b, K)
    for k in (1, 2, 5, K):
        xk, it, _ = P.krylov(b, 0.0, k, solver="cg")
        assert it == k
        # the Galerkin solution on the Krylov space minimises ||x* - x||_A
        y = np.linalg.solve(Q[:, :k].T @ J @ Q[:, :k], Q[:, :k].T @ b)
        assert np.linalg.norm(xk - Q[:, :k] @ y) <= 1e-8 * np.linalg.norm(xk), k
        e = xs - xk
        assert e @ J @ e <= (xs @ J @ xs) * (1 + 1e-12)

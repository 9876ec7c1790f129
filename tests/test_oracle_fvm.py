"""CPU: the oracle's restatement of fvm_matrix::fill / fvm_operator::apply (oracle/fvm.py) against properties that
do not depend on it: the Laplace edge core has vanishing row sums, is symmetric positive semi-definite and
reproduces linear functions exactly (a Voronoi finite-volume scheme is exact for them); the boundary-vertex set
of a box is the set of vertices on its faces; the Poisson example (examples/poisson/poisson.py) against a
sparse direct solve; the Bratu operators (examples/bratu/bratu.py) against finite differences."""
import os
import sys

import numpy as np
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import OracleProblem, fvm, meshgen  # noqa: E402


def _problem(n=7, jitter=0.0):
    coords, cells = meshgen.tetgrid(n, jitter=jitter)
    return coords, cells, OracleProblem(coords, cells, ("constcurl", (0.0, 0.0, 1.0), None))


def test_boundary_vertices_of_a_box():
    coords, cells, P = _problem(6)
    b = fvm.boundary_vertices(cells, P.N)
    on_face = (np.abs(np.abs(coords).max(axis=1) - 5.0) < 1e-12).astype(np.int32)
    assert np.array_equal(b, on_face)
    c2, t2 = meshgen.trigrid(7, 4)
    b2 = fvm.boundary_vertices(t2, c2.shape[0])
    assert b2.sum() == 2 * 7 + 2 * 4 - 4


def test_laplace_core_properties_and_exactness_for_linear_functions():
    coords, cells, P = _problem(7, jitter=0.15)
    A, rhs = fvm.fill(P)
    assert np.abs(A @ np.ones(P.N)).max() < 1e-12 and np.all(rhs == 0.0)
    assert abs(A - A.T).max() < 1e-13
    x = np.random.default_rng(0).standard_normal(P.N)
    assert x @ (A @ x) > 0
    # Dirichlet problem with the linear solution u = 1 + 2x - y + 0.5z
    u = 1.0 + 2.0 * coords[:, 0] - coords[:, 1] + 0.5 * coords[:, 2]
    mask = fvm.boundary_vertices(cells, P.N)
    Ad, b = fvm.fill(P, dirichlet_mask=mask, dirichlet_values=u)
    assert np.abs(Ad @ u - b).max() < 1e-10
    sol = spla.spsolve(Ad.tocsc(), b)
    assert np.abs(sol - u).max() < 1e-9
    # rows: unit rows on the boundary, columns untouched (the reference eliminates rows only)
    assert np.array_equal(Ad[mask == 1].toarray(), np.eye(P.N)[mask == 1])
    assert abs(Ad - Ad.T).max() > 0


def test_poisson_example_and_cg():
    """examples/poisson/poisson.py: -Laplace(u) = sin(y), u = 0 on the boundary with y < 0, u = 1 on the rest."""
    coords, cells, P = _problem(8, jitter=0.1)
    bnd = fvm.boundary_vertices(cells, P.N)
    g0 = (bnd == 1) & (coords[:, 1] < 0)
    g1 = (bnd == 1) & (coords[:, 1] >= 0)
    A, b = fvm.fill(P, vertex_rhs=P.cv * np.sin(coords[:, 1]), dirichlet_mask=(g0 | g1).astype(np.int32),
                    dirichlet_values=np.where(g1, 1.0, 0.0))
    sol = spla.spsolve(A.tocsc(), b)
    assert np.abs(sol[g0]).max() < 1e-13 and np.abs(sol[g1] - 1.0).max() < 1e-13
    lift = np.where(g1, 1.0, 0.0)
    x, it, rr = fvm.cg(A, b, 1e-10, 2000, x0=lift)
    assert rr <= 1e-10 and np.abs(x - sol).max() < 1e-7 and 0 < it < 500
    _, it0, rr0 = fvm.cg(A, b, 1e-10, 300)           # from 0 the rows-only elimination is not symmetric: no convergence
    assert rr0 > 1e-3


def test_bratu_operators():
    """examples/bratu/bratu.py: F(u) = NLaplace(u) - int alpha e^u dV, its Jacobian and dF/dalpha."""
    coords, cells, P = _problem(6, jitter=0.1)
    A, _ = fvm.fill(P)
    bnd = fvm.boundary_vertices(cells, P.N)
    alpha = 0.3
    rng = np.random.default_rng(1)
    u = 0.2 * rng.standard_normal(P.N)
    du = rng.standard_normal(P.N)
    F = lambda v, a=alpha: fvm.operator_apply(A, P.cv, v, 1, a, None, bnd, 1)  # noqa: E731
    J = fvm.operator_apply(A, P.cv, du, 2, alpha, u, bnd, 1)
    eps = 1e-6
    fd = (F(u + eps * du) - F(u - eps * du)) / (2 * eps)
    assert np.abs(J - fd).max() <= 1e-7 * np.abs(J).max()
    dFdp = fvm.operator_apply(None, P.cv, u, 1, 1.0, None, bnd, 2)
    fdp = (F(u, alpha + eps) - F(u, alpha - eps)) / (2 * eps)
    assert np.abs(dFdp - fdp).max() <= 1e-7 * np.abs(dFdp).max()
    assert np.all(F(u)[bnd == 1] == u[bnd == 1]) and np.all(dFdp[bnd == 1] == 0.0)

"""CPU: pin the oracle against the reference's own known-answer numbers.

Every literal comes from /root/reference/test/*.cpp via tests/golden/reference_known_answers.json.
The reference compares with Catch's Approx (relative ~1.2e-5); the oracle matches far tighter.
Sums that cancel (1^T K 1 on cubesmall: entries O(5) summing to 1.7e-4) are checked at 1e-9.
"""
import numpy as np
import pytest

from oracle import OracleProblem, meshgen

MU = 0.01


def _problem(name):
    coords, cells = getattr(meshgen, name)()
    psi, A = meshgen.plain_gl_fields(coords)
    P = OracleProblem(coords, cells, ("explicit", A), V=-1.0, thickness=1.0)
    return P, psi, A


def _basis(N):
    one = np.ones(2 * N)
    er = np.zeros(2 * N)
    er[0::2] = 1.0
    ei = np.zeros(2 * N)
    ei[1::2] = 1.0
    return one, er, ei


@pytest.mark.parametrize("name", ["rectanglesmall", "cubesmall"])
def test_mesh_and_io(name, golden):
    P, psi, A = _problem(name)
    g = golden[name]
    assert P.N == g["mesh"]["num_nodes"]
    assert np.abs(P.cv).sum() == pytest.approx(g["mesh"]["cv_norm1"], rel=1e-13)
    assert np.linalg.norm(P.cv) == pytest.approx(g["mesh"]["cv_norm2"], rel=1e-13)
    assert np.abs(P.cv).max() == pytest.approx(g["mesh"]["cv_norminf"], rel=1e-13)
    z = psi[0::2] + 1j * psi[1::2]
    assert np.abs(z).sum() == pytest.approx(g["io"]["psi_norm1"])
    assert np.abs(z).max() == pytest.approx(g["io"]["psi_norminf"])
    assert np.allclose(np.abs(A).max(axis=0), g["io"]["A_norminf"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", ["rectanglesmall", "cubesmall"])
def test_keo(name, golden):
    P, psi, _ = _problem(name)
    g = golden[name]["keo"]
    P.keo_fill(MU)
    one, er, ei = _basis(P.N)
    assert one @ P.keo_apply(one) == pytest.approx(g["sum"], rel=1e-9)
    assert er @ P.keo_apply(er) == pytest.approx(g["sum_real"], rel=1e-9)
    # Hermitian <=> symmetric in the real layout (test/keo.cpp:110-113 intends this)
    import scipy.sparse as sp
    M = sp.csr_matrix((P.vals, P.cols, P.rowptr), shape=(2 * P.N, 2 * P.N))
    assert abs(M - M.T).max() == 0.0
    # sum of the "imaginary parts": e_r^T K e_i = 0
    assert abs(er @ P.keo_apply(ei)) < 1e-15


def test_keo_matrix_rectangle(golden):
    """The full 8x8 matrix printed at test/keo.cpp:121-130 (6 significant digits)."""
    P, _, _ = _problem("rectanglesmall")
    g = golden["rectanglesmall"]["keo_matrix"]
    P.keo_fill(MU)
    rp, cols, K = P.complex_blocks(P.vals)
    dense = np.zeros((4, 4), complex)
    for i in range(4):
        dense[i, cols[rp[i]:rp[i + 1]]] = K[rp[i]:rp[i + 1]]
    assert np.allclose(np.diag(dense).real, g["diag"], rtol=1e-6)
    assert np.all(np.diag(dense).imag == 0)
    lengths = {}
    for a, b in P.edges:
        lengths[(a, b)] = np.linalg.norm(P.coords[a] - P.coords[b])
    for (a, b), ln in lengths.items():
        z = dense[a, b]
        if abs(ln - 1.0) < 1e-12:      # short edges (length 1): alpha = 5
            exp = g["long_edge"]
        elif abs(ln - 10.0) < 1e-12:   # long edges (length 10): alpha = 0.05
            exp = g["short_edge"]
        else:                          # the diagonal: coefficient ~ 0
            assert abs(z) < g["hypotenuse_abs_max"]
            continue
        assert z.real == pytest.approx(exp[0], rel=2e-6)
        assert abs(z.imag) == pytest.approx(abs(exp[1]), rel=2e-6)
        assert dense[b, a] == np.conj(z)


@pytest.mark.parametrize("name", ["rectanglesmall", "cubesmall"])
def test_compute_f(name, golden):
    P, psi, _ = _problem(name)
    g = golden[name]["compute_f"]
    P.keo_fill(MU)
    f = P.compute_f(1.0, psi)
    assert np.abs(f).sum() == pytest.approx(g["norm1"], rel=1e-9)
    assert np.linalg.norm(f) == pytest.approx(g["norm2"], rel=1e-9)
    assert np.abs(f).max() == pytest.approx(g["norminf"], rel=1e-9)


@pytest.mark.parametrize("name", ["rectanglesmall", "cubesmall"])
def test_jacobian(name, golden):
    P, psi, _ = _problem(name)
    g = golden[name]["jac"]
    P.keo_fill(MU)
    P.jac_rebuild(1.0, psi)
    one, er, ei = _basis(P.N)
    assert one @ P.jac_apply(one) == pytest.approx(g["t0"], rel=1e-12)
    assert er @ P.jac_apply(er) == pytest.approx(g["t1"], rel=1e-12)
    assert ei @ P.jac_apply(ei) == pytest.approx(g["t2"], rel=1e-9)


def test_dfdp_finite_difference():
    """test/dfdp.cpp:12-49,124-138: dF/dg against a central difference, eps = 1e-8, mu = 0."""
    P, psi, _ = _problem("cubesmall")
    eps = 1e-8
    P.keo_fill(0.0)
    fd = (P.compute_f(1.0 + eps, psi) - P.compute_f(1.0 - eps, psi)) * (0.5 / eps)
    P.dkeo_fill(0.0, dname="g")  # explicit_values: zero derivative for any name but "mu"
    assert np.all(P.dvals == 0.0)
    an = P.compute_dfdp(psi, is_g=True)
    assert np.abs(fd - an).max() < 1e-6


def test_dkeo_dmu_finite_difference():
    coords, cells = meshgen.tetgrid(5)
    psi, A = meshgen.plain_gl_fields(coords)
    x = meshgen.random_state(coords.shape[0])
    P = OracleProblem(coords, cells, ("explicit", A))
    mu, eps = 0.7, 1e-6
    kp = P.keo_fill(mu + eps).copy()
    km = P.keo_fill(mu - eps).copy()
    dk = P.dkeo_fill(mu, dname="mu")
    fd = (P.csr_apply(kp, x) - P.csr_apply(km, x)) / (2 * eps)
    an = P.csr_apply(dk, x)
    assert np.abs(fd - an).max() / np.abs(an).max() < 1e-8


def test_constcurl_equals_explicit():
    """constantCurl(B) == explicit_values(0.5 B x X)."""
    coords, cells = meshgen.tetgrid(5)
    _, A = meshgen.plain_gl_fields(coords, B=(0.0, 0.0, 1.0))
    Pe = OracleProblem(coords, cells, ("explicit", A))
    Pc = OracleProblem(coords, cells, ("constcurl", (0.0, 0.0, 1.0), None))
    ae, _ = Pe.edge_projection(0.3)
    ac, _ = Pc.edge_projection(0.3)
    assert np.abs(ae - ac).max() < 1e-14
    # rotation about u by theta, derivative vs finite difference
    Pr = OracleProblem(coords, cells, ("constcurl", (0.0, 0.0, 1.0), (1.0, 0.0, 0.0)))
    th, eps = 0.4, 1e-6
    _, dth = Pr.edge_projection(0.3, th, "theta")
    ap, _ = Pr.edge_projection(0.3, th + eps)
    am, _ = Pr.edge_projection(0.3, th - eps)
    # NB the reference's dRotateDTheta_ carries (1+sin) where the derivative of (1-cos) is sin
    # (src/vector_field_constant_curl.cpp:192-195); with u perpendicular to b that term vanishes.
    assert np.abs((ap - am) / (2 * eps) - dth).max() < 1e-8


def test_structural_identities():
    """sum of control volumes = |domain|, mu = 0 => K 1 = 0, J symmetric."""
    coords, cells = meshgen.tetgrid(6)
    psi, A = meshgen.plain_gl_fields(coords)
    P = OracleProblem(coords, cells, ("explicit", A))
    assert P.cv.sum() == pytest.approx(1000.0, rel=1e-12)
    assert P.cv.min() > 0
    P.keo_fill(0.0)
    er = np.zeros(2 * P.N)
    er[0::2] = 1.0
    assert np.abs(P.keo_apply(er)).max() < 1e-12
    P.keo_fill(1.0)
    x = meshgen.random_state(P.N, 1)
    y = meshgen.random_state(P.N, 2)
    P.jac_rebuild(1.0, x)
    assert x @ P.jac_apply(y) == pytest.approx(y @ P.jac_apply(x), rel=1e-12)


def test_minres_against_scipy():
    """Iterates of the restated MINRES agree with SciPy's Paige-Saunders MINRES."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    coords, cells = meshgen.tetgrid(6)
    psi, A = meshgen.plain_gl_fields(coords)
    P = OracleProblem(coords, cells, ("explicit", A))
    P.keo_fill(1.0)
    x0 = meshgen.random_state(P.N, 3)
    P.jac_rebuild(1.0, x0)
    n = 2 * P.N
    D = sp.lil_matrix((n, n))
    k = np.arange(P.N)
    D[2 * k, 2 * k] = P.d0[0::2]
    D[2 * k + 1, 2 * k + 1] = P.d0[1::2]
    D[2 * k, 2 * k + 1] = P.d1b
    D[2 * k + 1, 2 * k] = P.d1b
    J = sp.csr_matrix((P.vals, P.cols, P.rowptr), shape=(n, n)) + D.tocsr()
    b = meshgen.random_state(P.N, 4)
    assert np.abs(J @ b - P.jac_apply(b)).max() < 1e-12
    x, it, rr, hist = P.krylov(b, 1e-10, 2000, history=True)
    assert rr <= 1e-10 and it < 2000
    assert np.linalg.norm(J @ x - b) / np.linalg.norm(b) < 1e-9
    assert np.all(np.diff(hist) <= 1e-15)  # MINRES residuals are monotone
    xs, info = spla.minres(J, b, rtol=1e-12, maxiter=5000)
    assert info == 0
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-7


def test_newton_converges():
    coords, cells = meshgen.tetgrid(6)
    psi, A = meshgen.plain_gl_fields(coords)
    P = OracleProblem(coords, cells, ("explicit", A))
    P.keo_fill(0.1)
    x, steps, lin, fn = P.newton(1.0, psi, nl_tol=1e-8, nl_maxit=20, lin_tol=1e-10, lin_maxit=2000)
    assert fn[-1] < 1e-8 and steps < 20
    assert np.linalg.norm(P.compute_f(1.0, x)) < 1e-8


def test_threaded_baseline_matches_serial():
    coords, cells = meshgen.tetgrid(8)
    psi, A = meshgen.plain_gl_fields(coords)
    P1 = OracleProblem(coords, cells, ("explicit", A), nthreads=1)
    P4 = OracleProblem(coords, cells, ("explicit", A), nthreads=4)
    v1 = P1.keo_fill(0.3).copy()
    v4 = P4.keo_fill(0.3).copy()
    # threaded fill adds with atomics: the diagonal sums may be reordered (rounding only)
    assert np.abs(v1 - v4).max() <= 1e-14 * np.abs(v1).max()
    x = meshgen.random_state(P1.N)
    P1.jac_rebuild(1.0, x)
    P4.jac_rebuild(1.0, x)
    y1, y4 = P1.jac_apply(x), P4.jac_apply(x)
    assert np.abs(y1 - y4).max() <= 1e-13 * np.abs(y1).max()

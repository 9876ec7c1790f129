"""Full-size checks (BASELINE.json configs[1..2] sizes: tetgrid n=200 = 8.0M vertices) through properties that
do not need the oracle -- it would take hours at this size: volume conservation, Hermitian K, the null
vector at mu = 0, the reference tests' quadratic-form identity (test/keo.cpp:134-135), dK/dmu and J against
central differences (the check test/dfdp.cpp makes for dF/dg), linearity, MINRES' implicit residual against
the true one, and the V-cycle's symmetry / definiteness -- and against tests/golden/fullsize_n200.json, known-answer
hashes (norms and quadratic forms, the style of the reference's own tests) that the CPU oracle produced for
exactly this mesh (tests/golden/make_fullsize_golden.py; minutes of CPU time, so not run on the GPU box).
Vectors live on the device (torch) and are used in place by the C ABI."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_AXIS = int(os.environ.get("NOSH_FULLSIZE_N", "200"))


class Ordered:
    """The ctx enqueues on its own stream and device-pointer calls return without waiting; torch works on
    its current stream.  This proxy orders the two (a caller would pass torch's stream to nosh_ctx_create or
    call nosh_ctx_synchronize): wait for torch before a call, for the ctx after it."""

    def __init__(self, ctx, torch):
        self._ctx, self._torch = ctx, torch

    def __getattr__(self, name):
        fn = getattr(self._ctx, name)
        if not callable(fn):
            return fn

        def call(*a, **kw):
            self._torch.cuda.synchronize()
            r = fn(*a, **kw)
            self._ctx.synchronize()
            return r
        return call


@pytest.fixture(scope="module")
def setup():
    import torch
    import nosh_b200
    ctx = Ordered(nosh_b200.Context(), torch)
    mi = ctx.mesh_tetgrid(N_AXIS)
    ctx.set_thickness(None, 1.0)
    ctx.set_potential_constant(-1.0)
    ctx.set_mvp_constcurl((0.0, 0.0, 1.0))
    No = int(mi.n_owned)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    rnd = lambda: torch.randn(2 * No, generator=g, device="cuda", dtype=torch.float64)  # noqa: E731
    yield ctx, mi, No, rnd, torch, nosh_b200
    ctx._ctx.close()


def test_volume_and_sizes(setup):
    ctx, mi, No, rnd, torch, nb = setup
    assert mi.n_global == N_AXIS ** 3 == No
    assert mi.n_cells == 6 * (N_AXIS - 1) ** 3
    cv = ctx.control_volumes()
    assert cv.min() > 0
    # |[-5,5]^3| = 1000 (test/mesh.cpp's ||c||_1).  The reference's signed-covolume construction
    # (src/mesh_tetra.cpp:273-331) is exact for well-shaped cells only: on the 8M-vertex jittered mesh the
    # oracle itself sums to 1000.00000057 (golden below), on <= 1M vertices to exactly 1000.
    assert cv.sum() == pytest.approx(1000.0, rel=1e-8)


def test_against_full_size_oracle_hashes(setup):
    """the oracle's known answers for this very mesh (generated offline, see the module docstring)"""
    ctx, mi, No, rnd, torch, nb = setup
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_n%d.json" % N_AXIS)
    if not os.path.exists(path):
        pytest.skip("no golden file for n = %d" % N_AXIS)
    G = json.load(open(path))
    from oracle import meshgen
    assert mi.n_owned == G["num_nodes"] and mi.n_edges == G["num_edges"]
    cv = ctx.control_volumes()
    assert np.abs(cv).sum() == pytest.approx(G["cv_norm1"], rel=1e-13)
    assert np.linalg.norm(cv) == pytest.approx(G["cv_norm2"], rel=1e-13)
    assert np.abs(cv).max() == pytest.approx(G["cv_norminf"], rel=1e-13)
    assert cv.min() == pytest.approx(G["cv_min"], rel=1e-10)
    al = ctx.alpha_cache()
    assert al.sum() == pytest.approx(G["alpha_sum"], rel=1e-11)
    assert np.linalg.norm(al) == pytest.approx(G["alpha_norm2"], rel=1e-12)
    par = {"g": G["g"], "mu": G["mu"], "theta": G["theta"]}
    dev = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
    one = torch.ones(2 * No, device="cuda", dtype=torch.float64)
    er = torch.zeros_like(one)
    er[0::2] = 1.0
    ei = one - er
    x = dev(meshgen.random_state(No, 42))
    out = torch.empty_like(one)
    ctx.keo_fill(par)
    # quadratic forms with heavy cancellation (K 1 is almost 0): absolute scale = |x^T K x| of a generic x
    ctx.keo_apply(x, out)
    xkx = torch.dot(x, out).item()
    assert xkx == pytest.approx(G["keo_xKx"], rel=1e-12)
    assert torch.linalg.vector_norm(out).item() == pytest.approx(G["keo_Kx_norm2"], rel=1e-12)
    ctx.keo_apply(one, out)
    assert torch.dot(one, out).item() == pytest.approx(G["keo_1K1"], rel=1e-7, abs=1e-13 * abs(xkx))
    ctx.keo_apply(er, out)
    assert torch.dot(er, out).item() == pytest.approx(G["keo_erKer"], rel=1e-7, abs=1e-13 * abs(xkx))
    f = torch.empty_like(one)
    ctx.compute_f(par, er, f)
    assert f.abs().sum().item() == pytest.approx(G["F_norm1"], rel=1e-10)
    assert torch.linalg.vector_norm(f).item() == pytest.approx(G["F_norm2"], rel=1e-10)
    assert f.abs().max().item() == pytest.approx(G["F_norminf"], rel=1e-10)
    ctx.compute_f(par, x, f)
    assert torch.linalg.vector_norm(f).item() == pytest.approx(G["F_random_state_norm2"], rel=1e-12)
    ctx.jac_rebuild(par, x)
    for name, s in (("one", one), ("er", er), ("ei", ei), ("x", x)):
        ctx.jac_apply(s, out)
        assert torch.dot(s, out).item() == pytest.approx(G["jac_%s" % name], rel=1e-10), name
    ctx.compute_dfdp(par, "mu", x, out)
    assert torch.linalg.vector_norm(out).item() == pytest.approx(G["dfdmu_norm2"], rel=1e-12)
    # entry level at 8.0M vertices: K x, F(x), J y, dF/dmu at 4096 strided vertices against the oracle's values
    # (every sampled row involves its ~15 matrix blocks); north_star's bar: 1e-12 relative
    S = G.get("sample_rows")
    if S:
        rows = torch.arange(0, No, S["stride"], device="cuda")[:S["count"]]
        idx = torch.stack([2 * rows, 2 * rows + 1], 1).reshape(-1)
        y = dev(meshgen.random_state(No, 43))
        worst = {}
        ctx.keo_apply(x, out)
        worst["Kx"] = (out[idx].cpu().numpy(), S["Kx"], S["Kx_max"])
        ctx.compute_f(par, x, f)
        worst["Fx"] = (f[idx].cpu().numpy(), S["Fx"], S["Fx_max"])
        ctx.jac_apply(y, out)
        worst["Jy"] = (out[idx].cpu().numpy(), S["Jy"], S["Jy_max"])
        ctx.compute_dfdp(par, "mu", x, out)
        worst["dFdmu"] = (out[idx].cpu().numpy(), S["dFdmu"], S["dFdmu_max"])
        for name, (got, want, scale) in worst.items():
            err = np.abs(got - np.asarray(want)).max() / scale
            print("  sampled rows, %s: max |gpu - oracle| / max|oracle| = %.2e" % (name, err))
            assert err <= 1e-12, (name, err)


def test_keo_structure(setup):
    ctx, mi, No, rnd, torch, nb = setup
    one = torch.zeros(2 * No, device="cuda", dtype=torch.float64)
    one[0::2] = 1.0
    er = one.clone()
    x, y = rnd(), rnd()
    out, out2 = torch.empty_like(x), torch.empty_like(x)
    # mu = 0: constants are in the null space of K (src/parameter_matrix_keo.cpp:119-126 with a = 0)
    ctx.keo_fill({"mu": 0.0, "theta": 0.0})
    ctx.keo_apply(one, out)
    ctx.keo_apply(x, out2)
    assert out.abs().max().item() <= 1e-11 * out2.abs().max().item()
    # mu = 1: Hermitian (real form symmetric), and 1^T K 1 = 2 e_r^T K e_r (test/keo.cpp:134-135: one is
    # (1,1,...) over BOTH components there; here `full` is that vector)
    ctx.keo_fill({"mu": 1.0, "theta": 0.0})
    kx, ky = torch.empty_like(x), torch.empty_like(x)
    ctx.keo_apply(x, kx)
    ctx.keo_apply(y, ky)
    a, b = torch.dot(y, kx).item(), torch.dot(x, ky).item()
    assert abs(a - b) <= 1e-12 * abs(a)
    full = torch.ones(2 * No, device="cuda", dtype=torch.float64)
    ctx.keo_apply(full, out)
    ctx.keo_apply(er, out2)
    q_full, q_er = torch.dot(full, out).item(), torch.dot(er, out2).item()
    assert q_full == pytest.approx(2.0 * q_er, rel=1e-9)
    assert torch.dot(x, kx).item() > 0                            # positive (semi-)definite


def test_derivatives_against_central_differences(setup):
    ctx, mi, No, rnd, torch, nb = setup
    par = {"g": 1.0, "mu": 0.7, "theta": 0.0}
    psi, d = rnd(), rnd()
    # dF/dg with a constantCurl field: get_d_edge_projection_dp knows "mu" and "theta" only and throws for
    # anything else (src/vector_field_constant_curl.cpp:135-139)
    with pytest.raises(ValueError, match="Illegal parameter"):
        ctx.compute_dfdp(par, "g", psi)
    # dF/dmu (computeDFDP_) vs (F(p+e) - F(p-e)) / 2e, e = 1e-6
    for name in ("mu",):
        e = 1e-6
        fp, fm, df = torch.empty_like(psi), torch.empty_like(psi), torch.empty_like(psi)
        ctx.compute_f(dict(par, **{name: par[name] + e}), psi, fp)
        ctx.compute_f(dict(par, **{name: par[name] - e}), psi, fm)
        ctx.compute_dfdp(par, name, psi, df)
        fd = (fp - fm) / (2 * e)
        assert (fd - df).abs().max().item() <= 1e-7 * df.abs().max().item(), name
    # J(psi) d vs (F(psi + e d) - F(psi - e d)) / 2e
    e = 1e-6
    fp, fm, jd = torch.empty_like(psi), torch.empty_like(psi), torch.empty_like(psi)
    ctx.compute_f(par, psi + e * d, fp)
    ctx.compute_f(par, psi - e * d, fm)
    ctx.jac_rebuild(par, psi)
    ctx.jac_apply(d, jd)
    fd = (fp - fm) / (2 * e)
    assert (fd - jd).abs().max().item() <= 1e-7 * jd.abs().max().item()
    # real-linear and symmetric
    x, y = rnd(), rnd()
    jx, jy, jxy = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    ctx.jac_apply(x, jx)
    ctx.jac_apply(y, jy)
    ctx.jac_apply(2.0 * x - 3.0 * y, jxy)
    assert (jxy - (2.0 * jx - 3.0 * jy)).abs().max().item() <= 1e-12 * jxy.abs().max().item()
    a, b = torch.dot(y, jx).item(), torch.dot(x, jy).item()
    assert abs(a - b) <= 1e-12 * abs(a)
    # the reductions agree with torch's (different summation trees)
    assert ctx.dot(x, y) == pytest.approx(torch.dot(x, y).item(), rel=1e-10)


def test_minres_residuals_and_preconditioner(setup):
    ctx, mi, No, rnd, torch, nb = setup
    par = {"g": 1.0, "mu": 0.1, "theta": 0.0}
    psi = torch.zeros(2 * No, device="cuda", dtype=torch.float64)
    psi[0::2] = 1.0
    b = rnd()
    ctx.jac_rebuild(par, psi)
    x, jx = torch.empty_like(b), torch.empty_like(b)
    # implicit residual of MINRES == true residual (no preconditioner: 2-norm)
    x, res, hist = ctx.minres(b, x, tol=0.0, maxit=60, history=True)
    assert res.iterations == 60 and np.all(np.diff(hist) <= 1e-14)          # monotone
    ctx.jac_apply(x, jx)
    true = (torch.linalg.vector_norm(b - jx) / torch.linalg.vector_norm(b)).item()
    assert true == pytest.approx(hist[-1], rel=1e-6)
    # V-cycle: symmetric, positive; preconditioned MINRES reaches a true residual of 1e-8
    ctx.keoreg_rebuild(par, psi)
    u, v = rnd(), rnd()
    mu_, mv = torch.empty_like(u), torch.empty_like(u)
    ctx.keoreg_apply(u, mu_)
    ctx.keoreg_apply(v, mv)
    a, c = torch.dot(v, mu_).item(), torch.dot(u, mv).item()
    assert abs(a - c) <= 1e-11 * abs(a)
    assert torch.dot(u, mu_).item() > 0 and torch.dot(v, mv).item() > 0
    info = ctx.amg_info()
    assert info.levels >= 3 and info.nodes[1] < No / 15
    x, res = ctx.minres(b, x, tol=1e-10, maxit=500, prec=nb.PREC_KEOREG_AMG)
    assert res.converged == 1 and res.iterations < 200
    ctx.jac_apply(x, jx)
    assert (torch.linalg.vector_norm(b - jx) / torch.linalg.vector_norm(b)).item() <= 1e-8


# ---------------------------------------------------------------------------------------------------------------
# Iteration counts at benchmark size (north_star: "Newton/MINRES iteration counts must be identical").
# tests/golden/counts_n100.json / counts_n200.json hold the oracle's MINRES and Newton counts on the 1.0M-vertex mesh
# of BASELINE.json configs[1] and the 8.0M-vertex mesh of configs[2] for dot products split into 1, 2, 7 and 16 parts (the reference's own counts depend on
# its MPI rank count in the same way; it tests with 1, 2 and 7 ranks).  The oracle's own spread over the
# summation order is printed next to the GPU's numbers: where the oracle agrees with itself the GPU must agree
# exactly; where it does not, the GPU must lie inside the oracle's spread widened by that same spread.
# ---------------------------------------------------------------------------------------------------------------
def _counts_golden(n):
    p = os.path.join(os.path.dirname(__file__), "golden", "counts_n%d.json" % n)
    if not os.path.exists(p):
        pytest.skip("no counts golden for n=%d" % n)
    return json.load(open(p))


def _in_spread(value, oracle_values):
    """The oracle's counts for different summation orders span [lo, hi]; the GPU (yet another summation order)
    must lie within that interval widened by max(2, 3 (hi - lo))."""
    lo, hi = min(oracle_values), max(oracle_values)
    slack = max(2, 3 * (hi - lo))
    return lo - slack <= value <= hi + slack


@pytest.mark.parametrize("n", [100, 200])
def test_minres_and_newton_iteration_counts_at_benchmark_size(n):
    import nosh_b200
    sys_path_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import sys
    sys.path.insert(0, sys_path_root)
    from oracle import meshgen
    gold = _counts_golden(n)
    ctx = nosh_b200.Context()
    mi = ctx.mesh_tetgrid(n)
    N = int(mi.n_owned)
    assert N == gold["num_nodes"]
    ctx.set_thickness(None, 1.0)
    ctx.set_potential_constant(-1.0)
    ctx.set_mvp_constcurl((0.0, 0.0, 1.0))
    # MINRES on the benchmark operator (mu = 1, random state, random right-hand side), tol 1e-10
    gm = gold["minres"]
    par = {"g": gm["g"], "mu": gm["mu"], "theta": 0.0}
    psi = meshgen.random_state(N, 42)
    b = meshgen.random_state(N, 43)
    ctx.jac_rebuild(par, psi)
    x, res, hist = ctx.minres(b, tol=gm["tol"], maxit=gm["maxit"], history=True)
    ocounts = {k: v["iterations"] for k, v in gm["by_parts"].items()}
    print("MINRES n=%d: GPU %d iterations; oracle by dot partition %s" % (n, res.iterations, ocounts))
    assert res.converged == 1
    assert _in_spread(res.iterations, list(ocounts.values())), (res.iterations, ocounts)
    # The residual history is the same Krylov process: over the first 10 iterations GPU and oracle agree to 1e-8
    # (measured: 1e-14 ... 3e-12).  After that the Lanczos recurrence amplifies every rounding-level difference --
    # between the oracle's own runs (1.4 % at iteration 100 on the 1.0M-vertex operator, 1.6 % at iteration 50 on the
    # 8.0M-vertex one) and between oracle and GPU, whose operators agree to ~1e-12 only on the handful of badly
    # shaped cells of the 8.0M-vertex mesh (6x6 full-pivot solves, see fullsize_n200.json's tolerances): 1.7e-4 at
    # iteration 20 there.  Those deviations are printed (GPU next to the oracle's spread over dot partitions, accurate
    # dot products and a reversed row order of the SpMV); asserted is that the GPU stays within a factor 2 of the
    # oracle's median, i.e. converges at the same rate (measured: 0.64 at iteration 1000 of the 8.0M-vertex solve --
    # the GPU's tree-summed dot products are more accurate than the oracle's long sequential sums and it needs 3 %
    # fewer iterations; the oracle's own counts fall in the same direction as its sums get shorter).
    runs = list(gm["by_parts"].values())
    for j, k in enumerate(gm["hist_at"]):
        if k > res.iterations:
            continue
        vals = [v["hist"][j] for v in runs if j < len(v["hist"])]
        centre = float(np.median(vals))
        own = (max(vals) - min(vals)) / centre
        dev = abs(hist[k] - centre) / centre
        print("  relres at iteration %4d: GPU %.6e, oracle median %.6e, GPU deviation %.2e, oracle spread %.2e"
              % (k, hist[k], centre, dev, own))
        if k <= 10:
            assert dev <= 1e-8, (k, hist[k], vals)
        else:
            assert centre / 2.0 <= hist[k] <= centre * 2.0, (k, hist[k], vals)
    # the solve is verified independently of any history: true residual of the GPU's solution
    r = ctx.jac_apply(x)
    true_relres = np.linalg.norm(r - b) / np.linalg.norm(b)
    print("  true relative residual of the GPU solution: %.3e (oracle runs: %s)"
          % (true_relres, ["%.3e" % v["true_relres"] for v in runs]))
    assert true_relres < 2e-10
    ref = gm["by_parts"]["1"]
    assert np.linalg.norm(x) == pytest.approx(ref["x_norm2"], rel=1e-7)
    for v in runs:                                          # every run solved the system to the tolerance it was given
        assert v["true_relres"] < 2e-10
    # full Newton-MINRES solve (configs[2] at 1.0M vertices): psi0 = 1, mu = 0.1
    gn = gold["newton"]
    psi0 = np.zeros(2 * N)
    psi0[0::2] = 1.0
    nres, lin, fn = ctx.newton({"g": gn["g"], "mu": gn["mu"], "theta": 0.0}, psi0, gn["nl_tol"], 20, gn["lin_tol"],
                               gn["lin_maxit"])
    osteps = {k: v["steps"] for k, v in gn["by_parts"].items()}
    olin = {k: v["minres_iterations"] for k, v in gn["by_parts"].items()}
    print("Newton n=%d: GPU %d steps %s; oracle by dot partition %s" % (n, nres.steps, list(lin), olin))
    assert nres.converged == 1 and nres.linear_solve_status == 0
    assert set(osteps.values()) == {int(nres.steps)}
    for j in range(nres.steps):
        assert _in_spread(int(lin[j]), [v[j] for v in olin.values()]), (j, int(lin[j]), olin)
    for k, v in gn["by_parts"].items():
        for j in range(nres.steps):                      # the nonlinear residuals themselves are well conditioned
            assert fn[j] == pytest.approx(v["fnorms"][j], rel=1e-6), (k, j)
    assert np.linalg.norm(psi0) == pytest.approx(gn["by_parts"]["1"]["x_norm2"], rel=1e-10)
    ctx.close()

"""Small end-to-end exercise of every kernel family for compute-sanitizer (no torch import):
   compute-sanitizer --tool memcheck python profiles/sanitizer_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nosh_b200  # noqa: E402

n = 11
ctx = nosh_b200.Context()
mi = ctx.mesh_tetgrid(n)
ctx.set_thickness(None, 1.0)
ctx.set_potential_constant(-1.0)
ctx.set_mvp_constcurl((0.0, 0.0, 1.0))
N = int(mi.n_owned)
rng = np.random.default_rng(0)
psi = rng.standard_normal(2 * N)
b = rng.standard_normal(2 * N)
par = {"g": 1.0, "mu": 0.3, "theta": 0.0}
ctx.jac_rebuild(par, psi)
ctx.compute_f(par, psi)
ctx.compute_dfdp(par, "mu", psi)
ctx.jac_apply(b)
for persistent in (1, 0):
    ctx.set_tuning("persistent_minres", persistent)
    x, res = ctx.minres(b, tol=1e-10, maxit=400)
    assert res.converged == 1
ctx.cg(b, op=nosh_b200.OP_KEO, tol=1e-8, maxit=400)
ctx.amg_set_options(coarse_max=40)
ctx.keoreg_rebuild(par, psi)
ctx.keoreg_apply(b)
x, res = ctx.minres(b, tol=1e-10, maxit=200, prec=nosh_b200.PREC_KEOREG_AMG)
assert res.converged == 1
x, res = ctx.gmres(b, tol=1e-10, maxit=200, restart=30, prec=nosh_b200.PREC_KEOREG_AMG)
assert res.converged == 1
x, res = ctx.cg(b, op=nosh_b200.OP_KEOREG, tol=1e-10, maxit=200, prec=nosh_b200.PREC_KEOREG_AMG)
one = np.zeros(2 * N)
one[0::2] = 1.0
ctx.set_preconditioner(nosh_b200.PREC_KEOREG_AMG)
ctx.newton({"g": 1.0, "mu": 0.1, "theta": 0.0}, one.copy())
ctx.continuation_arclength({"g": 1.0, "mu": 0.0, "theta": 0.0}, "mu", one.copy(), initial_step_size=0.05,
                           max_step_size=0.1, max_steps=2)
print("probe ok, %d launches" % ctx.launch_count())
ctx.close()

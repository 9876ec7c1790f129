"""Profiling driver (run under ncu with --profile-from-start off): one bench step of the
hot path on tetgrid n, bracketed by cudaProfilerStart/Stop so that set-up kernels are excluded.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/profile_step.py --n 100 --iters 20
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:k_apply_sell -c 2 -o gpurun_out/prof_apply python profiles/profile_step.py --n 200 --iters 3
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nosh_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", "--mesh-n", dest="n", type=int, default=100)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--layout", default="sell32")
ap.add_argument("--precond", default="none", choices=["none", "amg"])
ap.add_argument("--persistent", type=int, default=0,
                help="0: the multi-launch MINRES loop (one launch per phase: readable launch lists); 1: the default "
                     "one-launch cooperative kernel")
ap.add_argument("--amg-mixed", type=int, default=0, help="1: the mixed-precision cycle (tuning key amg_mixed)")
ap.add_argument("--apply-variant", type=int, default=0)
ap.add_argument("--amg-degree", type=int, default=1)
ap.add_argument("--amg-coarse-degree", type=int, default=2)
a = ap.parse_args()

ctx = nosh_b200.Context(layout={"csr": 0, "sell32": 1}[a.layout])
mi = ctx.mesh_tetgrid(a.n)
ctx.set_thickness(None, 1.0)
ctx.set_potential_constant(-1.0)
ctx.set_mvp_constcurl((0.0, 0.0, 1.0))
No = int(mi.n_owned)
g = torch.Generator(device="cuda")
g.manual_seed(1)
psi = torch.randn(2 * No, generator=g, device="cuda", dtype=torch.float64)
b = torch.randn(2 * No, generator=g, device="cuda", dtype=torch.float64)
x = torch.empty_like(b)
par = {"g": 1.0, "mu": 1.0, "theta": 0.0}


ctx.set_tuning("persistent_minres", a.persistent)
ctx.set_tuning("apply_variant", a.apply_variant)
if a.precond == "amg":
    ctx.amg_set_options(degree=a.amg_degree, coarse_degree=a.amg_coarse_degree)
    ctx.set_tuning("amg_mixed", a.amg_mixed)


def step(k):
    par["mu"] = 1.0 + 1e-9 * k
    ctx.keo_fill(par)
    ctx.jac_rebuild(par, psi)
    if a.precond == "amg":
        ctx.keoreg_rebuild(par, psi)
        ctx.minres(b, x, tol=0.0, maxit=a.iters, prec=nosh_b200.PREC_KEOREG_AMG)
    else:
        ctx.minres(b, x, tol=0.0, maxit=a.iters)


step(0)
ctx.synchronize()
torch.cuda.profiler.start()
step(1)
ctx.synchronize()
torch.cuda.profiler.stop()
print("profiled one step: n=%d iters=%d launches=%d" % (a.n, a.iters, ctx.launch_count()))

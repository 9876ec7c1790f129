"""Measure the compiled variants of the SELL-32 apply kernel (apply.cu: U pairs in flight, register budget,
column prefetch) on tetgrid n: fused Jacobian apply alone and 200 MINRES iterations, CUDA events on the ctx
stream.  python profiles/apply_variants.py --n 200 > gpurun_out/apply_variants.json"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nosh_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200)
ap.add_argument("--reps", type=int, default=50)
a = ap.parse_args()
DESC = {0: "U=5, 64 regs, 2 CTAs/SM (default)", 1: "U=8, 64 regs", 2: "U=4 + column prefetch", 3: "U=8, 1 CTA/SM",
        4: "U=8, 1 CTA/SM + prefetch", 5: "U=6, 64 regs", 6: "U=4, 64 regs (the round-1 default)", 7: "U=12, 1 CTA/SM"}
ctx = nosh_b200.Context()
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
ctx.jac_rebuild(par, psi)
nb_blocks = int(mi.n_blocks)
bytes_apply = nb_blocks * 20 + (No + 1) * 8 + No * 56
out = []
ref = None
for v in range(8):
    ctx.set_tuning("apply_variant", v)
    for _ in range(3):
        ctx.jac_apply(b, x)
    ctx.timer_start()
    for _ in range(a.reps):
        ctx.jac_apply(b, x)
    ms = ctx.timer_stop() / a.reps
    y = x.clone()
    if ref is None:
        ref = y
    same = bool(torch.equal(ref, y))
    ctx.minres(b, x, tol=0.0, maxit=50)
    ctx.timer_start()
    ctx.minres(b, x, tol=0.0, maxit=200)
    ms_minres = ctx.timer_stop() / 200
    out.append({"variant": v, "what": DESC[v], "apply_ms": ms, "apply_gbs": bytes_apply / ms / 1e6,
                "minres_ms_per_iteration": ms_minres, "bit_identical_to_variant_0": same})
    print(out[-1], file=sys.stderr)
# the MINRES loop as one cooperative launch (krylov.cu k_minres_persistent) vs five launches per iteration
loop = {}
xs = {}
for mode, var in ((0, 0), (1, 0), (1, 6), (0, 0), (1, 0), (1, 6)):
    ctx.set_tuning("apply_variant", var)
    ctx.set_tuning("persistent_minres", mode)
    ctx.minres(b, x, tol=0.0, maxit=50)
    ctx.timer_start()
    ctx.minres(b, x, tol=0.0, maxit=400)
    key = ("persistent" if mode else "multi_launch") + ("_U4" if var == 6 else "")
    loop.setdefault(key, []).append(ctx.timer_stop() / 400)
    xs[key] = x.clone()
    print(mode, loop, file=sys.stderr)
loop["bits_equal"] = bool(all(torch.equal(v, xs["multi_launch"]) for v in xs.values()))
ctx.set_tuning("apply_variant", 0)
print(json.dumps({"n": a.n, "vertices": No, "bytes_per_apply": bytes_apply, "variants": out,
                  "minres_ms_per_iteration": loop}))

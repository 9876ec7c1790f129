"""Profiling driver for the plain fused Jacobian apply (nosh_jac_apply, the kernel bench.py's
roofline object times):
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:k_apply_sell -c 2 -o gpurun_out/prof_apply_plain python profiles/profile_plain.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nosh_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = nosh_b200.Context()
mi = ctx.mesh_tetgrid(n)
ctx.set_thickness(None, 1.0)
ctx.set_potential_constant(-1.0)
ctx.set_mvp_constcurl((0.0, 0.0, 1.0))
No = int(mi.n_owned)
g = torch.Generator(device="cuda")
g.manual_seed(1)
psi = torch.randn(2 * No, generator=g, device="cuda", dtype=torch.float64)
x = torch.randn(2 * No, generator=g, device="cuda", dtype=torch.float64)
y = torch.empty_like(x)
ctx.jac_rebuild({"g": 1.0, "mu": 1.0, "theta": 0.0}, psi)
for _ in range(3):
    ctx.jac_apply(x, y)
ctx.synchronize()
torch.cuda.profiler.start()
for _ in range(3):
    ctx.jac_apply(x, y)
ctx.synchronize()
torch.cuda.profiler.stop()
print("profiled 3 plain Jacobian applies, n=%d" % n)

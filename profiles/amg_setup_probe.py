"""AMG hierarchy set-up time in a FRESH process (the driver's `hierarchy_setup_seconds`): tetgrid n, psi = 1.
NOSH_B200_AMG_TIMING=1 prints the phases; NOSH_B200_AMG_ARENA_MB sets the size of the temporaries' arena.
    python profiles/amg_setup_probe.py [n]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nosh_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = nosh_b200.Context()
mi = ctx.mesh_tetgrid(n)
ctx.set_thickness(None, 1.0)
ctx.set_potential_constant(-1.0)
ctx.set_mvp_constcurl((0.0, 0.0, 1.0))
import torch  # noqa: E402
psi = torch.zeros(2 * int(mi.n_owned), device="cuda", dtype=torch.float64)
psi[0::2] = 1.0
torch.cuda.synchronize()
par = {"g": 1.0, "mu": 0.1, "theta": 0.0}
ctx.keoreg_rebuild(par, psi)
ctx.synchronize()
t0 = time.perf_counter()
ctx.amg_setup()
ctx.synchronize()
wall = time.perf_counter() - t0
ai = ctx.amg_info()
out = {"n": n, "setup_seconds": float(ai.setup_seconds), "wall_seconds": wall,
       "arena_mb_env": os.environ.get("NOSH_B200_AMG_ARENA_MB"), "levels": int(ai.levels)}
for k in ("amg.arena_bytes", "amg.arena_high_bytes", "amg.arena_fallback_bytes", "amg.setup.arena", "amg.store_bytes",
          "amg.setup.move to permanent storage", "amg.alloc.malloc_calls", "amg.alloc.malloc_s",
          "amg.alloc.free_calls", "amg.alloc.free_s"):
    try:
        out[k] = ctx.stat(k)
    except KeyError:
        pass
# a second build in the same process (reuse = none would do this per rebuild)
ctx.amg_set_options(reuse=0)
ctx.keoreg_rebuild(par, psi)
t0 = time.perf_counter()
ctx.amg_setup()
ctx.synchronize()
out["second_build_wall_seconds"] = time.perf_counter() - t0
out["second_build_setup_seconds"] = float(ctx.amg_info().setup_seconds)
print(json.dumps(out), flush=True)
ctx.close()

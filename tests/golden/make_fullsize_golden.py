"""Generates tests/golden/fullsize_n200.json: known-answer "hashes" (norms and quadratic forms, the style of
the reference's own tests: test/mesh.cpp, test/keo.cpp, test/compute_f.cpp, test/jac.cpp) of the CPU oracle
on the FULL-SIZE synthetic mesh of BASELINE.json (tetgrid n=200 = 8.0M vertices), so that the GPU path can be
checked at that size without running the oracle on the GPU box.  Takes a few minutes and ~25 GB of RAM:

    python tests/golden/make_fullsize_golden.py [n]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import meshgen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
t0 = time.time()
coords, cells = meshgen.tetgrid(n)
P = oracle.OracleProblem(coords, cells, ("constcurl", (0.0, 0.0, 1.0), None), nthreads=oracle.num_threads())
N = P.N
print("oracle problem built: N = %d, E = %d, %.0f s" % (N, P.E, time.time() - t0), flush=True)
g, mu = 1.0, 0.1
out = {"n": n, "num_nodes": N, "num_edges": int(P.E), "g": g, "mu": mu, "theta": 0.0,
       "fields": "psi = 1, V = -1, thickness = 1, constantCurl b = (0,0,1)",
       "cv_norm1": float(np.abs(P.cv).sum()), "cv_norm2": float(np.linalg.norm(P.cv)),
       "cv_norminf": float(np.abs(P.cv).max()), "cv_min": float(P.cv.min()),
       "alpha_sum": float(P.alpha.sum()), "alpha_norm2": float(np.linalg.norm(P.alpha))}
P.keo_fill(mu)
one = np.ones(2 * N)
er = np.zeros(2 * N)
er[0::2] = 1.0
ei = np.zeros(2 * N)
ei[1::2] = 1.0
x = meshgen.random_state(N, 42)
out["keo_1K1"] = float(one @ P.keo_apply(one))
out["keo_erKer"] = float(er @ P.keo_apply(er))
kx = P.keo_apply(x)
out["keo_xKx"] = float(x @ kx)
out["keo_Kx_norm2"] = float(np.linalg.norm(kx))
psi = er.copy()
F = P.compute_f(g, psi)
out["F_norm1"] = float(np.abs(F).sum())
out["F_norm2"] = float(np.linalg.norm(F))
out["F_norminf"] = float(np.abs(F).max())
Fx = P.compute_f(g, x)
out["F_random_state_norm2"] = float(np.linalg.norm(Fx))
P.jac_rebuild(g, x)
for name, s in (("one", one), ("er", er), ("ei", ei), ("x", x)):
    out["jac_%s" % name] = float(s @ P.jac_apply(s))
P.dkeo_fill(mu, 0.0, "mu")
out["dfdmu_norm2"] = float(np.linalg.norm(P.compute_dfdp(x, False, np.zeros(N))))
# entry-level evidence at full size: the values of K x, F(x), J y and dF/dmu at 4096 strided vertices (both
# components) -- each depends on the ~15 matrix blocks of its row, so a single wrong entry anywhere in those rows
# shows up here, which the norms above would average away
stride = max(1, N // 4096)
rows = np.arange(0, N, stride)[:4096]
idx = np.stack([2 * rows, 2 * rows + 1], 1).ravel()
y = meshgen.random_state(N, 43)
jy = P.jac_apply(y)
dfdmu = P.compute_dfdp(x, False, np.zeros(N))
out["sample_rows"] = {"stride": int(stride), "count": int(rows.size),
                      "Kx": [float(v) for v in kx[idx]], "Kx_max": float(np.abs(kx).max()),
                      "Fx": [float(v) for v in Fx[idx]], "Fx_max": float(np.abs(Fx).max()),
                      "Jy": [float(v) for v in jy[idx]], "Jy_max": float(np.abs(jy).max()),
                      "dFdmu": [float(v) for v in dfdmu[idx]], "dFdmu_max": float(np.abs(dfdmu).max()),
                      "note": "x = random_state(N, 42) (also the state J is built at), y = random_state(N, 43)"}
out["seconds"] = time.time() - t0
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_n%d.json" % n)
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))

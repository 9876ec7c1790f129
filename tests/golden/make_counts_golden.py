"""Generates tests/golden/counts_n<N>.json: MINRES and Newton ITERATION COUNTS (and residual-history
samples) of the CPU oracle on the synthetic benchmark meshes (tetgrid n=100 = 1.0M vertices = BASELINE.json
configs[1]; n=200 = 8.0M vertices = configs[2]), so that "identical iteration counts" can be checked on the GPU
box at benchmark size without running the oracle there (minutes to hours of CPU time).

Iteration counts of a Krylov method depend on the summation order of its dot products once the residual
is rounding dominated.  The reference sums per MPI rank and all-reduces, so its own counts depend on the rank
count (it tests with 1, 2 and 7 ranks, test/CMakeLists.txt:18-23).  The oracle is therefore run with its dot
products split into 1, 2, 7 and 16 contiguous parts (oracle.set_dot_parts); the spread between those runs is
the rounding sensitivity of the count and is stored next to it.

A second family of runs perturbs the OPERATOR APPLY at rounding level instead: every row of the sparse products
is summed right to left (oracle.set_row_reverse) -- what a different column order in Tpetra's local graph, or the
GPU's fused complex arithmetic, amounts to.  Because K x cancels strongly on smooth vectors this perturbation is
orders of magnitude larger than that of the dot products, and it is the one the GPU's deviation has to be
compared with.  Keys "<parts>r" in by_parts.

    python tests/golden/make_counts_golden.py [n] [parts,parts,...] [reverse parts,...]
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
parts_list = [int(v) for v in sys.argv[2].split(",") if v] if len(sys.argv) > 2 else [1, 2, 7, 16]
rev_list = [int(v) for v in sys.argv[3].split(",") if v] if len(sys.argv) > 3 else [1, 7]
runs = [(p, False) for p in parts_list] + [(p, True) for p in rev_list]
newton = os.environ.get("NOSH_GOLDEN_NEWTON", "1") != "0"
t0 = time.time()
coords, cells = meshgen.tetgrid(n)
P = oracle.OracleProblem(coords, cells, ("constcurl", (0.0, 0.0, 1.0), None), nthreads=oracle.num_threads())
N = P.N
print("oracle problem built: N = %d, %.0f s" % (N, time.time() - t0), flush=True)
out = {"n": n, "num_nodes": N, "threads": oracle.num_threads(),
       "minres": {"g": 1.0, "mu": 1.0, "psi": "random_state(N, 42)", "b": "random_state(N, 43)", "tol": 1e-10,
                  "maxit": 20000, "hist_at": [], "by_parts": {}},
       "newton": {"g": 1.0, "mu": 0.1, "psi0": "1", "nl_tol": 1e-8, "lin_tol": 1e-10, "lin_maxit": 20000,
                  "by_parts": {}}}
psi = meshgen.random_state(N, 42)
b = meshgen.random_state(N, 43)
P.keo_fill(1.0)
P.jac_rebuild(1.0, psi)
for parts, rev in runs:
    oracle.set_dot_parts(parts)
    oracle.set_row_reverse(rev)
    t = time.time()
    x, it, rr, hist = P.krylov(b, 1e-10, 20000, history=True)
    at = [k for k in (1, 2, 5, 10, 20, 50, 100, 200, 300, 400, 500, 750, 1000, 1500, 2000) if k <= it]
    out["minres"]["hist_at"] = at if len(at) > len(out["minres"]["hist_at"]) else out["minres"]["hist_at"]
    out["minres"]["by_parts"]["%d%s" % (parts, "r" if rev else "")] = {"iterations": it, "relres": rr, "hist": [float(hist[k]) for k in at],
                                             "x_norm2": float(np.linalg.norm(x)),
                                             "true_relres": float(np.linalg.norm(P.jac_apply(x) - b) / np.linalg.norm(b))}
    print("minres parts=%d reverse=%d: %d iterations, relres %.3e, %.0f s" % (parts, rev, it, rr, time.time() - t), flush=True)
if newton:
    psi0 = np.zeros(2 * N)
    psi0[0::2] = 1.0
    P.keo_fill(0.1)
    for parts, rev in runs:
        oracle.set_dot_parts(parts)
        oracle.set_row_reverse(rev)
        t = time.time()
        xn, steps, lin, fn = P.newton(1.0, psi0, 1e-8, 20, 1e-10, 20000)
        out["newton"]["by_parts"]["%d%s" % (parts, "r" if rev else "")] = {"steps": steps, "minres_iterations": [int(v) for v in lin],
                                                 "fnorms": [float(v) for v in fn],
                                                 "x_norm2": float(np.linalg.norm(xn))}
        print("newton parts=%d reverse=%d: %d steps, %s, %.0f s" % (parts, rev, steps, list(lin), time.time() - t), flush=True)
oracle.set_dot_parts(0)
oracle.set_row_reverse(False)
out["seconds"] = time.time() - t0
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "counts_n%d.json" % n)
if os.environ.get("NOSH_GOLDEN_MERGE") == "1" and os.path.exists(path):      # add runs to an existing file
    old = json.load(open(path))
    for sec in ("minres", "newton"):
        old[sec]["by_parts"].update(out[sec]["by_parts"])
    if len(out["minres"]["hist_at"]) > len(old["minres"]["hist_at"]):
        old["minres"]["hist_at"] = out["minres"]["hist_at"]
    out = old
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)

"""Robustness of the SELL-32 apply kernel against the vertex numbering and the row-length distribution
(SURVEY.md 7.3 "gather locality").  All meshes have the same vertices (tetgrid n):

  kuhn/natural      the benchmark mesh: Kuhn split, lexicographic ids (device generator)
  kuhn/shuffled     same cells, vertex ids randomly permuted (what an unordered mesh file would give)
  kuhn/morton       the shuffled mesh renumbered with nosh_morton_order (the stand-in for mbpart / RCM)
  five/natural      alternating 5-tet split: block rows of 7 and 19 entries alternate (varying valence)
  five/natural+sigma  the same with the rows of every 512-row window sorted by length (SELL-32-sigma)
  five/morton(+sigma) shuffled + Morton renumbered

For each: n_stored / n_blocks (padding), fused Jacobian apply ms and algorithmic GB/s (CUDA events on the ctx
stream, 50 applies), and ms per iteration of a 50-iteration MINRES run.

    python profiles/unstructured_bench.py --n 200 > gpurun_out/unstructured.json
    python profiles/unstructured_bench.py --n 200 --only kuhn/morton --reps 5      (under ncu)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nosh_b200  # noqa: E402
from oracle import meshgen  # noqa: E402   (numpy mesh generators only: nothing of the oracle is timed)

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200)
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--only", default=None)
a = ap.parse_args()
PEAK = 6452.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def measure(name, build, sigma):
    if a.only and a.only != name:
        return None
    t0 = time.time()
    ctx = nosh_b200.Context()
    ctx.set_tuning("sell_sigma", sigma)
    mi = build(ctx)
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
    nb = int(mi.n_blocks)
    bytes_apply = nb * 20 + (No + 1) * 8 + No * 56
    for _ in range(3):
        ctx.jac_apply(b, x)
    ctx.synchronize()
    ctx.timer_start()
    for _ in range(a.reps):
        ctx.jac_apply(b, x)
    ms = ctx.timer_stop() / a.reps
    ctx.minres(b, x, tol=0.0, maxit=10)
    ctx.synchronize()
    ctx.timer_start()
    ctx.minres(b, x, tol=0.0, maxit=50)
    ms_it = ctx.timer_stop() / 50
    r = {"mesh": name, "n_vertices": No, "n_blocks": nb, "n_stored": int(mi.n_stored),
         "stored_over_blocks": int(mi.n_stored) / nb, "sell_sigma": ctx.stat("sell.sigma"),
         "apply_ms": ms, "apply_algorithmic_gbs": bytes_apply / ms / 1e6, "frac_of_measured_peak": bytes_apply / ms / 1e6 / PEAK,
         "stored_gbs": (int(mi.n_stored) * 20 + (No + 1) * 8 + No * 56) / ms / 1e6,
         "minres_ms_per_iteration": ms_it, "setup_s": time.time() - t0}
    ctx.close()
    print(json.dumps(r), flush=True)
    return r


def host_mesh(kind, order):
    coords, cells = (meshgen.tetgrid if kind == "kuhn" else meshgen.tetgrid5)(a.n)
    if order == "natural":
        return coords, cells
    rng = np.random.default_rng(7)
    perm = rng.permutation(coords.shape[0])
    coords, cells, _ = nosh_b200.api.renumber(coords, cells, perm)
    if order == "morton":
        perm = nosh_b200.api.morton_order(coords)
        coords, cells, _ = nosh_b200.api.renumber(coords, cells, perm)
    return coords, cells


def from_host(kind, order):
    def build(ctx):
        coords, cells = host_mesh(kind, order)
        return ctx.mesh_set(coords, cells)
    return build


measure("kuhn/natural", lambda ctx: ctx.mesh_tetgrid(a.n), 0)
measure("kuhn/shuffled", from_host("kuhn", "shuffled"), -1)
measure("kuhn/morton", from_host("kuhn", "morton"), -1)
measure("five/natural", from_host("five", "natural"), 0)
measure("five/natural+sigma", from_host("five", "natural"), 1)
measure("five/morton", from_host("five", "morton"), 0)
measure("five/morton+sigma", from_host("five", "morton"), 1)

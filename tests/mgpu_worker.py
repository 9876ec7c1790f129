"""Multi-rank parity worker (launched by torchrun, one rank per GPU -- or, with NOSH_TEST_COMM=host, any
number of ranks SHARING the available GPUs; see test_multi_gpu.py).

Checks, on a vertex-partitioned tetgrid with peer-memory halo exchange and group-sum all-gathers:
  * every rank's owned slice of cv / F / J.x / dF/dp matches the oracle on the GLOBAL mesh
  * MINRES / CG / Newton iteration counts equal the oracle's
  * partition independence: results are BIT-IDENTICAL to a single-GPU context on the same mesh, for the
    persistent one-launch MINRES loop and for the multi-launch loop, and for whole continuation runs

Environment:
  NOSH_TEST_COMM      nccl (default): torch NCCL process group + nosh_ctx_comm_init (library-owned NCCL
                      communicator for set-up, CUDA IPC peer memory for the data path)
                      host: torch gloo process group + nosh_ctx_comm_init_host (set-up through the caller's
                      communicator, no NCCL anywhere; ranks may share a GPU)
  NOSH_TEST_N         grid size (default 20)
  NOSH_TEST_SECTIONS  comma list of core,local,amg,gmres,cont,tiny (default: all)
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nosh_b200  # noqa: E402
from oracle import OracleProblem, meshgen  # noqa: E402

RTOL = 1e-12


def relerr(a, b):
    s = np.abs(b).max()
    return np.abs(a - b).max() / (s if s > 0 else 1.0)


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    mode = os.environ.get("NOSH_TEST_COMM", "nccl")
    sections = set(os.environ.get("NOSH_TEST_SECTIONS", "core,local,amg,gmres,cont,tiny").split(","))
    dev = local % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    if mode == "host":
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    n = int(os.environ.get("NOSH_TEST_N", "20"))
    group = 512
    par = {"g": 1.0, "mu": 0.3, "theta": 0.0}
    t_start = time.time()

    def log(msg):
        if rank == 0:
            print("[mgpu %6.1fs] %s" % (time.time() - t_start, msg), flush=True)

    def new_ctx(gv=group):
        c = nosh_b200.Context(device=dev, group_vertices=gv)
        if mode == "host":
            c.comm_init_torch()
        else:
            obj = [nosh_b200.Context.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            c.comm_init(obj[0], rank, world)
        return c

    def fields(c):
        c.set_thickness(None, 1.0)
        c.set_potential_constant(-1.0)
        c.set_mvp_constcurl((0.0, 0.0, 1.0))

    ctx = new_ctx()
    mi = ctx.mesh_tetgrid(n, n, n + 3)
    fields(ctx)
    p2p = ctx.stat("p2p") == 1.0
    if mode == "host" or os.environ.get("NOSH_B200_P2P", "1") != "0":
        assert p2p, "peer memory (CUDA IPC) path is not active"
    vb, No = int(mi.owned_begin), int(mi.n_owned)
    sl = slice(2 * vb, 2 * (vb + No))
    assert mi.n_ghost > 0 and No > 0, (mi.n_ghost, No)

    coords, cells = meshgen.tetgrid(n, n, n + 3)
    N = coords.shape[0]
    assert mi.n_global == N
    P = OracleProblem(coords, cells, ("constcurl", (0.0, 0.0, 1.0), None), nthreads=2)

    # one-GPU context on the same mesh (no communicator): the bit-identity reference
    single = nosh_b200.Context(device=dev, group_vertices=group)
    single.mesh_tetgrid(n, n, n + 3)
    fields(single)

    x = meshgen.random_state(N, 42)
    y = meshgen.random_state(N, 43)
    b = meshgen.random_state(N, 4)
    summary = {}

    if "core" in sections:
        gids = ctx.local_gids()
        assert np.array_equal(gids[:No], np.arange(vb, vb + No))
        assert np.array_equal(ctx.coords(), coords[gids])
        assert relerr(ctx.control_volumes(), P.cv[vb:vb + No]) <= RTOL
        # local edges <-> global edges
        e, ln, cov = ctx.edges()
        ge = gids[e]
        key = ge[:, 0] * N + ge[:, 1]
        okey = P.edges[:, 0].astype(np.int64) * N + P.edges[:, 1]
        pos = np.searchsorted(okey, key)
        assert np.array_equal(okey[pos], key)
        assert relerr(cov, P.covolume[pos]) <= RTOL and relerr(ctx.alpha_cache(), P.alpha[pos]) <= RTOL

        P.keo_fill(par["mu"])
        P.jac_rebuild(par["g"], x)
        ctx.keo_fill(par)
        ctx.jac_rebuild(par, x[sl].copy())
        F = ctx.compute_f(par, x[sl].copy())
        Jy = ctx.jac_apply(y[sl].copy())
        assert relerr(F, P.compute_f(par["g"], x)[sl]) <= RTOL
        assert relerr(Jy, P.jac_apply(y)[sl]) <= RTOL
        # device vectors of the caller are used in place (no ghost room, no staging copy with peer memory)
        yd = torch.from_numpy(y[sl].copy()).cuda()
        od = torch.empty_like(yd)
        ctx.jac_apply(yd, od)
        ctx.synchronize()
        assert np.array_equal(od.cpu().numpy(), Jy)
        Ky = ctx.keo_apply(y[sl].copy(), alpha=2.0, beta=0.0)
        assert relerr(Ky, 2.0 * P.keo_apply(y)[sl]) <= RTOL
        P.dkeo_fill(par["mu"], 0.0, "mu")
        dF = ctx.compute_dfdp(par, "mu", x[sl].copy())
        assert relerr(dF, P.compute_dfdp(x, False, np.zeros(N))[sl]) <= RTOL
        d = ctx.dot(x[sl].copy(), y[sl].copy())
        assert abs(d - x @ y) <= 1e-12 * abs(x @ y)
        log("partitioned F / J.x / dF/dmu / dot == oracle")

        xo, ito, _ = P.krylov(b, 1e-10, 3000)
        xo2, ito2, _ = P.krylov(b, 1e-8, 3000, solver="cg", jacobian=False)
        single.keo_fill(par)
        single.jac_rebuild(par, x)
        assert np.array_equal(single.compute_f(par, x)[sl], F)
        assert np.array_equal(single.jac_apply(y)[sl], Jy)
        assert np.array_equal(single.compute_dfdp(par, "mu", x)[sl], dF)
        assert single.dot(x, y) == d
        xs, rs, hs = single.minres(b, tol=1e-10, maxit=3000, history=True)
        launches = {}
        for persistent, lean in ((1, 0), (1, 1), (0, 0)):      # lean: the two-grid-syncs schedule of the persistent kernel
            ctx.set_tuning("persistent_mgpu", persistent)
            ctx.set_tuning("mgpu_lean", lean)
            l0 = ctx.launch_count()
            xg, res, hg = ctx.minres(b[sl].copy(), tol=1e-10, maxit=3000, history=True)
            launches[persistent] = ctx.launch_count() - l0
            assert res.iterations == ito and res.converged == 1, (persistent, res.iterations, ito)
            assert relerr(xg, xo[sl]) <= 1e-8
            assert rs.iterations == res.iterations
            assert np.array_equal(hs, hg), "residual history differs between 1 and %d ranks (persistent=%d)" % (
                world, persistent)
            assert np.array_equal(xs[sl], xg)
            # a solve cut off by maxit, and the KEO operator
            xc, rc = ctx.minres(b[sl].copy(), tol=1e-10, maxit=17)
            xsc, rsc = single.minres(b, tol=1e-10, maxit=17)
            assert rc.iterations == rsc.iterations == 17 and np.array_equal(xsc[sl], xc)
            xk, rk = ctx.minres(b[sl].copy(), tol=1e-8, maxit=3000, op=nosh_b200.OP_KEO)
            xsk, rsk = single.minres(b, tol=1e-8, maxit=3000, op=nosh_b200.OP_KEO)
            assert rk.iterations == rsk.iterations and np.array_equal(xsk[sl], xk)
        ctx.set_tuning("mgpu_lean", 0)
        if p2p:
            assert launches[1] <= 10 < launches[0], launches   # one cooperative launch per solve and rank
        summary["minres"] = ito
        summary["launches_persistent"] = launches[1]
        summary["launches_multi"] = launches[0]
        log("MINRES %d iterations == oracle, bits == one GPU, launches persistent/multi = %s" % (ito, launches))
        xg2, res2 = ctx.cg(b[sl].copy(), op=nosh_b200.OP_KEO, tol=1e-8, maxit=3000)
        assert res2.iterations == ito2, (res2.iterations, ito2)
        xs2, rs2 = single.cg(b, op=nosh_b200.OP_KEO, tol=1e-8, maxit=3000)
        assert np.array_equal(xs2[sl], xg2)

        psi0, _ = meshgen.plain_gl_fields(coords)
        parn = {"g": 1.0, "mu": 0.1, "theta": 0.0}
        P.keo_fill(parn["mu"])
        xn, steps, lin, fn = P.newton(1.0, psi0, 1e-8, 20, 1e-10, 3000)
        psi = psi0[sl].copy()
        nres, glin, gfn = ctx.newton(parn, psi, 1e-8, 20, 1e-10, 3000)
        # the last Newton step solves for a correction below the rounding floor of its right-hand side: its
        # MINRES count is rounding noise in the oracle itself (139/162/203/149 vs .../242 on n = 12)
        assert nres.steps == steps and list(glin)[:-1] == list(lin)[:-1], (nres.steps, steps, glin, lin)
        assert nres.linear_solve_status == 0 and nres.converged == 1
        assert relerr(psi, xn[sl]) <= 1e-8
        psis = psi0.copy()
        single.newton(parn, psis, 1e-8, 20, 1e-10, 3000)
        assert np.array_equal(psis[sl], psi)
        summary["newton"] = [int(v) for v in lin]
        log("Newton %s == oracle, bits == one GPU" % list(lin))

    if "local" in sections:
        # ---- partitioned ingestion (MOAB's READ_PART, src/mesh_reader.cpp:32-35): every rank passes only the
        # cells touching its owned vertex range + their vertices; identical to uploading the global mesh
        cg = new_ctx()
        mg = cg.mesh_set(coords, cells)
        cl = new_ctx()
        b0, e0, _ = nosh_b200.partition_range(N, world, rank, group)
        gids_l, coords_l, cells_l = nosh_b200.Context.local_part(coords, cells, b0, e0)
        perm = np.random.default_rng(rank).permutation(gids_l.size)       # any order of the local vertex list
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.size)
        ml = cl.mesh_set_local(N, gids_l[perm], coords_l[perm], inv[cells_l].astype(np.int32))
        assert gids_l.size < N                                         # really only a part
        for f in ("n_global", "owned_begin", "n_owned", "n_ghost", "n_cells", "n_edges", "n_blocks", "n_stored"):
            assert getattr(mg, f) == getattr(ml, f) == getattr(mi, f), f
        assert np.array_equal(cg.local_gids(), cl.local_gids()) and np.array_equal(cg.coords(), cl.coords())
        assert np.array_equal(cg.control_volumes(), cl.control_volumes())
        outs = []
        for c in (cg, cl):
            fields(c)
            c.keo_fill(par)
            outs.append((c.block_csr(), c.compute_f(par, x[sl].copy())))
        for a_, b_ in zip(outs[0][0], outs[1][0]):
            assert np.array_equal(a_, b_)
        assert np.array_equal(outs[0][1], outs[1][1])
        with np.testing.assert_raises(RuntimeError):                     # a global id outside [0, N)
            bad = new_ctx()
            g2 = gids_l.copy()
            g2[0] = N + 5
            bad.mesh_set_local(N, g2, coords_l, cells_l)
        cg.close()
        cl.close()
        log("partitioned ingestion (mesh_set_local) == global upload, bit for bit")

    if "amg" in sections:
        # ---- preconditioned MINRES: every rank applies the AMG V-cycle of ITS diagonal block of the
        # regularised KEO (no communication inside the preconditioner); the oracle does the same with one
        # hierarchy per rank's vertex range
        import scipy.sparse as sp
        from oracle import amg
        _, ito, _ = P.krylov(b, 1e-10, 3000) if "minres" not in summary else (None, summary["minres"], None)
        ctx.amg_set_options(coarse_max=64)
        ctx.keo_fill(par)
        ctx.jac_rebuild(par, x[sl].copy())
        ctx.keoreg_rebuild(par, x[sl].copy())
        P.keo_fill(par["mu"])
        P.jac_rebuild(par["g"], x)
        K = sp.csr_matrix((P.vals.copy(), P.cols, P.rowptr), shape=(2 * N, 2 * N))
        rr = np.arange(N)
        Dj = sp.csr_matrix((np.concatenate([P.d0[0::2], P.d0[1::2], P.d1b, P.d1b]),
                            (np.concatenate([2 * rr, 2 * rr + 1, 2 * rr, 2 * rr + 1]),
                             np.concatenate([2 * rr, 2 * rr + 1, 2 * rr + 1, 2 * rr]))), shape=(2 * N, 2 * N))
        J = (K + Dj).tocsr()
        Pm = sp.csr_matrix((P.keoreg_fill(par["mu"], par["g"], x), P.cols, P.rowptr), shape=(2 * N, 2 * N))
        blocks = []
        for r in range(world):
            b0, e0, _ = nosh_b200.partition_range(N, world, r, group)
            blocks.append((b0, e0, amg.Hierarchy(Pm[2 * b0:2 * e0, 2 * b0:2 * e0], coarse_max=64, degree=1)))

        def M(v):
            out = np.empty_like(v)
            for b0, e0, H in blocks:
                out[2 * b0:2 * e0] = H.vcycle(v[2 * b0:2 * e0])
            return out

        zg = ctx.keoreg_apply(y[sl].copy())
        assert relerr(zg, M(y)[sl]) <= 1e-11
        xpo, itpo, _, _ = amg.pminres(lambda t: J @ t, M, b, 1e-10, 1000)
        xpg, pres = ctx.minres(b[sl].copy(), tol=1e-10, maxit=1000, prec=nosh_b200.PREC_KEOREG_AMG)
        assert pres.iterations == itpo and pres.converged == 1, (pres.iterations, itpo)
        assert relerr(xpg, xpo[sl]) <= 1e-8
        assert itpo < ito / 2
        summary["pminres"] = itpo
        log("AMG-preconditioned MINRES %d iterations == oracle" % itpo)

    if "gmres" in sections:
        # restarted GMRES: batched Gram-Schmidt reductions through the same tree + one sum of the group sums
        from oracle import gmres as og
        P.keo_fill(par["mu"])
        P.jac_rebuild(par["g"], x)
        single.keo_fill(par)
        single.jac_rebuild(par, x)
        ctx.keo_fill(par)
        ctx.jac_rebuild(par, x[sl].copy())
        xs3, rs3, hs3 = single.gmres(b, tol=1e-10, maxit=400, restart=40, history=True)
        xg3, rg3, hg3 = ctx.gmres(b[sl].copy(), tol=1e-10, maxit=400, restart=40, history=True)
        assert rs3.iterations == rg3.iterations and rg3.converged == 1
        assert np.array_equal(hs3, hg3) and np.array_equal(xs3[sl], xg3)
        _, ito3, _, _ = og.gmres(lambda t: P.jac_apply(t), None, b, 1e-10, 400, restart=40)
        assert rg3.iterations == ito3, (rg3.iterations, ito3)
        summary["gmres"] = ito3
        log("GMRES(40) %d iterations == oracle, bits == one GPU" % ito3)

    if "cont" in sections:
        # ---- configs[3]: parameter continuation on several ranks.  Every reduction is partition independent,
        # so whole runs (step records AND solutions) must equal the one-GPU run bit for bit; against the
        # oracle: same parameters / Newton steps / energies (the last corrector solves are rounding dominated,
        # see test_gpu_parity.py, so their MINRES counts are compared within a band)
        from oracle import continuation as oc
        psi0, _ = meshgen.plain_gl_fields(coords)
        p0 = {"g": 1.0, "mu": 0.0, "theta": 0.0}
        for c in (ctx, single):
            c.set_preconditioner(nosh_b200.PREC_NONE)
        xo, recs = P.continuation(1.0, "mu", 0.0, 0.05, 3, psi0, lin_maxit=3000)
        xg = psi0[sl].copy()
        st_m = ctx.continuation(p0, "mu", 0.05, 3, xg, lin_maxit=3000)
        xs1 = psi0.copy()
        st_s = single.continuation(p0, "mu", 0.05, 3, xs1, lin_maxit=3000)
        assert len(st_m) == len(st_s) == len(recs) == 4
        for a, s, r in zip(st_m, st_s, recs):
            assert (a.step, a.converged, a.newton_steps, a.linear_iterations, a.predictor_linear_iterations) == \
                   (s.step, s.converged, s.newton_steps, s.linear_iterations, s.predictor_linear_iterations)
            assert a.param == s.param and a.gibbs_energy == s.gibbs_energy and a.norm == s.norm and a.fnorm == s.fnorm
            assert a.converged == 1 and a.newton_steps == r["newton_steps"]
            assert abs(a.param - r["param"]) <= 1e-15
            assert abs(a.linear_iterations - r["linear_iterations"]) <= 0.25 * max(1, r["linear_iterations"])
            assert abs(a.gibbs_energy - r["gibbs_energy"]) <= 1e-9 * abs(r["gibbs_energy"])
            assert abs(a.norm - r["norm"]) <= 1e-9 * abs(r["norm"])
        assert np.array_equal(xs1[sl], xg)
        assert relerr(xg, xo[sl]) <= 1e-6
        summary["continuation_minres"] = [int(a.linear_iterations) for a in st_m]
        log("natural continuation: records and solution bits == one GPU; oracle agrees (%s MINRES its/step)"
            % summary["continuation_minres"])
        kw = dict(initial_step_size=0.05, min_step_size=1e-7, max_step_size=0.1, aggressiveness=2.0, max_steps=3,
                  lin_maxit=3000)
        xo, recs = oc.arclength(P, 1.0, 0.0, psi0, kw["initial_step_size"], kw["min_step_size"], kw["max_step_size"],
                                kw["aggressiveness"], kw["max_steps"], lin_maxit=3000)
        xg = psi0[sl].copy()
        al_m = ctx.continuation_arclength(p0, "mu", xg, **kw)
        xs1 = psi0.copy()
        al_s = single.continuation_arclength(p0, "mu", xs1, **kw)
        assert len(al_m) == len(al_s) == len(recs)
        for a, s, r in zip(al_m, al_s, recs):
            for f in ("step", "converged", "newton_steps", "linear_iterations", "predictor_linear_iterations", "param",
                      "gibbs_energy", "norm", "fnorm", "step_size", "dparam_ds"):
                assert getattr(a, f) == getattr(s, f), (f, getattr(a, f), getattr(s, f))
            assert a.converged == 1 and a.newton_steps == r["newton_steps"]
            assert abs(a.param - r["param"]) <= 1e-7 * max(1.0, abs(r["param"]))
            assert abs(a.step_size - r["step_size"]) <= 1e-12 * max(1.0, abs(r["step_size"]))
            assert abs(a.gibbs_energy - r["gibbs_energy"]) <= 1e-7 * abs(r["gibbs_energy"])
        assert np.array_equal(xs1[sl], xg)
        assert relerr(xg, xo[sl]) <= 1e-5
        summary["arclength_mu"] = [float(a.param) for a in al_m]
        log("arc-length continuation: records and solution bits == one GPU; oracle agrees (mu = %s)"
            % summary["arclength_mu"])
        # the same sweep with the per-rank AMG V-cycle as preconditioner (block-Jacobi over ranks): counts
        # depend on the rank count by construction, the branch does not
        ctx.amg_set_options(coarse_max=64)
        ctx.set_preconditioner(nosh_b200.PREC_KEOREG_AMG)
        xg2 = psi0[sl].copy()
        st_p = ctx.continuation(p0, "mu", 0.05, 3, xg2, lin_maxit=3000)
        ctx.set_preconditioner(nosh_b200.PREC_NONE)
        assert [a.converged for a in st_p] == [1, 1, 1, 1]
        for a, s in zip(st_p, st_s):
            assert abs(a.gibbs_energy - s.gibbs_energy) <= 1e-8 * abs(s.gibbs_energy)
        assert sum(a.linear_iterations for a in st_p) < sum(a.linear_iterations for a in st_s) / 2
        log("AMG-preconditioned continuation follows the same branch")

    single.close()
    ctx.close()

    if "tiny" in sections:
        # ---- a mesh so small that some ranks own nothing (1 group of 512 vertices for 5^3 = 125) ----
        tiny = new_ctx(512)
        mt = tiny.mesh_tetgrid(5)
        owner = world - 1                      # one group only: floor(1*r/P) puts it on the last rank
        assert (mt.n_owned == 125) == (rank == owner) and (mt.n_owned == 0) == (rank != owner)
        fields(tiny)
        c5, t5 = meshgen.tetgrid(5)
        P5 = OracleProblem(c5, t5, ("constcurl", (0.0, 0.0, 1.0), None))
        x5 = meshgen.random_state(125, 1)
        b5 = meshgen.random_state(125, 2)
        s5 = slice(0, 250) if rank == owner else slice(0, 0)
        P5.keo_fill(par["mu"])
        P5.jac_rebuild(par["g"], x5)
        tiny.jac_rebuild(par, x5[s5].copy())
        f5 = tiny.compute_f(par, x5[s5].copy())
        if rank == owner:
            assert relerr(f5, P5.compute_f(par["g"], x5)) <= RTOL
        assert abs(tiny.dot(x5[s5].copy(), b5[s5].copy()) - x5 @ b5) <= 1e-12 * abs(x5 @ b5)
        # 250 unknowns: the Lanczos process amplifies rounding differences quickly on this matrix (the
        # oracle run with 1 and 8 summation threads already differs by 1e-4 after 40 iterations), so
        # compare a fixed, short run instead of a converged one
        xo5, it5, _ = P5.krylov(b5, 1e-14, 20)
        xg5, r5 = tiny.minres(b5[s5].copy(), tol=1e-14, maxit=20)
        assert r5.iterations == it5 == 20, (r5.iterations, it5)
        if rank == owner:
            ro = np.linalg.norm(P5.jac_apply(xo5) - b5)
            rg = np.linalg.norm(P5.jac_apply(xg5) - b5)
            assert relerr(xg5, xo5) <= 1e-10 and abs(rg - ro) <= 1e-10 * ro, (relerr(xg5, xo5), rg, ro)
        tiny.close()
        log("ranks that own nothing: ok")
    dist.barrier()
    if rank == 0:
        print("MGPU OK world=%d n=%d comm=%s p2p=%d %s" % (world, n, mode, int(p2p), summary), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

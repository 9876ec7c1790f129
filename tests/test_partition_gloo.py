"""CPU, world_size 2 and 3 over gloo: the host-side logic of the multi-GPU path.

What can run without a GPU is the *plan*: (1) the group-aligned contiguous partition the library
computes (nosh_partition_range, a host-only C-ABI call), (2) the ghost/halo exchange plan derived
from it, (3) the fixed three-level reduction tree.  Each rank owns its rows of the ORACLE's global
Jacobian, exchanges halo entries of x with gloo point-to-point messages exactly as halo_setup /
halo_exchange do over NCCL (ghosts sorted by global id => one contiguous run per owner), applies its
rows, and reduces dots with the tree of common.cuh.  Asserted: the partition tiles the mesh, the
distributed apply equals the global one bit for bit, and dot products are bit-identical for 1, 2
and 3 ranks (partition independence).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHUNK = 512


def _halve(t):
    """Lane 0 of an xor-shuffle reduction (offsets n/2 ... 1): repeated t[:h] + t[h:]."""
    t = np.asarray(t, np.float64).copy()
    while t.size > 1:
        h = t.size // 2
        t = t[:h] + t[h:]
    return t[0]


def tree_dot(x, y, vb, n_global, group, allreduce):
    """The reduction tree of csrc/common.cuh in numpy, for the owned range starting at vb."""
    No = x.size // 2
    assert -(-n_global // group) <= 1024
    prod = (x * y).reshape(-1, 2).sum(1)            # per-vertex contribution re*re + im*im
    gs = np.zeros(1024)
    cpg = group // CHUNK
    for g0 in range(0, No, group):                  # the owned range is group aligned
        part = np.zeros(cpg)
        for c in range(cpg):
            seg = prod[g0 + c * CHUNK: g0 + (c + 1) * CHUNK]
            v = np.zeros(CHUNK)
            v[:seg.size] = seg
            t = v[:256] + v[256:]                   # level 1: thread t owns vertices t, t+256
            warps = np.array([_halve(t[32 * w:32 * w + 32]) for w in range(8)])
            s = 0.0
            for w in warps:                         # warp sums added in warp order
                s += w
            part[c] = s
        per = -(-cpg // 32)                         # level 2: lane l sums `per` consecutive partials
        lanes = np.zeros(32)
        for l in range(32):
            for k in range(l * per, min((l + 1) * per, cpg)):
                lanes[l] += part[k]
        gs[(vb + g0) // group] = _halve(lanes)
    gs = allreduce(gs)                              # exact: every entry has exactly one owner
    ws = np.array([_halve(gs[32 * w:32 * w + 32]) for w in range(32)])   # level 3
    return _halve(ws)


def _worker(rank, world, port, n, group, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import nosh_b200
    from oracle import OracleProblem, meshgen
    import scipy.sparse as sp

    coords, cells = meshgen.tetgrid(n)
    N = coords.shape[0]
    P = OracleProblem(coords, cells, ("constcurl", (0.0, 0.0, 1.0), None))
    P.keo_fill(0.7)
    x = meshgen.random_state(N, 42)
    y = meshgen.random_state(N, 43)
    P.jac_rebuild(1.0, x)
    J_ref = P.jac_apply(y)

    vb, ve, G = nosh_b200.partition_range(N, world, rank, group)
    # (1) the ranges tile [0, N) and are group aligned
    rng = [None] * world
    dist.all_gather_object(rng, (vb, ve))
    assert rng[0][0] == 0 and rng[-1][1] == N
    for a, b in zip(rng[:-1], rng[1:]):
        assert a[1] == b[0] and a[1] % G == 0
    # (2) halo plan: ghosts = columns of my rows outside my range, sorted by global id
    K = sp.csr_matrix((P.vals, P.cols, P.rowptr), shape=(2 * N, 2 * N))
    rows = K[2 * vb:2 * ve]
    vcols = np.unique(rows.indices // 2)
    ghosts = vcols[(vcols < vb) | (vcols >= ve)]
    owner = np.searchsorted(np.array([r[1] for r in rng]), ghosts, side="right")
    assert np.all(np.diff(owner) >= 0)              # one contiguous run per owner
    want = [ghosts[owner == r] for r in range(world)]
    allwant = [None] * world
    dist.all_gather_object(allwant, want)
    send_idx = [allwant[r][rank] for r in range(world)]   # what rank r wants from me
    ylocal = np.zeros(2 * N)
    ylocal[2 * vb:2 * ve] = y[2 * vb:2 * ve]
    reqs = []
    recv_bufs = {}
    for r in range(world):
        if r == rank:
            continue
        if len(send_idx[r]):
            idx = np.asarray(send_idx[r])
            buf = torch.from_numpy(np.stack([y[2 * idx], y[2 * idx + 1]], 1).copy())
            reqs.append(dist.isend(buf, r))
        if len(want[r]):
            recv_bufs[r] = torch.empty(len(want[r]), 2, dtype=torch.float64)
            reqs.append(dist.irecv(recv_bufs[r], r))
    for q in reqs:
        q.wait()
    for r, buf in recv_bufs.items():
        ylocal[2 * want[r]] = buf[:, 0].numpy()
        ylocal[2 * want[r] + 1] = buf[:, 1].numpy()
    Jloc = rows @ ylocal
    k = np.arange(vb, ve)
    Jloc[0::2] += P.d0[2 * k] * ylocal[2 * k] + P.d1b[k] * ylocal[2 * k + 1]
    Jloc[1::2] += P.d1b[k] * ylocal[2 * k] + P.d0[2 * k + 1] * ylocal[2 * k + 1]
    if ve > vb:   # with 7 ranks and 5 groups two ranks own nothing (the reference's -n 7 runs on 82..744 vertices)
        assert np.abs(Jloc - J_ref[2 * vb:2 * ve]).max() <= 1e-13 * np.abs(J_ref).max()

    # (3) partition-independent reduction
    def allreduce(a):
        t = torch.from_numpy(a.copy())
        dist.all_reduce(t)
        return t.numpy()

    d = tree_dot(x[2 * vb:2 * ve], y[2 * vb:2 * ve], vb, N, G, allreduce)
    if rank == 0:
        out.put((world, float(d), float(x @ y)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_partition_halo_and_reduction_tree_over_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    results = {}
    n, group = 13, 512                       # 2197 vertices -> 5 groups of 512
    # 1, 2, 3 and 7 ranks: the reference runs every test serial, with mpiexec -n 2 and -n 7 (test/CMakeLists.txt:18-23)
    for world, port in ((1, 29701), (2, 29702), (3, 29703), (7, 29707)):
        procs = [ctx.Process(target=_worker, args=(r, world, port, n, group, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        w, d, ref = out.get(timeout=10)
        results[w] = d
        assert abs(d - ref) <= 1e-13 * abs(ref)
    assert results[1] == results[2] == results[3] == results[7], results   # bit-identical for any rank count


def test_partition_range_properties():
    sys.path.insert(0, ROOT)
    import nosh_b200
    for N, P, g in ((8_000_000, 8, 65536), (64_000_000, 8, 65536), (1_000_000, 4, 65536), (2197, 3, 512),
                    (128_000_000, 8, 65536), (409, 7, 512), (8_000_000, 7, 65536)):
        prev = 0
        for r in range(P):
            b, e, G = nosh_b200.partition_range(N, P, r, g)
            assert b == prev and e >= b and (e % G == 0 or e == N)
            assert -(-N // G) <= 1024 and G % g == 0
            prev = e
        assert prev == N
    with pytest.raises(ValueError):
        nosh_b200.partition_range(100, 2, 2, 512)
    with pytest.raises(ValueError):
        nosh_b200.partition_range(100, 2, 0, 100)

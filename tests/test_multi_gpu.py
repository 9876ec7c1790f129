"""Several ranks: run tests/mgpu_worker.py under torchrun.

Two launch modes:
  * one rank per GPU, NCCL process group (needs >= 2 devices: skipped on a one-GPU box)
  * SHARED-GPU ranks with the host communicator (nosh_ctx_comm_init_host over a gloo group): 2 and 3
    processes on whatever GPUs exist -- the whole multi-rank path (partitioned mesh, CUDA-IPC peer memory,
    in-kernel halo push and group-sum all-gather, the persistent multi-rank MINRES kernel, partition-
    independent bits) runs on a ONE-GPU box too; the ranks' kernels time-slice on the device.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(world, env_extra, port, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0 and "MGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_parity_and_bit_identity(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    _run(world, {"NOSH_TEST_COMM": "nccl"}, 29611 + world)


@pytest.mark.parametrize("world", [2, 4])
def test_host_communicator_one_rank_per_gpu(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    _run(world, {"NOSH_TEST_COMM": "host", "NOSH_TEST_SECTIONS": "core,local,cont"}, 29631 + world)


@pytest.mark.parametrize("world", [2, 3])
def test_shared_gpu_ranks_host_communicator(world):
    """Runs on ONE GPU: `world` processes share it (and any further GPUs round-robin)."""
    out = _run(world, {"NOSH_TEST_COMM": "host", "NOSH_TEST_N": "12",
                       "NOSH_TEST_SECTIONS": os.environ.get("NOSH_SHARED_SECTIONS", "core,local,cont,tiny")},
               29651 + world)
    assert "p2p=1" in out

#!/bin/bash
# What was run on the GPU boxes this round (gpurun -- 'bash profiles/run_gpu_checks.sh [1|2|8]'):
#   1 GPU : the whole GPU suite, the default bench line, the reference arm on the same mesh
#   N GPUs: the multi-rank parity program (both communicator modes) and the weak-scaling bench
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_full.log 2>&1; echo rc=$? >> gpurun_out/pytest_full.log
  timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  NOSH_TEST_COMM=nccl timeout 600 $TR --master-port 29811 tests/mgpu_worker.py > gpurun_out/mgpu${N}_nccl.log 2>&1
  NOSH_TEST_COMM=host timeout 600 $TR --master-port 29812 tests/mgpu_worker.py > gpurun_out/mgpu${N}_host.log 2>&1
  timeout 900 $TR --master-port 29813 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
fi

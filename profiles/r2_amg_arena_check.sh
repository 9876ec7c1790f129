set -x
timeout 600 python -m pytest tests/test_gpu_amg.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do NOSH_B200_AMG_TIMING=1 timeout 300 python profiles/amg_setup_probe.py 200 2>&1 | tail -30; done > gpurun_out/amg_arena_probe.txt 2>&1
tail -5 gpurun_out/amg_arena_probe.txt

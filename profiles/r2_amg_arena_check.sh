mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_amg.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do NOSH_B200_AMG_TIMING=1 timeout 300 python profiles/amg_setup_probe.py 200 2>&1 | tail -30; done > gpurun_out/amg_arena_probe.txt 2>&1
grep '^{' gpurun_out/amg_arena_probe.txt
NOSH_TEST_SECTIONS=amg,gmres timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --no-parity --no-strong-probe > gpurun_out/bench_arena2.json 2> gpurun_out/bench_arena2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_arena2.json").read().strip().splitlines()[-1])
a=d["newton_solve"]["amg"]; print(a["hierarchy_setup_seconds"], a["hierarchy_setup_phases_s"], d["setup_s"])
PY

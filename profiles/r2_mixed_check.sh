mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_amg.py -m gpu -x -q > gpurun_out/mixed_pytest.log 2>&1
tail -5 gpurun_out/mixed_pytest.log
timeout 900 python bench.py --no-parity > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_mixed.json').read().strip().splitlines()[-1])
for k in ('amg','amg_mixed','none'):
    v=d['newton_solve'].get(k); print(k, v and {q:v[q] for q in ('solve_seconds','newton_steps','minres_iterations_per_step','fnorm')})
print(d['newton_solve'].get('error'))
PY

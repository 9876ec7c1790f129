mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_amg.py tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log )
for C in 1 0 1 0; do
( NOSH_B200_COMPRESS_COLS=$C timeout 600 python bench.py --no-newton --no-parity --no-cpu-baseline --steps 5 > gpurun_out/bench_cc$C.json 2> gpurun_out/bench_cc$C.err )
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cc$C.json').read().strip().splitlines()[-1])
print('compress=$C value %.2f it/s %.1f ms %.2f apply %.4f frac %.3f'%(d['value'],d['minres_iters_per_s'],d['ms_per_step'],d['roofline']['ms_per_launch'],d['roofline']['frac']))
PY
done
( timeout 600 python profiles/unstructured_bench.py --n 200 > gpurun_out/unstructured_cc.json 2> gpurun_out/unstructured_cc.err )
tail -n 3 gpurun_out/pytest.log; cat gpurun_out/unstructured_cc.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['mesh'], '%.3f ms frac %.3f stored/blocks %.3f it %.3f'%(d['apply_ms'],d['frac_of_measured_peak'],d['stored_over_blocks'],d['minres_ms_per_iteration']))"

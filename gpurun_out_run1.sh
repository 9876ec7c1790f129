mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_amg.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log )
for i in 1 2 3; do ( NOSH_B200_AMG_TIMING=1 timeout 300 python profiles/amg_setup_probe.py 200 > gpurun_out/amg_probe_$i.json 2> gpurun_out/amg_probe_$i.err ); done
( timeout 900 python bench.py --no-cpu-baseline --no-parity --steps 3 > gpurun_out/bench_1gpu_b.json 2> gpurun_out/bench_1gpu_b.err; echo rc=$? >> gpurun_out/bench_1gpu_b.err )
tail -n 4 gpurun_out/pytest.log; cat gpurun_out/amg_probe_?.json; grep "amg setup" gpurun_out/amg_probe_1.err | head -12

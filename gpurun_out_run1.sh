mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_amg.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log )
( timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo rc=$? >> gpurun_out/bench_1gpu.err )
tail -n 4 gpurun_out/pytest.log; tail -c 300 gpurun_out/bench_1gpu.err

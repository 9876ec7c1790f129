mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log )
( timeout 900 python bench.py --strong --mesh-n 400 --steps 3 --warmup 3 --no-newton --no-parity --no-cpu-baseline > gpurun_out/bench_1gpu_strong64M.json 2> gpurun_out/bench_1gpu_strong64M.err; echo rc=$? >> gpurun_out/bench_1gpu_strong64M.err )
( timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo rc=$? >> gpurun_out/bench_1gpu.err )
tail -n 5 gpurun_out/pytest.log

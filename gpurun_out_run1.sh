mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo rc=$? >> gpurun_out/pytest.log )
( timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo rc=$? >> gpurun_out/bench_1gpu.err )
( NOSH_B200_AMG_GRAPH=0 timeout 900 python bench.py --workload newton --precond amg --steps 2 --warmup 1 --no-parity > gpurun_out/newton_amg_nograph.json 2> gpurun_out/newton_amg_nograph.err )
( timeout 900 python bench.py --workload newton --precond amg --steps 2 --warmup 1 --no-parity > gpurun_out/newton_amg_graph.json 2> gpurun_out/newton_amg_graph.err )
tail -n 5 gpurun_out/pytest.log

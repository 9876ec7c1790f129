mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( NOSH_TEST_COMM=host timeout 600 $TR --master-port 29811 tests/mgpu_worker.py > gpurun_out/mgpu8_host.log 2>&1; echo rc=$? >> gpurun_out/mgpu8_host.log )
( timeout 900 $TR --master-port 29812 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo rc=$? >> gpurun_out/bench_8gpu.err )
( timeout 900 $TR --master-port 29815 bench.py --gpus 8 --steps 1 --warmup 1 --workload continuation --precond amg --no-parity > gpurun_out/cont8_amg_weak.json 2> gpurun_out/cont8_amg_weak.err; echo rc=$? >> gpurun_out/cont8_amg_weak.err )
grep -h "MGPU\|rc=" gpurun_out/mgpu8_host.log | tail -3; tail -c 150 gpurun_out/bench_8gpu.err gpurun_out/cont8_amg_weak.err

mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( NOSH_TEST_COMM=host timeout 600 $TR --master-port 29712 tests/mgpu_worker.py > gpurun_out/mgpu2_host.log 2>&1; echo rc=$? >> gpurun_out/mgpu2_host.log )
( timeout 900 $TR --master-port 29721 bench.py --gpus 2 --steps 10 --warmup 3 --no-newton > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo rc=$? >> gpurun_out/bench_2gpu.err )
grep -h "MGPU\|rc=" gpurun_out/mgpu2_*.log; tail -c 200 gpurun_out/bench_2gpu.err

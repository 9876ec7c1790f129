mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
( NOSH_TEST_COMM=nccl timeout 600 $TR --master-port 29811 tests/mgpu_worker.py > gpurun_out/mgpu8_nccl.log 2>&1; echo rc=$? >> gpurun_out/mgpu8_nccl.log )
( timeout 900 $TR --master-port 29812 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo rc=$? >> gpurun_out/bench_8gpu.err )
( NOSH_B200_PERSISTENT_MGPU=0 timeout 600 $TR --master-port 29813 bench.py --gpus 8 --steps 5 --warmup 3 --no-newton --no-parity > gpurun_out/bench_8gpu_multilaunch.json 2> gpurun_out/bench_8gpu_multilaunch.err; echo rc=$? >> gpurun_out/bench_8gpu_multilaunch.err )
( timeout 600 $TR --master-port 29814 bench.py --gpus 8 --strong --mesh-n 400 --steps 5 --warmup 3 --no-newton --no-parity > gpurun_out/bench_8gpu_strong64M.json 2> gpurun_out/bench_8gpu_strong64M.err; echo rc=$? >> gpurun_out/bench_8gpu_strong64M.err )
( timeout 900 $TR --master-port 29815 bench.py --gpus 8 --steps 1 --warmup 1 --workload continuation --precond amg --no-parity > gpurun_out/cont8_amg_weak.json 2> gpurun_out/cont8_amg_weak.err; echo rc=$? >> gpurun_out/cont8_amg_weak.err )
( timeout 900 $TR --master-port 29816 bench.py --gpus 8 --steps 1 --warmup 1 --workload continuation --precond amg --strong --mesh-n 200 --no-parity > gpurun_out/cont8_amg_strong8M.json 2> gpurun_out/cont8_amg_strong8M.err; echo rc=$? >> gpurun_out/cont8_amg_strong8M.err )
( timeout 900 $TR --master-port 29817 bench.py --gpus 8 --steps 1 --warmup 1 --workload arclength --precond amg --no-parity > gpurun_out/arc8_amg_weak.json 2> gpurun_out/arc8_amg_weak.err; echo rc=$? >> gpurun_out/arc8_amg_weak.err )
grep -h "MGPU\|rc=" gpurun_out/mgpu8_nccl.log | tail -3; tail -c 150 gpurun_out/*8*.err

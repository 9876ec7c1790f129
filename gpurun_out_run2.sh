mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 900 python -m pytest tests -m gpu -x -q -k "fvm or mesh_set_local or reference_style or multi_gpu" > gpurun_out/pytest_new.log 2>&1; echo rc=$? >> gpurun_out/pytest_new.log )
( timeout 600 $TR --master-port 29731 bench.py --gpus 2 --mesh-n 60 --steps 1 --warmup 1 --workload continuation --precond amg > gpurun_out/cont2.json 2> gpurun_out/cont2.err; echo rc=$? >> gpurun_out/cont2.err )
( timeout 600 $TR --master-port 29732 bench.py --gpus 2 --mesh-n 60 --steps 1 --warmup 1 --workload arclength --precond amg > gpurun_out/arc2.json 2> gpurun_out/arc2.err; echo rc=$? >> gpurun_out/arc2.err )
( timeout 600 $TR --master-port 29733 bench.py --gpus 2 --mesh-n 100 --strong --steps 3 --warmup 3 --no-newton > gpurun_out/strong2.json 2> gpurun_out/strong2.err; echo rc=$? >> gpurun_out/strong2.err )
( timeout 600 $TR --master-port 29734 bench.py --gpus 2 --mesh-n 60 --steps 1 --warmup 1 --workload newton > gpurun_out/newton2.json 2> gpurun_out/newton2.err; echo rc=$? >> gpurun_out/newton2.err )
tail -n 4 gpurun_out/pytest_new.log; tail -c 300 gpurun_out/cont2.err gpurun_out/arc2.err gpurun_out/strong2.err gpurun_out/newton2.err

mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( NOSH_TEST_COMM=host NOSH_TEST_SECTIONS=core timeout 600 $TR --master-port 29712 tests/mgpu_worker.py > gpurun_out/mgpu2_host.log 2>&1; echo rc=$? >> gpurun_out/mgpu2_host.log )
for L in 1 0 1 0; do
( NOSH_B200_MGPU_LEAN=$L timeout 900 $TR --master-port 2972$L bench.py --gpus 2 --steps 10 --warmup 3 --no-newton --no-parity > gpurun_out/bench_2gpu_lean$L.json 2> gpurun_out/bench_2gpu_lean$L.err; echo rc=$? >> gpurun_out/bench_2gpu_lean$L.err )
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_2gpu_lean$L.json').read().strip().splitlines()[-1])
print('lean=$L value %.2f it/s %.1f ms %.2f e2e %.2f apply %.4f'%(d['value'],d['minres_iters_per_s'],d['ms_per_step'],d['e2e']['value'],d['roofline']['ms_per_launch']))
PY
done
grep -h "MGPU\|rc=" gpurun_out/mgpu2_host.log | tail -2

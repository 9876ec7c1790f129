mkdir -p gpurun_out
for i in 1 2; do ( NOSH_B200_AMG_TIMING=1 timeout 300 python profiles/amg_setup_probe.py 200 > gpurun_out/amg_probe_$i.json 2> gpurun_out/amg_probe_$i.err ); done
( NOSH_B200_AMG_TIMING=1 timeout 900 python bench.py --no-cpu-baseline --no-parity --steps 3 > gpurun_out/bench_1gpu_b.json 2> gpurun_out/bench_1gpu_b.err; echo rc=$? >> gpurun_out/bench_1gpu_b.err )
cat gpurun_out/amg_probe_?.json; grep "amg setup" gpurun_out/bench_1gpu_b.err | head -30

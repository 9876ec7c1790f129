mkdir -p gpurun_out
( timeout 900 python profiles/unstructured_bench.py --n 200 > gpurun_out/unstructured.json 2> gpurun_out/unstructured.err; echo rc=$? >> gpurun_out/unstructured.err )
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_bench_n200.csv python bench.py --steps 2 --warmup 1 --no-newton --no-parity --no-cpu-baseline --apply-reps 5 > gpurun_out/ncu_bench.log 2>&1 )
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_sell -s 3 -c 2 -o gpurun_out/r2_apply_natural python profiles/unstructured_bench.py --n 200 --only kuhn/natural --reps 3 > gpurun_out/ncu_nat.log 2>&1 )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_sell -s 3 -c 2 -o gpurun_out/r2_apply_morton python profiles/unstructured_bench.py --n 200 --only kuhn/morton --reps 3 > gpurun_out/ncu_morton.log 2>&1 )
( timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo rc=$? >> gpurun_out/bench_1gpu.err )
( timeout 1200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_n200.json 2> gpurun_out/bench_ref_n200.err; echo rc=$? >> gpurun_out/bench_ref_n200.err )
ls -la gpurun_out

mkdir -p gpurun_out
# fp32-value smoother apply: pairs in flight per thread (apply_variant 1: 4, 2: 5, 0: 6, 3: 8)
for V in 0 1 2 3; do
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/launches_amg_mixed1_v$V.csv -k regex:k_apply_sell \
    python profiles/profile_step.py --n 200 --iters 3 --precond amg --amg-mixed 1 --apply-variant $V > gpurun_out/ncu_amg_mixed1_v$V.log 2>&1
  tail -1 gpurun_out/ncu_amg_mixed1_v$V.log
done

mkdir -p gpurun_out
T0=$(date +%s)
timeout 1000 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
echo "bench wall seconds: $(( $(date +%s) - T0 ))"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity"]["ok"])
print(d.get("strong_scaling_64M"))
print({k:d["newton_solve"][k]["solve_seconds"] for k in ("amg","amg_mixed","none")}, d["newton_solve"]["amg"]["hierarchy_setup_seconds"])
PY

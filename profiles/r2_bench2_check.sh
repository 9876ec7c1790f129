mkdir -p gpurun_out
T0=$(date +%s)
export N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-2} --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus ${N:-2} --steps 5 --warmup 3 > gpurun_out/bench_${N:-2}gpu.json 2> gpurun_out/bench_${N:-2}gpu.err
echo "rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json, os
d=json.loads(open("gpurun_out/bench_%sgpu.json" % os.environ.get("N", "2")).read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity"]["ok"], d["gpu_launches"])
print(d.get("strong_scaling_64M"))
a=d["newton_solve"]; print({k:a[k]["solve_seconds"] for k in ("amg","amg_mixed","none")}, a["amg"]["hierarchy_setup_seconds"])
PY

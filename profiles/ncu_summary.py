"""Turn an .ncu-rep (ncu --set full) into the compact per-launch JSON kept under profiles/:
    python profiles/ncu_summary.py gpurun_out/x.ncu-rep > profiles/r2_ncu_full_x.json
Reads the report here (no GPU needed) with `ncu -i ... --page raw --csv`."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed.sum", "launch__waves_per_multiprocessor"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = {}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            d[k] = [r[i], units[i]]
    res.append(d)
print(json.dumps(res, indent=1))

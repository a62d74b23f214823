#!/usr/bin/env python
"""profiles/r1_traffic.json from an `ncu --set full` report of scripts/ncu_traffic.py: per kernel, the LAST captured launch's
dram bytes (read + write), duration and headline utilisation metrics."""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
names = {"grid_scatter": "grid_scatter", "roi_align": "roi_align", "gemm_ps": "gemm_ffn_up"}
res = {}
def num(d, u, k):
    v = float(d[k].replace(",", "")); unit = u[k].lower()
    for pre, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if unit.startswith(pre): return v * m
    return v
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    for key, name in names.items():
        if key in d["Kernel Name"]:
            res[name] = {"kernel": d["Kernel Name"].split("(")[0], "dram_bytes": num(d, u, "dram__bytes_read.sum") + num(d, u, "dram__bytes_write.sum"),
                         "dram_read": num(d, u, "dram__bytes_read.sum"), "dram_write": num(d, u, "dram__bytes_write.sum"),
                         "duration_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) / (1e3 if u["gpu__time_duration.sum"].startswith("ns") else 1),
                         "tensor_pipe_active_pct": float(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "0") or 0),
                         "dram_throughput_pct": float(d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "0") or 0)}
res["_how"] = "ncu --set full --clock-control none, scripts/ncu_traffic.py (cfg2 shapes, pre-split planes, L2 flushed before each launch); last captured launch per kernel"
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))

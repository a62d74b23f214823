#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into a short text file for profiles/:
per captured launch the headline metrics, plus the top warp-stall instructions of the first launch.

    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/r1_xxx.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration_us"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_throughput_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
]


def ncu(path, page):
    return subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    path = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(path, "raw"))))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full --clock-control none (per-launch, cold-cache, serialised)")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"\n{d.get('Kernel Name', '?')[:110]}  grid={d.get('Grid Size')} block={d.get('Block Size')}")
        for m, name in METRICS:
            if m in d:
                print(f"    {name:<26} {d[m]:>16} {u.get(m, '')}")
    src = list(csv.reader(io.StringIO(ncu(path, "source"))))
    hdr, out, k = None, [], 0
    for r in src:
        if r and r[0] == "Kernel Name":
            k += 1
            if k > 1:
                break
            kname = r[1] if len(r) > 1 else ""
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) >= len(hdr) - 5:
            out.append(r)
    if hdr and out:
        i_s, i_src = hdr.index("# Samples"), hdr.index("Source")
        cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[i_s]) for r in out) or 1
        agg = {}
        for r in out:
            for i in cols:
                agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
        print(f"\n## warp-stall samples, first launch ({kname[:80]}): {tot} samples")
        print("   " + ", ".join(f"{k}={100.0 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:6]))
        for r in sorted(out, key=lambda r: -int(r[i_s]))[:12]:
            top = max(((int(r[i] or 0), hdr[i]) for i in cols), default=(0, ""))
            print(f"   {100.0 * int(r[i_s]) / tot:5.1f}%  {r[i_src][:70]:<70} {top[1]}")


if __name__ == "__main__":
    main()

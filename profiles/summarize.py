#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/<launch list>.csv  profiles/<name>.md
    python profiles/summarize.py kernel   gpurun_out/<capture>.ncu-rep  profiles/<name>.md [kernel regex]

`launches`: per-kernel count / total / mean / share from `ncu --metrics gpu__time_duration.sum` (cold-cache,
serialised launches: shares are meaningful, absolutes are not).  `kernel`: key raw metrics + top stall lines of one
`ncu --set full --import-source on` capture; also rewrites profiles/bp_sweep_traffic.json when the kernel is the
BP sweep (bench.py reports it as roofline.traffic)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum"]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        k = row["Kernel Name"].split("(")[0][-70:]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as out:
        out.write(f"# launch list: {os.path.basename(src)}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` "
                  "(cold-cache, serialised: compare shares).\n\n| kernel | launches | total ms | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.3f} | {v[1] / v[0]:.1f} | {v[1] / tot:.3f} |\n")
    print(open(dst).read())


def kernel(src, dst, pattern=""):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    data = [r for r in data if pattern in r[name_i]] or data
    with open(dst, "w") as out:
        out.write(f"# ncu --set full: {os.path.basename(src)}\n\n")
        for r in data:
            out.write(f"## {r[name_i]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    out.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            out.write("\n| stall reason (warps per issue-active cycle) | value |\n|---|---:|\n")
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and "per_issue_active" in h:
                    out.write(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {r[i]} |\n")
            out.write("\n")
        r = data[0]
        if "k_msgs" in r[name_i] and "bp_run" not in r[name_i]:          # a per-sweep launch only (a bp_run launch = many sweeps)
            def mb(k):
                v, u = float(r[hdr.index(k)]), units[hdr.index(k)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            t = {"dram_bytes_per_launch": mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"),
                 "source": os.path.basename(src), "kernel": r[name_i],
                 "launch_us": float(r[hdr.index("gpu__time_duration.sum")]) *
                 {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[units[hdr.index("gpu__time_duration.sum")]]}
            with open(os.path.join(os.path.dirname(dst), "bp_sweep_traffic.json"), "w") as f:
                json.dump(t, f, indent=1)
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](*sys.argv[2:])

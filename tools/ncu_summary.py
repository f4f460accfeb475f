"""Summarise an `ncu --set full` report (.ncu-rep) per launch: duration, DRAM bytes read / written, achieved DRAM GB/s
against the measured copy peak (MEASURED_PEAKS.json), tensor-pipe activity, registers.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [algorithmic-bytes-python-expr per kernel-name substring ...]

Writes a markdown table to stdout (what profiles/*.md hold)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=float("nan")):
    i = col.get(name)
    if i is None or r[i] == "":
        return default
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(u, 1.0)
    return v * scale


print("| # | kernel | grid | µs | DRAM read MB | DRAM write MB | DRAM GB/s | of measured peak | L2 hit % | tensor pipe % | regs |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for k, r in enumerate(rows[2:]):
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void fcwdm::", "").replace("fcwdm::", "")
    us = get(r, "gpu__time_duration.sum")
    rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / (us * 1e-6) / 1e9
    tens = get(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
               get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
    hit = get(r, "lts__t_sector_hit_rate.pct")
    print(f"| {k} | `{name[:48]}` | {r[col['Grid Size']]} | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | "
          f"{gbs / peaks['hbm_gbs']:.2f} | {hit:.0f} | {tens:.1f} | {get(r, 'launch__registers_per_thread'):.0f} |")

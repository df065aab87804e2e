#!/usr/bin/env python
"""Summarise an `ncu --set full` report (one line per captured launch) for profiles/.

usage: ncu_summary.py <report.ncu-rep> [out.md]
"""
import csv, io, subprocess, sys, re

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}

def g(r, name, default=float("nan")):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default

def unit(name):
    return units[col[name]] if name in col else ""

def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)

def to_mb(v, u):
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)

out = []
out.append(f"# ncu --set full summary of {rep.split('/')[-1]} (per launch; cold-cache, serialised replays)\n")
out.append("| # | kernel | grid | block | regs | dur us | DRAM rd MB | DRAM wr MB | DRAM %peak | L2 (lts) MB | SM thr % | issue act % | fma pipe % | alu pipe % | warps act % | tensor % |")
out.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for n, r in enumerate(data):
    name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]) if "Kernel Name" in col else "?"
    name = name.replace("void ", "").replace("hsp::", "")
    dur = to_us(g(r, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"))
    rd = to_mb(g(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum"))
    wr = to_mb(g(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"))
    lts = to_mb(g(r, "lts__t_bytes.sum"), unit("lts__t_bytes.sum")) if "lts__t_bytes.sum" in col else float("nan")
    out.append("| %d | `%s` | %s | %s | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
        n, name[:60], r[col["Grid Size"]], r[col["Block Size"]], g(r, "launch__registers_per_thread", 0), dur, rd, wr,
        g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), lts,
        g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        g(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        g(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)))
# machine-readable DRAM traffic per launch (bench.py's roofline.traffic reads profiles/kernel_traffic.json)
import json
traffic = []
for r in data:
    name = re.sub(r"\(.*$", "", r[col["Kernel Name"]]).replace("void ", "")
    traffic.append({"kernel": name, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
                    "dram_bytes": (to_mb(g(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum")) +
                                   to_mb(g(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"))) * 1e6,
                    "duration_us": to_us(g(r, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"))})
if len(sys.argv) > 3:
    json.dump(traffic, open(sys.argv[3], "w"), indent=1)
txt = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
print(txt)

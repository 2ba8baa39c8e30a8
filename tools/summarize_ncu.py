#!/usr/bin/env python3
"""Summarises ncu captures into small text files under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

    python tools/summarize_ncu.py launches <launches.csv> <out.md>
    python tools/summarize_ncu.py kernel <report.ncu-rep> <out.md> [--traffic-key workload]
"""
import collections
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct"]


def launches(csv_path, out_path):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        name = r[ki].split("(")[0].split("::")[-1]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    total = sum(v[1] for v in agg.values())
    lines = [f"# ncu launch list summary ({Path(csv_path).name})", "", "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold cache, serialised: compare SHARES, not absolutes).", "",
             "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / total:.3f} |")
    Path(out_path).write_text("\n".join(lines) + "\n")
    print("\n".join(lines))


def kernel(rep_path, out_path, traffic_key=None, top=0):
    raw = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if top:  # keep the `top` longest launches (tail launches of a wavefront step carry few rays)
        ti = hdr.index("gpu__time_duration.sum")
        data = sorted(data, key=lambda r: -float(r[ti]))[:top]
    lines = [f"# ncu --set full summary ({Path(rep_path).name})", "", f"kernel: `{data[0][hdr.index('Kernel Name')]}`", "",
             "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |", "|---|---|" + "---:|" * len(data)]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                wts = [float(r[hdr.index("gpu__time_duration.sum")]) for r in data]  # weighted by launch duration
                stalls.append((sum(float(r[i]) * wt for r, wt in zip(data, wts)) / sum(wts), h))
            except ValueError:
                pass
    lines += ["", "Top stall reasons (warps stalled per issued instruction, averaged over the launches by duration):", ""]
    for v, h in sorted(stalls, reverse=True)[:6]:
        lines.append(f"* {h.split('issue_stalled_')[1].split('_per_issue')[0]}: {v:.2f}")
    Path(out_path).write_text("\n".join(lines) + "\n")
    print("\n".join(lines))
    if traffic_key:
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        per_launch = sum(float(r[ri]) * scale.get(units[ri], 1) + float(r[wi]) * scale.get(units[wi], 1) for r in data) / len(data)
        p = REPO / "profiles" / "traffic.json"
        d = json.loads(p.read_text()) if p.exists() else {}
        d.setdefault(traffic_key, {})["extend_kernel_dram_bytes_per_launch"] = per_launch
        d[traffic_key]["source"] = Path(out_path).name
        d[traffic_key]["launches_averaged"] = len(data)
        # the rooflines the kernel actually sits under (weighted by launch duration): issue slots, the L1 load/store data pipe, L2
        ti = hdr.index("gpu__time_duration.sum")
        weights = [float(r[ti]) for r in data]
        def weighted(key):
            if key not in hdr:
                return None
            i = hdr.index(key)
            return sum(float(r[i].replace(",", "")) * wt for r, wt in zip(data, weights)) / sum(weights)
        d[traffic_key]["issue_slots_active_pct"] = weighted("smsp__issue_active.avg.pct_of_peak_sustained_active")
        d[traffic_key]["active_lanes_per_instruction"] = weighted("smsp__thread_inst_executed_per_inst_executed.ratio")
        d[traffic_key]["l1_lsu_data_pipe_pct"] = weighted("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")
        d[traffic_key]["l2_throughput_pct"] = weighted("lts__throughput.avg.pct_of_peak_sustained_elapsed")
        d[traffic_key]["dram_throughput_pct"] = weighted("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        p.write_text(json.dumps(d, indent=1) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        key = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
        top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
        kernel(sys.argv[2], sys.argv[3], key, top)

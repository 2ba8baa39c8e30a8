#!/usr/bin/env python3
"""Instruction-level view of an `ncu --set full --import-source on` capture: where the issue slots and the stall samples go.

    python tools/ncu_hotspots.py <report.ncu-rep> [--launch N] [--top K]

Prints, for one launch (default: the first), the SASS regions between branch targets ranked by executed warp instructions,
with their average active lanes and stall samples, and the hottest single instructions by stall samples."""
import csv
import io
import subprocess
import sys


def load(rep, launch):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]  # one block per captured launch
    block = rows[starts[2 * launch]:starts[2 * launch + 1]]  # ncu prints every launch twice
    hdr = block[1]
    return block[0][1], hdr, [r for r in block[2:] if len(r) == len(hdr)]


def main():
    rep = sys.argv[1]
    launch = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else 0
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    name, hdr, rows = load(rep, launch)
    col = {h: i for i, h in enumerate(hdr)}
    num = lambda r, k: float(r[col[k]].replace(",", "") or 0)
    total_inst = sum(num(r, "Instructions Executed") for r in rows)
    total_samples = sum(num(r, "# Samples") for r in rows)
    print(f"kernel: {name}\nlaunch {launch}: {total_inst:.3e} warp instructions, {total_samples:.0f} stall samples, {len(rows)} SASS lines")
    # regions: split at instructions whose execution count differs markedly from the previous one (loop / branch boundaries)
    regions, cur = [], None
    for i, r in enumerate(rows):
        n = num(r, "Instructions Executed")
        if cur is None or (max(n, cur["n0"]) > 1.25 * min(n, cur["n0"]) + 16):
            cur = {"first": i, "n0": n, "inst": 0.0, "thread": 0.0, "samples": 0.0, "lines": 0}
            regions.append(cur)
        cur["inst"] += n; cur["thread"] += num(r, "Thread Instructions Executed"); cur["samples"] += num(r, "# Samples"); cur["lines"] += 1
    print(f"\ntop {top} regions by executed warp instructions (share of kernel, lines, avg active lanes, share of stall samples, first SASS line):")
    for g in sorted(regions, key=lambda g: -g["inst"])[:top]:
        lanes = g["thread"] / g["inst"] if g["inst"] else 0
        print(f"  {g['inst'] / total_inst:6.1%}  {g['lines']:4d} lines  {lanes:5.1f} lanes  {g['samples'] / max(total_samples, 1):6.1%} samples   "
              f"{rows[g['first']][col['Address']][-5:]}  {rows[g['first']][col['Source']].strip()[:60]}")
    print(f"\ntop {top} instructions by stall samples:")
    for r in sorted(rows, key=lambda r: -num(r, "# Samples"))[:top]:
        print(f"  {num(r, '# Samples') / max(total_samples, 1):6.1%}  {num(r, 'Avg. Threads Executed'):5.1f} lanes  {r[col['Address']][-5:]}  {r[col['Source']].strip()[:70]}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""BASELINE.json configs[0]: DefaultShading / GGX_R / OrenNayar / Burley evaluate + sample + PDF over 2^22 random tuples.
Device-resident SoA inputs, CUDA events on the library's stream; reports Mtuples/s and GB/s against the 104 B/tuple
algorithmic traffic (60 B in + 44 B out), next to the reference's own headers on the host cores."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bifrost3d_b200 as b
from bifrost3d_b200.workloads import bsdf_tuples

N = 1 << 22
KINDS = {"DefaultShading": 0, "GGX_R": 1, "OrenNayar": 2, "Burley": 3}


def main():
    ctx = b.Bpt(0)
    t = bsdf_tuples(N, seed=1234)
    dev = {k: torch.from_numpy(t[k]).cuda().contiguous() for k in ("wo", "wi", "tint", "rms", "u")}
    out = {"eval_f": torch.empty((N, 3), device="cuda"), "eval_pdf": torch.empty(N, device="cuda"), "sample_f": torch.empty((N, 3), device="cuda"),
           "sample_pdf": torch.empty(N, device="cuda"), "sample_dir": torch.empty((N, 3), device="cuda")}
    ptrs = {k: v.data_ptr() for k, v in {**dev, **out}.items()}
    stream = torch.cuda.ExternalStream(ctx.stream)
    peak = json.loads((REPO / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (REPO / "MEASURED_PEAKS.json").exists() else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    results = {}
    for name, kind in KINDS.items():
        for _ in range(3):
            ctx.bsdf_eval_sample_pdf_device(kind, N, ptrs)
        ctx.synchronize(); torch.cuda.synchronize()
        times = []
        for _ in range(10):
            flush.zero_(); torch.cuda.synchronize()  # 256 MB > L2: inputs come from HBM
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.bsdf_eval_sample_pdf_device(kind, N, ptrs); e1.record(stream)
            ctx.synchronize(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        results[name] = {"ms": ms, "mtuples_per_s": N / ms / 1e3, "gb_per_s": N * 104 / ms / 1e6, "frac_of_measured_hbm": N * 104 / ms / 1e6 / peak}
    try:
        from tests import oracle_lib
        ref = oracle_lib.load()
        n_cpu = 1 << 20
        args = [t[k][:n_cpu] for k in ("wo", "wi", "tint", "rms", "u")]
        for threads, label in ((1, "cpu_1_thread"), (0, "cpu_all_threads")):
            t0 = time.perf_counter(); ref.bsdf_eval_sample_pdf(0, *args, threads=threads); dt = time.perf_counter() - t0
            results[f"DefaultShading_{label}"] = {"mtuples_per_s": n_cpu / dt / 1e6, "threads": threads or ref.max_threads(), "kind": "reference"}
    except Exception as e:  # oracle not built
        results["cpu"] = str(e)
    print(json.dumps({"workload": "configs[0]: 2^22 BSDF tuples", "tuples": N, "algorithmic_bytes_per_tuple": 104, "hbm_peak_gbs": peak, "results": results}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Benchmark of the path-tracing hot path (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cornell|materials|...]

A "step" is ONE progressive sample of every pixel of the frame (one Renderer::render call in the reference,
Renderer.cpp:1250-1265). Metric: Msamples/s = width * height * steps / seconds (whole job, all GPUs); Mrays/s
(closest-hit + shadow rays from device counters) is reported next to it.
  value   : device-timed, scene + BVH resident in HBM, CUDA events on the library's stream.
  e2e     : through the C ABI with host buffers: every step uploads the camera + settings and reads the resolved
            half4 frame back to pinned-size host memory (what DX11OptiXAdaptor displays each frame).
  roofline: closest-hit traversal kernel, algorithmic bytes (48 + 64*ceil(log2(N/4)) + 192 B/ray) over its
            CUDA-event duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline / --impl reference: the CPU oracle (reference shading headers + restated integrator, OpenMP) on host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def bytes_per_ray(triangles):
    return 48 + 64 * max(1, math.ceil(math.log2(max(triangles, 8) / 4))) + 192


def measured_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


def build_scene(name):
    from bifrost3d_b200 import scenes
    if name == "cornell":
        return scenes.cornell_box(), {"max_bounces": 4, "nee_samples": 3, "pdf_scale": 0.5}
    if name == "cornell_small":
        return scenes.cornell_box(sphere_quads=(24, 12), width=256, height=256), {"max_bounces": 4, "nee_samples": 3, "pdf_scale": 0.5}
    if name == "materials":      # configs[2]: 10x10 material grid + HDR environment, 1920x1080
        return scenes.material_grid(), {"max_bounces": 4, "nee_samples": 3, "pdf_scale": 0.5}
    if name == "materials_small":
        return scenes.material_grid(320, 180, grid=4, sphere_quads=(24, 12), env_size=(256, 128), env_samples=512), {"max_bounces": 4, "nee_samples": 3, "pdf_scale": 0.5}
    if name == "terrain":        # configs[3]: 50M triangles, 3840x2160, 8 bounces
        return scenes.instanced_terrain(), {"max_bounces": 8, "nee_samples": 3, "pdf_scale": 0.5}
    if name == "terrain_small":
        return scenes.instanced_terrain(480, 270, (5, 4), 32, 3), {"max_bounces": 8, "nee_samples": 3, "pdf_scale": 0.5}
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML during the timed region (every 10 ms)."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.sm, self.reasons, self.stop_flag, self.max_mhz = device, [], set(), threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[device]) if visible and visible.split(",")[device].isdigit() else device
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is None:
            return
        n = self.nvml
        names = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(self.handle))
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.01)

    def summary(self):
        self.stop_flag.set()
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_oracle_run(scene, settings, steps, warmup, target_seconds_per_step=2.0, threads=0):
    """Times the CPU oracle on a bounded sample of the workload: a band of pixel rows at one sample per pixel per step."""
    from tests import oracle_lib
    sc = oracle_lib.OracleScene(scene)
    w, h = scene["width"], scene["height"]
    cores = sc.ref.max_threads() if threads <= 0 else threads
    # calibrate: a few rows around the image centre
    probe_rows = max(1, min(h, 4))
    r0 = h // 2 - probe_rows // 2
    t0 = time.perf_counter()
    sc.render(scene["camera"], w, h, 1, 1, rows=(r0, r0 + probe_rows), threads=threads, **settings)
    per_row = (time.perf_counter() - t0) / probe_rows
    rows = int(max(1, min(h, target_seconds_per_step / max(per_row, 1e-9))))
    r0 = (h - rows) // 2
    counters_total = np.zeros(2, np.uint64)
    times = []
    for step in range(warmup + steps):
        t0 = time.perf_counter()
        _, counters = sc.render(scene["camera"], w, h, step, 1, rows=(r0, r0 + rows), threads=threads, **settings)
        dt = time.perf_counter() - t0
        if step >= warmup:
            times.append(dt); counters_total += counters
    sc.close()
    seconds = float(sum(times))
    samples = rows * w * steps
    return {"msamples_per_s": samples / seconds / 1e6, "mrays_per_s": float(counters_total.sum()) / seconds / 1e6, "cores": int(cores),
            "sample": f"rows {r0}..{r0 + rows} of {h} ({rows}x{w} pixels), 1 sample per pixel per step, {steps} steps",
            "ms_per_step": seconds / steps * 1e3, "rows": rows}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene, settings = build_scene(args.workload)
    if args.russian_roulette:
        settings["russian_roulette_start"] = args.russian_roulette
    steps, warmup = max(1, min(args.steps, 8)), max(0, min(args.warmup, 1))
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: ask for every core this process may run on instead
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    r = cpu_oracle_run(scene, settings, steps, warmup, target_seconds_per_step=max(2.0, 60.0 / (steps + warmup)), threads=threads)
    line = {"impl": "reference", "metric": "Msamples/s", "value": r["msamples_per_s"], "unit": "Msamples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mrays_per_s": r["mrays_per_s"],
            "config": workload_config(args, scene, settings),
            "cpu_baseline": {"value": r["msamples_per_s"], "unit": "Msamples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["msamples_per_s"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU oracle: the reference's own host-compiled shading/light/RNG headers driven by a restatement of its OptiX integrator over a CPU BVH (the OptiX programs themselves cannot run here); OpenMP, all host cores"}
    print(json.dumps(line), flush=True)
    return 0


BSDF_TUPLES = 1 << 22          # BASELINE.json configs[0]
BSDF_BYTES_PER_TUPLE = 104      # 60 B in (wo, wi, tint, {roughness, metallic, specularity}, u) + 44 B out
BSDF_KINDS = {"DefaultShading": 0, "GGX_R": 1, "OrenNayar": 2, "Burley": 3}


def bsdf_config(args):
    return {"workload": f"bsdf: DefaultShading evaluate_with_PDF + sample over 2^22 random (wo, wi, tint, roughness, metallic, specularity, u) tuples per step and GPU, seed 1234",
            "baseline_config": "configs[0] (BSDF evaluate + sample + PDF over 2^22 tuples)",
            "l2_policy": "a 256 MB buffer is written between the timed steps (larger than L2): every step reads its inputs from HBM",
            "parallelism": f"independent tuple batches x{args.gpus} (no collective)"}


def run_bsdf_reference(args):
    """--impl reference --workload bsdf: the reference's own host-compiled DefaultShading over a bounded sample of the tuples."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from tests import oracle_lib
    from bifrost3d_b200.workloads import bsdf_tuples
    ref = oracle_lib.load()
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = 1 << 20
    t = bsdf_tuples(n, seed=1234)
    tuple_args = [t[k] for k in ("wo", "wi", "tint", "rms", "u")]
    steps, warmup = max(1, min(args.steps, 8)), max(0, min(args.warmup, 1))
    times = []
    for step in range(warmup + steps):
        t0 = time.perf_counter(); ref.bsdf_eval_sample_pdf(0, *tuple_args, threads=threads); dt = time.perf_counter() - t0
        if step >= warmup:
            times.append(dt)
    value = n * steps / sum(times) / 1e6
    sample = f"the first 2^20 of the 2^22 tuples per step, {steps} steps"
    print(json.dumps({"impl": "reference", "metric": "Mtuples/s", "value": value, "unit": "Mtuples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                      "ms_per_step": sum(times) / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": bsdf_config(args),
                      "cpu_baseline": {"value": value, "unit": "Mtuples/s", "cores": threads, "kind": "reference", "sample": sample},
                      "e2e": {"value": value, "unit": "Mtuples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "note": "the reference's own headers (DefaultShading.h, GGX.h, OrenNayar.h) compiled for the host, OpenMP over the tuples"}), flush=True)
    return 0


def run_bsdf(args):
    """--workload bsdf = configs[0]: a step is one pass of DefaultShading evaluate + sample + PDF over 2^22 tuples resident in HBM."""
    import torch
    import bifrost3d_b200 as b
    from bifrost3d_b200.workloads import bsdf_tuples
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA extension is the product and there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = b.Bpt(local_rank)
    N, K, Wm = BSDF_TUPLES, args.steps, args.warmup
    t = bsdf_tuples(N, seed=1234 + rank)
    names_in, names_out = ("wo", "wi", "tint", "rms", "u"), {"eval_f": 3, "eval_pdf": 1, "sample_f": 3, "sample_pdf": 1, "sample_dir": 3}
    dev = {k: torch.from_numpy(t[k]).cuda().contiguous() for k in names_in}
    out = {k: torch.empty((N, c) if c > 1 else (N,), device="cuda") for k, c in names_out.items()}
    ptrs = {k: v.data_ptr() for k, v in {**dev, **out}.items()}
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        ctx.synchronize(); torch.cuda.synchronize()
        if distributed:
            dist.barrier(); torch.cuda.synchronize()

    def timed(kind, steps):
        total = 0.0
        for _ in range(steps):
            flush.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.bsdf_eval_sample_pdf_device(kind, N, ptrs); e1.record(stream)
            ctx.synchronize()
            total += e0.elapsed_time(e1)
        return total

    for _ in range(Wm):
        ctx.bsdf_eval_sample_pdf_device(0, N, ptrs)
    barrier()
    ctx.counters(reset=True)
    sampler = ClockSampler(local_rank); sampler.start()
    device_ms = timed(0, K)
    barrier()
    clocks = sampler.summary()
    launches = ctx.counters()["kernel_launches"]
    tm = torch.tensor([device_ms], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    device_ms = float(tm.item())
    value = N * K * world / (device_ms * 1e-3) / 1e6
    others = {name: N / (timed(kind, 5) / 5) / 1e3 for name, kind in BSDF_KINDS.items() if kind != 0}  # Mtuples/s of the single-lobe BSDFs

    # end to end: host arrays in, host arrays out through the C ABI (bpt_bsdf_eval_sample_pdf with on_device = 0)
    host_in = [t[k] for k in names_in]
    ctx.bsdf_eval_sample_pdf(0, *host_in)
    barrier()
    e2e_steps = min(K, 4)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        result = ctx.bsdf_eval_sample_pdf(0, *host_in)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = N * e2e_steps * world / float(te.item()) / 1e6

    if rank == 0:
        peak, peak_kind = measured_peak()
        achieved = N * BSDF_BYTES_PER_TUPLE / (device_ms / K * 1e-3) / 1e9 * 1.0
        line = {"metric": "Mtuples/s", "value": value, "unit": "Mtuples/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": device_ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bsdf_config(args),
                "e2e": {"value": e2e_value, "unit": "Mtuples/s", "h2d_bytes_per_step": int(N * 60), "d2h_bytes_per_step": int(N * 44), "steps": e2e_steps},
                "gpu_launches": int(launches), "gpu_launches_per_step": launches / max(K, 1),
                "roofline": {"bound": "hbm", "kernel": "bsdf_batch_kernel<DefaultShading>", "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": None, "algorithmic_bytes_per_launch": N * BSDF_BYTES_PER_TUPLE,
                             "algorithmic_bytes_per_tuple": BSDF_BYTES_PER_TUPLE, "avg_launch_ms": device_ms / K,
                             "note": "ALU bound, not HBM bound (IEEE division / square root and fp64 sin / cos / pow for parity with the host build)"},
                "other_bsdfs_mtuples_per_s": others, "clocks": clocks,
                "finite_fraction_of_outputs": float(np.isfinite(result["eval_f"]).mean())}
        if not args.no_cpu_baseline and world == 1 and cpu_baseline_available():
            from tests import oracle_lib
            ref = oracle_lib.load()
            n_cpu = 1 << 20
            cpu_args = [t[k][:n_cpu] for k in names_in]
            ref.bsdf_eval_sample_pdf(0, *[a[:4096] for a in cpu_args])
            t0 = time.perf_counter(); ref.bsdf_eval_sample_pdf(0, *cpu_args); dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_cpu / dt / 1e6, "unit": "Mtuples/s", "cores": int(ref.max_threads()), "kind": "reference",
                                    "sample": "the first 2^20 of the 2^22 tuples, one pass"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if distributed:
        dist.destroy_process_group()
    return 0


def workload_config(args, scene, settings):
    from bifrost3d_b200 import scenes
    return {"workload": f"{scene['name']}: {scene['width']}x{scene['height']}, {scenes.triangle_count(scene)} triangles, {len(scene['lights'])} light(s), "
                        f"max_bounce_count {settings['max_bounces']}, next_event_sample_count {settings['nee_samples']}, 1 sample per pixel per step"
                        + (f", Russian roulette from bounce {settings['russian_roulette_start']} (extension, not in the reference)" if settings.get("russian_roulette_start") else "")
                        + (", environment next event estimation by CDF inversion on the device" if scene.get("environment", {}).get("nee") == "cdf" else ""),
            "baseline_config": {"cornell": "configs[1] (SmallPT-style Cornell box)", "materials": "configs[2] (material grid + HDR environment, 1080p)", "terrain": "configs[3] (50M-triangle instanced scene, 4K, 8 bounces)"}.get(args.workload, args.workload),
            "l2_policy": "per-step working set (path state + frame buffers) exceeds L2; no explicit flush",
            "parallelism": f"sample-index sharding x{args.gpus}, replicated BVH"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="materials", help="materials = BASELINE.json configs[2] (1080p, 4 bounces: the configuration the metric is quoted on); cornell = configs[1]; terrain = configs[3]; bsdf = configs[0] (2^22 BSDF tuples)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spp-total", type=int, default=0, metavar="T",
                    help="configs[4]: render T samples per pixel in total, T / N on each of the N GPUs (strong scaling; overrides --steps)")
    ap.add_argument("--environment-nee", default=None, choices=["presampled", "cdf"],
                    help="how next event estimation samples the environment map: presampled lights (the reference renderer's way, default) or CDF inversion on the device (SURVEY.md 8(d) C3: report both)")
    ap.add_argument("--sort-hits", type=int, default=None, metavar="K",
                    help="sort the surface hits by (shading class, hit cell) before shading from wavefront iteration K on (bpt_set_hit_sorting); -1 = never; default: the library's")
    ap.add_argument("--russian-roulette", type=int, default=0, metavar="N",
                    help="opt-in Russian roulette from the N-th surface interaction on (configs[3] names it; 0 = off = the reference's behaviour)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.workload == "bsdf":
        return run_bsdf_reference(args) if args.impl == "reference" else run_bsdf(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import bifrost3d_b200 as b
    from bifrost3d_b200 import scenes, capi

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    distributed = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA extension is the product and there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene, settings = build_scene(args.workload)
    if args.russian_roulette:
        settings["russian_roulette_start"] = args.russian_roulette
    if args.environment_nee and scene.get("environment", {}).get("texels") is not None:
        scene["environment"]["nee"] = args.environment_nee
    W, H = scene["width"], scene["height"]
    ctx = b.Bpt(local_rank)
    if args.sort_hits is not None:
        ctx.set_hit_sorting(args.sort_hits)
    scenes.upload(ctx, scene)
    ctx.build_accel()  # second build: the reported build time excludes one-off module loading and allocator warm-up
    info = ctx.accel_info()
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    cam = capi.make_camera(*scene["camera"])
    if args.spp_total:
        if args.spp_total % world:
            raise SystemExit(f"--spp-total {args.spp_total} is not a multiple of the {world} GPUs")
        args.steps = args.spp_total // world  # sample ranges [g T / N, (g + 1) T / N), SURVEY.md 8(d) C5
    K, Wm = args.steps, args.warmup

    def barrier():
        ctx.synchronize(); torch.cuda.synchronize()
        if distributed:
            dist.barrier()
            torch.cuda.synchronize()

    # Each rank renders its own contiguous range of sample indices (weak scaling: K samples per GPU).
    from bifrost3d_b200.sharding import sample_range
    first = sample_range(rank, K, warmup=Wm)[0] - Wm
    # ---- device-timed region -----------------------------------------------------------------------
    if distributed:
        # The library's own communicator (NCCL bound inside libbpt.so): rank 0 creates the unique id, torch.distributed only carries it.
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], world, rank)
    ctx.render(cam, W, H, first, Wm, reset=True, **settings)  # warm-up
    if distributed:
        ctx.reduce_accumulation(0)  # warm-up of the collective too: NCCL sets its channels up on first use
    barrier()
    ctx.counters(reset=True)
    sampler = ClockSampler(local_rank); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.render(cam, W, H, first + Wm, K, reset=True, **settings)
    if distributed:
        # combine the per-GPU radiance sums: one ncclReduce over NVLink of the double4 accumulation buffer, enqueued on the
        # render stream behind the last sample (no host synchronisation inside the timed region)
        ctx.reduce_accumulation(0)
    e1.record(stream)
    barrier()
    if distributed:
        ctx.comm_check()  # an asynchronous NCCL error (a lost peer, a link error) fails the run instead of a wrong image
    device_ms = e0.elapsed_time(e1)
    clocks = sampler.summary()
    counters = ctx.counters()
    t = torch.tensor([device_ms, info["build_ms"]], dtype=torch.float64, device="cuda")
    rays = torch.tensor([float(counters["extend_rays"]), float(counters["shadow_rays"])], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    device_ms = float(t[0].item()); slowest_build_ms = float(t[1].item())
    total_samples = W * H * K * world
    value = total_samples / (device_ms * 1e-3) / 1e6
    mrays = float(rays.sum().item()) / (device_ms * 1e-3) / 1e6

    # ---- roofline for the dominant kernel (closest-hit traversal) on this rank ---------------------
    # The timed region above runs each sample as one CUDA graph (no host work inside), which leaves no place for CUDA events
    # between the stages. The per-kernel durations therefore come from an instrumented pass right behind it: the same kernels on
    # the same sample indices as plain stream launches with a CUDA event between the stages (bpt_set_profiling).
    barrier()
    ctx.set_profiling(True)
    Kp = min(K, 32)
    ctx.render(cam, W, H, first + Wm, 1, reset=True, **settings)
    ctx.counters(reset=True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    ctx.render(cam, W, H, first + Wm, Kp, reset=True, **settings)
    p1.record(stream)
    barrier()
    prof = ctx.counters()
    prof_ms = p0.elapsed_time(p1)
    ctx.set_profiling(False)
    peak, peak_kind = measured_peak()
    capture = ncu_capture(args.workload)
    bpr = bytes_per_ray(info["triangles"])
    extend_s = prof["extend_ms"] * 1e-3
    shadow_s = prof["shadow_ms"] * 1e-3
    launches_per_step = counters["kernel_launches"] / max(K, 1)
    extend_launches = max(int(prof["iterations"]), 1)
    achieved = prof["extend_rays"] * bpr / max(extend_s, 1e-12) / 1e9
    roofline = {"bound": "hbm", "kernel": "extend_kernel (closest-hit BVH traversal)", "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(args.workload), "algorithmic_bytes_per_launch": prof["extend_rays"] * bpr / extend_launches,
                "algorithmic_bytes_per_ray": bpr,
                "rays_per_launch": prof["extend_rays"] / extend_launches, "avg_launch_ms": prof["extend_ms"] / extend_launches,
                "measured_in": f"instrumented pass of {Kp} steps behind the timed region (stream launches + CUDA events between the stages), {prof_ms / Kp:.3f} ms/step against {device_ms / K:.3f} ms/step for the graph launches of the timed region",
                "share_of_step": {"extend": prof["extend_ms"] / prof_ms, "shade": prof["shade_ms"] / prof_ms, "shadow": prof["shadow_ms"] / prof_ms},
                "ms_per_step": {"extend": prof["extend_ms"] / Kp, "shade": prof["shade_ms"] / Kp, "shadow": prof["shadow_ms"] / Kp, "instrumented_total": prof_ms / Kp},
                # The HBM figure above is the model SURVEY.md 8(d) prescribes; the scene's hierarchy is served by L1 / L2, and what the
                # kernel is limited by are the issue slots (at the SIMT width below) and the L1 load/store data pipe: from the ncu capture.
                "measured_limits_from_ncu": {k: capture.get(k) for k in ("issue_slots_active_pct", "active_lanes_per_instruction", "l1_lsu_data_pipe_pct",
                                                                         "l2_throughput_pct", "dram_throughput_pct", "launches_averaged", "source")},
                "shadow_kernel_achieved": prof["shadow_rays"] * bpr / max(shadow_s, 1e-12) / 1e9,
                "grays_per_s_extend": prof["extend_rays"] / max(extend_s, 1e-12) / 1e9}

    # ---- end-to-end through the C ABI with host buffers ---------------------------------------------
    barrier()
    # A ring of pinned host frames (bpt_resolve_half4_async has BPT_FRAME_SLOTS = 4 slots): the read-back of frame k overlaps
    # the rendering of the following frames; every frame is complete in host memory before the timed region ends.
    SLOTS = 4
    frames_t = [torch.empty((H, W, 4), dtype=torch.int16).pin_memory() for _ in range(SLOTS)]
    frames = [f.numpy().view(np.uint16) for f in frames_t]
    for k in range(SLOTS):
        ctx.render(cam, W, H, first + k, 1, reset=(k == 0), **settings); ctx.resolve_half4_async(frames[k % SLOTS], k % SLOTS)
    for slot in range(SLOTS):
        ctx.wait_frame(slot)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = min(K, 32)
    for k in range(e2e_steps):
        cam_k = capi.make_camera(*scene["camera"])  # host-side camera state is rebuilt and uploaded every step
        ctx.render(cam_k, W, H, first + k, 1, reset=(k == 0), **settings)
        ctx.wait_frame(k % SLOTS)                              # the frame this slot delivered SLOTS steps ago has been consumed
        ctx.resolve_half4_async(frames[k % SLOTS], k % SLOTS)  # device -> host read of the frame
    for slot in range(SLOTS):
        ctx.wait_frame(slot)
    barrier()
    e2e_s = time.perf_counter() - t0
    frame = frames[(e2e_steps - 1) % SLOTS]
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = W * H * e2e_steps * world / float(te.item()) / 1e6
    frame_finite = bool(np.isfinite(frame.view(np.float16).astype(np.float32)).all())
    nonfinite = torch.tensor([float(ctx.counters()["nonfinite_samples"]) + float(counters["nonfinite_samples"]), 0.0 if frame_finite else 1.0], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(nonfinite, op=dist.ReduceOp.SUM)

    if rank == 0:
        line = {"metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": device_ms / K,
                "higher_is_better": True, "scaling": "strong" if args.spp_total else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "iterations_per_step": counters["iterations"] / K,
                "mrays_per_s": mrays, "extend_rays_per_step": counters["extend_rays"] / K, "shadow_rays_per_step": counters["shadow_rays"] / K,
                "config": workload_config(args, scene, settings),
                "bvh": {"triangles": info["triangles"], "nodes": info["nodes"], "build_ms": info["build_ms"], "mtris_per_s": info["triangles"] / max(info["build_ms"], 1e-6) / 1e3,
                        "node_width": info["node_width"], "traversed_nodes": info["traversed_nodes"], "levels": info["levels"],
                        "slowest_rank_build_ms": slowest_build_ms, "note": "every rank builds the same hierarchy over the replicated scene, outside the timed region"},
                "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(capi.C.sizeof(capi.Camera) + capi.C.sizeof(capi.Settings)),
                        "d2h_bytes_per_step": int(frame.nbytes), "steps": e2e_steps},
                "gpu_launches": int(counters["kernel_launches"]), "gpu_launches_per_step": launches_per_step,
                "nonfinite_samples": int(nonfinite[0].item()), "frames_with_nonfinite_pixels": int(nonfinite[1].item()),
                "roofline": roofline, "clocks": clocks}
        if not args.no_cpu_baseline and world == 1 and cpu_baseline_available():
            r = cpu_oracle_run(scene, settings, steps=3, warmup=1, target_seconds_per_step=4.0)
            line["cpu_baseline"] = {"value": r["msamples_per_s"], "unit": "Msamples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                    "mrays_per_s": r["mrays_per_s"]}
        print(json.dumps(line), flush=True)
    ctx.close()
    if distributed:
        dist.destroy_process_group()
    return 0


def ncu_capture(workload):
    """What the committed `ncu --set full` capture of this workload says about extend_kernel (profiles/traffic.json, written by
    tools/summarize_ncu.py from every launch of one sample): DRAM bytes per launch and the utilisation of the units the kernel
    actually sits under. {} when there is no capture."""
    p = REPO / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(workload, {})
        except Exception:
            return {}
    return {}


def ncu_traffic(workload):
    return ncu_capture(workload).get("extend_kernel_dram_bytes_per_launch")


def cpu_baseline_available():
    from tests import oracle_lib
    return oracle_lib.available()


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


def accumulation_tensor(ctx, width, height, device):
    """torch view (no copy) of the context-owned double4 accumulation buffer, for the NCCL reduce."""
    import torch
    holder = _CudaArray(ctx.accumulation_device_ptr(), (height, width, 4), "<f8")
    return torch.as_tensor(holder, device=torch.device("cuda", device))


if __name__ == "__main__":
    sys.exit(main())

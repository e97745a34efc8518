#!/usr/bin/env python3
"""Benchmark of the path-tracing hot path (BASELINE.json metric: Mrays/s, paths/s at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm (under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation on the host cores

A "step" is one pass of the hot path over one batch: `--spp-per-step` samples per pixel of the workload scene on every
GPU (weak scaling: per-GPU work is fixed, rank g traces the global frames f = g (mod N)), followed for N > 1
by the NCCL sum-reduce of the accumulators to rank 0.  Default workload = BASELINE config 4 / north_star target:
data/book2_final_scene_10000_samples.json at its authored 600x600, max_depth 50.

Rank 0 prints ONE JSON line.  `value` = rays traced by all ranks / device time (CUDA events on the launching stream, max
over ranks) with the scene resident in HBM; `e2e` = the same metric through the public C-ABI path with host buffers: scene
upload (H2D) + render + mean-image read-back (D2H) inside the timed region; `roofline` = the dominant kernel (k_traverse)
against the measured HBM peak; `cpu_baseline` = the reference CPU renderer timed on this box's host cores (N = 1 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

DEFAULT_SCENE = "book2_final_scene_10000_samples"
# algorithmic HBM bytes per ray of k_traverse: ray origin+time 16 + direction 16 read, closest-surface
# record 16 written
EXTEND_BYTES_PER_RAY = 16 + 16 + 16
# DRAM bytes per ray of k_traverse from the committed `ncu --set full` capture (profiles/r01d_ncu_summary.txt: 59.5 + 4.5 MB
# for the ~1.9 M rays of bounce 2, 44.5 + 2.4 MB for the ~1.4 M rays of bounce 3): used for roofline.traffic = per-ray
# traffic x rays per launch
EXTEND_DRAM_BYTES_PER_RAY_NCU = 34.0
# arithmetic credited per unit of algorithmic work: AABB slab test 30, sphere test 30 (to the
# discriminant; a lower bound), quad test 57, instance visit 42
FLOP_AABB, FLOP_SPHERE, FLOP_QUAD, FLOP_INSTANCE = 30, 30, 57, 42


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scene", default=DEFAULT_SCENE, help="scene name under data/ (or 'synthetic:<n_spheres>')")
    ap.add_argument("--spp-per-step", type=int, default=256, help="samples per pixel per step = one wavefront batch (92 M paths at 600x600)")
    ap.add_argument("--spp-total", type=int, default=10000, help="samples-per-pixel setting (fixes the stratification grid)")
    ap.add_argument("--max-depth", type=int, default=50)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--fast-math", action="store_true", help="FMA-contracted intersection arithmetic (not bit-exact)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-spp", type=int, default=0, help="spp of the bounded CPU-baseline sample (0 = auto)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_scene(rt, args):
    if args.scene.startswith("synthetic:"):
        n = int(args.scene.split(":")[1])
        # large instances skip the host SAH build (seconds per million spheres): the BVH is built on the device (run_ours adds
        # RT2_FLAG_GPU_LBVH for them)
        return (rt.Scene.synthetic_spheres(n, width=args.width or 3840, height=args.height or 2160, host_bvh=n < 2_000_000),
                f"synthetic {n}-sphere BVH stress scene")
    path = os.path.join(ROOT, "data", args.scene + ".json")
    return rt.Scene.load(path, data_dir=os.path.join(ROOT, "data")), f"data/{args.scene}.json"


# ---------------------------------------------------------------------------------------------------------------------
def reference_sample(scene_name, dims, spp, spp_total, max_depth, threads):
    """Times the reference's CPU implementation of the path (oracle/_ref = its unmodified sources; else the port) on a
    bounded sample.  Returns (Mrays/s, paths/s, kind, seconds, rays)."""
    from oracle import ref_oracle
    path = os.path.join(ROOT, "data", scene_name + ".json")
    if ref_oracle.available():
        try:
            sc = ref_oracle.RefScene(path, spp_total, dims=dims)
            kind = "reference"
        except RuntimeError:
            sc = None  # legacy-format files: HEAD's own loader throws
    else:
        sc = None
    if sc is None:
        from oracle import rt_oracle
        sc = rt_oracle.PortScene(path, spp_total, dims=dims)
        kind = "port"
    _, _, rays, sec = sc.render(0, spp, max_depth, threads, False)
    w, h = sc.width, sc.height
    return rays / sec * 1e-6, w * h * spp / sec, kind, sec, rays, (w, h)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    dims = (args.width, args.height) if args.width and args.height else None
    spp = max(1, args.spp_per_step // 32)  # bounded sample per step: the CPU is ~600x slower than one B200
    if args.scene.startswith("synthetic:"):
        print(json.dumps({"impl": "reference", "unavailable": "the synthetic scene has no JSON file the reference could load"}))
        return 0
    total_rays, total_sec, kind, wh = 0, 0.0, "port", (0, 0)
    for i in range(args.warmup + args.steps):
        mr, pps, kind, sec, rays, wh = reference_sample(args.scene, dims, spp, args.spp_total, args.max_depth, threads)
        if i >= args.warmup:
            total_rays += rays
            total_sec += sec
    value = total_rays / total_sec * 1e-6
    paths = wh[0] * wh[1] * spp * args.steps
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic camera samples on the repo's scene file",
        "config": {"workload": f"data/{args.scene}.json {wh[0]}x{wh[1]} max_depth {args.max_depth}, {spp} spp per step on the host CPU",
                   "spp_per_step": spp, "paths_per_s": paths / total_sec},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind,
                         "sample": f"{args.steps} x {spp} spp at {wh[0]}x{wh[1]}, {total_sec:.1f} s of CPU work"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import raytrace2_b200 as rt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available() or rt.load_library().rt2_device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device — the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    scene, scene_label = load_scene(rt, args)
    dims = (args.width, args.height) if args.width and args.height else None
    flags = rt.RT2_FLAG_FAST_MATH if args.fast_math else 0
    if args.scene.startswith("synthetic:") and int(args.scene.split(":")[1]) >= 2_000_000:
        flags |= rt.RT2_FLAG_GPU_LBVH
    tracer = rt.RayTracer(scene, num_samples=args.spp_total, max_depth=args.max_depth, device=local_rank, seed=20261018,
                          flags=flags, frames_per_batch=args.spp_per_step, frame_offset=rank, frame_stride=world, dims=dims)
    W, H = tracer.Dims()
    S = args.spp_per_step
    ext = torch.cuda.ExternalStream(tracer.stream(), device=dev)
    acc_ptr, acc_n = tracer.accum_device_ptr()
    from raytrace2_b200.distributed import _CudaBuffer
    accum = torch.as_tensor(_CudaBuffer(acc_ptr, acc_n), device=dev)
    scratch = torch.empty_like(accum)

    def step():
        tracer.Update(S)
        if world > 1:
            scratch.copy_(accum, non_blocking=True)
            dist.reduce(scratch, dst=0, op=dist.ReduceOp.SUM)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(ext):
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        st0 = tracer.stats()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ext)
        for _ in range(args.steps):
            step()
        ev1.record(ext)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = ev0.elapsed_time(ev1)
        st1 = tracer.stats()
    rays = st1["rays"] - st0["rays"]
    paths = st1["paths"] - st0["paths"]
    launches = st1["launches"] - st0["launches"]
    tot = torch.tensor([float(rays), float(paths), float(launches)], dtype=torch.float64, device=dev)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    rays_all, paths_all, launches_all = (float(x) for x in tot.tolist())
    ms_all = float(tmax.item())

    # ---- end-to-end through the public C-ABI path with host buffers: upload scene, render, read the mean image back ----
    def e2e_step():
        tracer.upload_scene()           # H2D: flattened scene from host memory
        tracer.Reset()
        tracer.Update(S)
        return tracer.NonConvertedPixels()  # D2H: W*H*3 floats (synchronises)

    for _ in range(2):
        e2e_step()
    barrier()
    e0 = tracer.stats()["rays"]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    e2e_rays = tracer.stats()["rays"]  # stats are reset by Reset(): rays of the last step only
    e2e_t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
    e2e_r = torch.tensor([float(e2e_rays) * args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_r, op=dist.ReduceOp.SUM)
    d = scene.desc
    scene_bytes = (d.n_spheres * 32 + d.n_quads * 80 + d.n_xforms * 96 + d.n_instances * 16 + d.n_media * 32 + d.n_materials * 32 +
                   d.n_textures * 48 + d.n_perlin * 7168 + d.n_prim_refs * 4 + d.n_node_pairs * 64)
    e2e = {"value": float(e2e_r.item()) / float(e2e_t.item()) * 1e-6, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes),
           "d2h_bytes_per_step": int(W * H * 3 * 4),
           "host_memory": "pinned staging arena inside the library (cudaMallocHost), host buffers on both sides of the C ABI"}
    del e0

    # ---- roofline of the dominant kernel: per-kernel CUDA-event split over extra (untimed) profiled steps ----
    tracer.Reset()
    tracer.set_profiling(True)
    for _ in range(min(args.steps, 4)):
        tracer.Update(S)
    ps = tracer.stats()
    tracer.set_profiling(False)
    peaks, peak_src = measured_peaks()
    ext_ms = ps["gpu_ms_extend"]
    prof_total = ps["gpu_ms_extend"] + ps["gpu_ms_shade"] + ps["gpu_ms_other"] + ps["gpu_ms_finish"]
    achieved = ps["rays"] * EXTEND_BYTES_PER_RAY / (ext_ms * 1e-3) * 1e-9 if ext_ms > 0 else 0.0
    prof_steps = min(args.steps, 4)
    n_trav_launches = max(1, prof_steps * args.max_depth)  # one k_traverse launch per bounce and batch
    rays_per_launch = ps["rays"] / n_trav_launches
    kernel_name = "k_traverse_flat" if ps["box_pair_tests"] == 0 else "k_traverse"
    roofline = {"kernel": kernel_name, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"],
                "traffic": (EXTEND_DRAM_BYTES_PER_RAY_NCU * rays_per_launch) if (EXTEND_DRAM_BYTES_PER_RAY_NCU and kernel_name == "k_traverse") else None,
                "traffic_note": "ncu dram__bytes_read+write per ray (profiles/) x mean rays per launch; algorithmic = 48 B/ray x the same",
                "peak_source": peak_src, "algorithmic_bytes_per_ray": EXTEND_BYTES_PER_RAY,
                "algorithmic_bytes_per_launch": EXTEND_BYTES_PER_RAY * rays_per_launch,
                "avg_launch_ms": ext_ms / n_trav_launches,
                "share_of_step": ext_ms / prof_total if prof_total > 0 else None,
                "note": "the extend kernel is instruction-issue bound under divergence (scene lives in L1/L2), so its HBM fraction is small "
                        "by construction; roofline_fp32 below is the roofline that binds it (DESIGN.md §4, profiles/)"}
    # the roofline that actually binds the extend stage: credited arithmetic of the algorithmic work done by active lanes
    # (device counters of the profiling build) against the FP32 FMA peak measured on this GPU by our micro-benchmark
    fp32_peak = C.c_double(0.0)
    rt.load_library().rt2_measure_fp32_peak(local_rank, C.byref(fp32_peak))
    flops = (2 * ps["box_pair_tests"] * FLOP_AABB + ps["sphere_tests"] * FLOP_SPHERE + ps["quad_tests"] * FLOP_QUAD +
             ps["instance_visits"] * FLOP_INSTANCE)
    ach_tf = flops / (ext_ms * 1e-3) * 1e-12 if ext_ms > 0 else 0.0
    roofline_fp32 = {"kernel": kernel_name, "bound": "fp32", "achieved": ach_tf, "peak": fp32_peak.value, "unit": "TFLOP/s",
                     "frac": ach_tf / fp32_peak.value if fp32_peak.value > 0 else None,
                     "peak_source": "measured here (rt2_measure_fp32_peak: FMA micro-benchmark, 2 flop per FMA)",
                     "flop_per_ray": flops / max(ps["rays"], 1),
                     "per_ray": {"aabb_tests": 2 * ps["box_pair_tests"] / max(ps["rays"], 1), "sphere_tests": ps["sphere_tests"] / max(ps["rays"], 1),
                                 "quad_tests": ps["quad_tests"] / max(ps["rays"], 1), "instance_visits": ps["instance_visits"] / max(ps["rays"], 1)},
                     "note": "credited flops per SURVEY Appendix C (30 / AABB test, 30 / sphere, 57 / quad, 42 / instance visit) of the work "
                             "done by ACTIVE lanes; unavailable (0) for the flat extend kernel, which has no counters"}
    kernel_split = {"extend_ms": ps["gpu_ms_extend"], "finish_shade_ms": ps["gpu_ms_finish"], "deferred_shade_ms": ps["gpu_ms_shade"],
                    "sort_ms": ps["gpu_ms_sort"], "other_ms": ps["gpu_ms_other"], "steps": prof_steps}

    line = {
        "metric": "Mrays/s", "value": rays_all / (ms_all * 1e-3) * 1e-6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic camera samples (Philox) on the repo's scene file; random Perlin tables",
        "config": {"workload": f"{scene_label} {W}x{H}, max_depth {args.max_depth}, {S} spp per step per GPU",
                   "spp_per_step_per_gpu": S, "paths_per_s": paths_all / (ms_all * 1e-3), "rays_per_path": rays_all / max(paths_all, 1),
                   "intersection_math": "fast (FMA)" if args.fast_math else "exact (bit-identical with the reference)",
                   "l2": "wavefront state per step (%.0f MB) exceeds L2; no explicit flush" % (W * H * S * 184 / 1e6),
                   "parallelism": f"sample-partition x{world}" if world > 1 else "single GPU"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_all), "roofline": roofline, "roofline_fp32": roofline_fp32,
        "kernel_split_profiled": kernel_split,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.scene.startswith("synthetic:"):
        threads = os.cpu_count() or 1
        spp_cpu = args.cpu_sample_spp or max(2, min(128, int(round(15 * threads / 8))))  # ~10 s of wall clock on all host threads
        mr, pps, kind, sec, crays, wh = reference_sample(args.scene, dims, spp_cpu, args.spp_total, args.max_depth, threads)
        line["cpu_baseline"] = {"value": mr, "unit": "Mrays/s", "cores": threads, "kind": kind, "paths_per_s": pps,
                                "sample": f"{spp_cpu} spp of the same workload at {wh[0]}x{wh[1]} ({sec:.1f} s)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""Benchmark of the path-tracing hot path (BASELINE.json metric: Mrays/s and paths/s at 1/2/4/8 B200; wall time to a 10k-spp
converged frame).

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm (under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU implementation on the host cores

Rank 0 prints ONE JSON line.

  value / ms_per_step   WEAK scaling, the headline rate: a "step" is one wavefront batch of `--spp-per-step` samples per pixel
                        of the workload scene on every GPU (rank g traces the global frames f = g (mod N)) followed
                        for N > 1 by the NCCL sum-reduce of the accumulators to rank 0.  CUDA events on the launching stream,
                        max over ranks, scene resident in HBM.  Default workload = BASELINE config 4 / north_star target:
                        data/book2_final_scene_10000_samples.json at its authored 600x600, max_depth 50.
  e2e                   the same metric through the public C-ABI path with HOST buffers: scene upload (H2D) + render +
                        mean-image read-back (D2H) inside the timed region (host wall clock).
  frame                 STRONG scaling, BASELINE's second metric: wall time to a `--frame-spp` (10 000) spp frame of the workload
                        with the samples split over the N GPUs — render + reduce + read-back, host wall clock, max over ranks.
                        Two legs: "ranks" (one process per GPU, torch.distributed / NCCL reduce) and "handle" (rank 0 alone
                        drives all N GPUs through ONE rt2 handle, rt2_config.n_gpus = N: the path `raytrace_2 <scene>` takes).
  configs               Mrays/s and paths/s of BASELINE configs C1, C2, C3 and C5 (1 M spheres) in the same run, all ranks.
  roofline              the dominant kernel (the world pass of the extend stage) against the roofline that BINDS it — FP32 /
                        instruction issue: credited flops of the algorithmic work done by active lanes (device counters) over
                        its CUDA-event time, against the FP32 FMA peak measured on this GPU.  roofline_hbm and roofline_l2 give the
                        same kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json) and the L2 bandwidth measured
                        here.  `traffic` comes from this round's committed ncu capture (profiles/), else null.
  cpu_baseline          the reference CPU renderer timed on this box's host cores (N = 1 only, bounded sample).
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
# algorithmic HBM bytes per ray of the extend kernel: ray origin+time 16 + direction 16 read, closest-surface
# record 16 written
EXTEND_BYTES_PER_RAY = 16 + 16 + 16
# DRAM traffic of the extend kernel per ray, from this round's `ncu --set full` capture (tools/ncu_extract.py writes the file)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_extend_traffic.json")
# arithmetic credited per unit of algorithmic work: AABB slab test 30, sphere test 30 (to the
# discriminant; a lower bound), quad test 57, instance visit 42
FLOP_AABB, FLOP_SPHERE, FLOP_QUAD, FLOP_INSTANCE = 30, 30, 57, 42
# scene-fetch bytes per unit of traversal work: one 64-byte node pair (32 B when the scene uses the quantised
# pairs, rt2_stats.compact_nodes), 32 B per sphere, 80 B per quad
NODE_PAIR_BYTES, SPHERE_BYTES, QUAD_BYTES = 64, 32, 80


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scene", default=DEFAULT_SCENE, help="scene name under data/ (or 'synthetic:<n_spheres>')")
    ap.add_argument("--spp-per-step", type=int, default=0, help="samples per pixel per step = one wavefront batch (0 = ~184 M paths: 512 at 600x600)")
    ap.add_argument("--spp-total", type=int, default=10000, help="samples-per-pixel setting (fixes the stratification grid)")
    ap.add_argument("--max-depth", type=int, default=50)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--fast-math", action="store_true", help="FMA-contracted intersection arithmetic (not bit-exact)")
    ap.add_argument("--flags", type=int, default=0, help="extra RT2_FLAG_* bits (A/B runs, e.g. 64 = no instance split)")
    ap.add_argument("--frame-spp", type=int, default=10000, help="strong-scaling leg: spp of the whole frame (0 = skip)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C2/C3/C5 table")
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


def is_synthetic(name):
    return name.startswith("synthetic:")


def load_scene(rt, name, width=0, height=0):
    if is_synthetic(name):
        n = int(name.split(":")[1])
        # large instances skip the host SAH build (seconds per million spheres): the BVH is built on the device (scene_flags adds
        # RT2_FLAG_GPU_LBVH for them)
        return (rt.Scene.synthetic_spheres(n, width=width or 3840, height=height or 2160, host_bvh=n < 500_000),
                f"synthetic {n}-sphere BVH stress scene")
    path = os.path.join(ROOT, "data", name + ".json")
    return rt.Scene.load(path, data_dir=os.path.join(ROOT, "data")), f"data/{name}.json"


def scene_flags(rt, name, args):
    flags = (rt.RT2_FLAG_FAST_MATH if args.fast_math else 0) | args.flags
    if is_synthetic(name) and int(name.split(":")[1]) >= 500_000:
        flags |= rt.RT2_FLAG_GPU_LBVH
    return flags


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
    spp = max(1, (args.spp_per_step or 512) // 64)  # bounded sample per step: the CPU is ~700x slower than one B200
    if is_synthetic(args.scene):
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
        # (the same `workload` string as the GPU arm prints: same scene file, size and depth; a step here is a bounded sample of it)
        "config": {"workload": f"data/{args.scene}.json {wh[0]}x{wh[1]}, max_depth {args.max_depth}",
                   "step": f"{spp} spp per step on the host CPU ({threads} threads)", "spp_per_step": spp, "paths_per_s": paths / total_sec},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind,
                         "sample": f"{args.steps} x {spp} spp at {wh[0]}x{wh[1]}, {total_sec:.1f} s of CPU work"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frame": {"spp": args.frame_spp, "wall_s": (wh[0] * wh[1] * args.frame_spp) / (paths / total_sec) if paths else None,
                  "extrapolated": True, "note": "cost is linear in spp (one parallel region per sample, RayTracer.cpp:69)"},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing: NCCL for the data path, a gloo group for host-side waits that must not occupy a GPU."""

    def __init__(self, torch, dist, world, rank, local_rank):
        self.torch, self.dist, self.world, self.rank = torch, dist, world, rank
        self.dev = torch.device("cuda", local_rank)
        self.cpu_group = None
        if world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=self.dev)
            self.cpu_group = dist.new_group(backend="gloo")

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def sum_max(self, sums, maxs):
        t = self.torch
        a = t.tensor([float(x) for x in sums], dtype=t.float64, device=self.dev)
        b = t.tensor([float(x) for x in maxs], dtype=t.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(a, op=self.dist.ReduceOp.SUM)
            self.dist.all_reduce(b, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in a.tolist()], [float(x) for x in b.tolist()]


def accum_tensor(torch, tracer, dev):
    from raytrace2_b200.distributed import _CudaBuffer
    ptr, n = tracer.accum_device_ptr()
    return torch.as_tensor(_CudaBuffer(ptr, n), device=dev)


def auto_spp_per_step(w, h):
    """One wavefront batch of the library's default size: ~192 M paths (Renderer::AllocState)."""
    return max(1, min(512, (192 * 1024 * 1024 + w * h - 1) // (w * h)))


def time_config(rt, torch, D, args, name, dims, spp_step, steps=2, warmup=1):
    """One row of the configs table: aggregate Mrays/s and paths/s of `name` over all ranks (weak: spp_step per GPU and step)."""
    scene, label = load_scene(rt, name, *(dims or (0, 0)))
    tracer = rt.RayTracer(scene, num_samples=args.spp_total, max_depth=args.max_depth, device=D.dev.index, seed=20261018 + D.rank,
                          flags=scene_flags(rt, name, args), frames_per_batch=spp_step, frame_offset=D.rank, frame_stride=D.world, dims=dims)
    W, H = tracer.Dims()
    ext = torch.cuda.ExternalStream(tracer.stream(), device=D.dev)
    with torch.cuda.stream(ext):
        for _ in range(warmup):
            tracer.Update(spp_step)
        D.barrier()
        st0 = tracer.stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ext)
        for _ in range(steps):
            tracer.Update(spp_step)
        ev1.record(ext)
        D.barrier()
        ms = ev0.elapsed_time(ev1)
        st1 = tracer.stats()
    (rays, paths), (ms_all,) = D.sum_max([st1["rays"] - st0["rays"], st1["paths"] - st0["paths"]], [ms])
    row = {"workload": f"{label} {W}x{H}, {spp_step} spp per step per GPU", "Mrays_per_s": rays / (ms_all * 1e-3) * 1e-6,
           "paths_per_s": paths / (ms_all * 1e-3), "rays_per_path": rays / max(paths, 1), "ms_per_step": ms_all / steps,
           "bvh_build_ms": st1["gpu_ms_bvh_build"], "instance_mode": st1["instance_mode"], "max_stack_need": st1["max_stack_need"],
           "compact_nodes": st1["compact_nodes"], "node_inflation": round(st1["node_inflation"], 4), "stack_overflows": st1["stack_overflows"]}
    del tracer, scene
    return row


def run_ours(args):
    import torch
    import torch.distributed as dist

    import raytrace2_b200 as rt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or rt.load_library().rt2_device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device — the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    D = Dist(torch, dist, world, rank, local_rank)
    dev = D.dev
    lib = rt.load_library()

    dims = (args.width, args.height) if args.width and args.height else None
    scene, scene_label = load_scene(rt, args.scene, *(dims or (0, 0)))
    flags = scene_flags(rt, args.scene, args)
    sw, sh = dims if dims else scene.dims
    S = args.spp_per_step or auto_spp_per_step(sw, sh)
    tracer = rt.RayTracer(scene, num_samples=args.spp_total, max_depth=args.max_depth, device=local_rank, seed=20261018,
                          flags=flags, frames_per_batch=S, frame_offset=rank, frame_stride=world, dims=dims)
    W, H = tracer.Dims()
    ext = torch.cuda.ExternalStream(tracer.stream(), device=dev)
    accum = accum_tensor(torch, tracer, dev)
    scratch = torch.empty_like(accum)

    def step():
        tracer.Update(S)  # S = one full batch: traced at once
        if world > 1:
            scratch.copy_(accum, non_blocking=True)
            dist.reduce(scratch, dst=0, op=dist.ReduceOp.SUM)

    with torch.cuda.stream(ext):
        for _ in range(max(args.warmup, 3)):
            step()
        D.barrier()
        st0 = tracer.stats()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(ext)
        for _ in range(args.steps):
            step()
        ev1.record(ext)
        D.barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = ev0.elapsed_time(ev1)
        st1 = tracer.stats()
    (rays_all, paths_all, launches_all), (ms_all,) = D.sum_max(
        [st1["rays"] - st0["rays"], st1["paths"] - st0["paths"], st1["launches"] - st0["launches"]], [ms])

    # ---- end-to-end through the public C-ABI path with host buffers: upload scene, render, read the mean image back ----
    def e2e_step():
        tracer.upload_scene()           # H2D: flattened scene from host memory
        tracer.Reset()
        tracer.Update(S)
        img = tracer.NonConvertedPixels()  # D2H: W*H*3 floats (synchronises)
        return tracer.stats()["rays"], img  # Reset() zeroed the counters: rays of this step

    for _ in range(2):
        e2e_step()
    D.barrier()
    e2e_rays = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_rays += e2e_step()[0]
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    (e2e_rays_all,), (e2e_sec_all,) = D.sum_max([e2e_rays], [e2e_sec])
    d = scene.desc
    scene_bytes = (d.n_spheres * 32 + d.n_quads * 80 + d.n_xforms * 96 + d.n_instances * 16 + d.n_media * 32 + d.n_materials * 32 +
                   d.n_textures * 48 + d.n_perlin * 7168 + d.n_prim_refs * 4 + d.n_node_pairs * 64)
    e2e = {"value": e2e_rays_all / e2e_sec_all * 1e-6, "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes),
           "d2h_bytes_per_step": int(W * H * 3 * 4),
           "host_memory": "pinned staging arena inside the library (cudaMallocHost), host buffers on both sides of the C ABI"}

    # ---- rooflines of the dominant kernel: per-kernel CUDA-event split over extra (untimed) profiled steps ----
    tracer.Reset()
    tracer.set_profiling(True)
    prof_steps = min(args.steps, 4)
    for _ in range(prof_steps):
        tracer.Update(S)
    ps = tracer.stats()
    tracer.set_profiling(False)
    peaks, peak_src = measured_peaks()
    ext_ms, inst_ms = ps["gpu_ms_extend"], ps["gpu_ms_extend_inst"]
    prof_total = ext_ms + inst_ms + ps["gpu_ms_shade"] + ps["gpu_ms_other"] + ps["gpu_ms_finish"] + ps["gpu_ms_sort"]
    n_trav_launches = max(1, prof_steps * args.max_depth)  # one world-pass launch per bounce and batch
    rays_per_launch = ps["rays"] / n_trav_launches
    mode = ps.get("instance_mode", 0)
    if ps["box_pair_tests"] == 0:
        kernel_name = "k_traverse_flat"
    elif mode == 2:
        kernel_name = "k_traverse<kTravWorld> (+ k_traverse<kTravInst>)"
    elif mode == 3:
        kernel_name = "k_traverse<kTravUnified>"
    else:
        kernel_name = "k_traverse<kTravInline>"
    compact = bool(ps.get("compact_nodes", 0))
    if compact and kernel_name.startswith("k_traverse<"):
        kernel_name = kernel_name[:-1] + ", kQuant>"
    node_pair_bytes = 32 if compact else NODE_PAIR_BYTES
    trav_ms = ext_ms + inst_ms  # both passes of the extend stage
    fp32_peak, l2_peak = C.c_double(0.0), C.c_double(0.0)
    lib.rt2_measure_fp32_peak(local_rank, C.byref(fp32_peak))
    lib.rt2_measure_l2_bandwidth(local_rank, C.byref(l2_peak))
    flops = (2 * ps["box_pair_tests"] * FLOP_AABB + ps["sphere_tests"] * FLOP_SPHERE + ps["quad_tests"] * FLOP_QUAD +
             ps["instance_visits"] * FLOP_INSTANCE)
    ach_tf = flops / (trav_ms * 1e-3) * 1e-12 if trav_ms > 0 else 0.0
    per_ray = {"aabb_tests": 2 * ps["box_pair_tests"] / max(ps["rays"], 1), "sphere_tests": ps["sphere_tests"] / max(ps["rays"], 1),
               "quad_tests": ps["quad_tests"] / max(ps["rays"], 1), "instance_entries": ps["instance_visits"] / max(ps["rays"], 1)}
    traffic, traffic_src = None, "no ncu capture of this round's kernel committed yet (profiles/r02_extend_traffic.json)"
    if os.path.exists(TRAFFIC_FILE) and kernel_name.startswith("k_traverse<"):
        try:
            tj = json.load(open(TRAFFIC_FILE))
            traffic = float(tj["dram_bytes_per_ray"]) * rays_per_launch
            traffic_src = f"{tj.get('source', TRAFFIC_FILE)}: {tj['dram_bytes_per_ray']:.1f} B/ray (ncu dram__bytes_read+write) x mean rays per launch"
        except Exception as e:  # noqa: BLE001
            traffic_src = f"unreadable {TRAFFIC_FILE}: {e}"
    roofline = {"kernel": kernel_name, "bound": "fp32", "achieved": ach_tf, "peak": fp32_peak.value, "unit": "TFLOP/s",
                "frac": ach_tf / fp32_peak.value if fp32_peak.value > 0 else None,
                "traffic": traffic, "traffic_note": traffic_src,
                "peak_source": "measured here (rt2_measure_fp32_peak: FMA micro-benchmark, 2 flop per FMA)",
                "flop_per_ray": flops / max(ps["rays"], 1), "per_ray": per_ray,
                "avg_launch_ms": trav_ms / n_trav_launches, "rays_per_launch": rays_per_launch,
                "share_of_step": trav_ms / prof_total if prof_total > 0 else None,
                "note": "the extend stage runs out of L1: ncu shows l1tex throughput at 88 % of peak (float nodes), 58-65 % of issue slots busy "
                        "at 13-17 of 32 lanes and long-scoreboard as the top stall (profiles/r02_notes.md) — SIMT divergence on "
                        "cache-resident data, not DRAM.  Of the rooflines bench.py can measure live the FP32 one is the closest: credited flops per "
                        "SURVEY Appendix C (30 / AABB test, 30 / sphere, 57 / quad, 42 / instance entry) of the work done by ACTIVE "
                        "lanes (device counters); 0 for the flat extend kernel, which has no counters"}
    ach_hbm = ps["rays"] * EXTEND_BYTES_PER_RAY / (trav_ms * 1e-3) * 1e-9 if trav_ms > 0 else 0.0
    roofline_hbm = {"kernel": kernel_name, "bound": "hbm", "achieved": ach_hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach_hbm / peaks["hbm_gbs"], "peak_source": peak_src, "algorithmic_bytes_per_ray": EXTEND_BYTES_PER_RAY,
                    "algorithmic_bytes_per_launch": EXTEND_BYTES_PER_RAY * rays_per_launch,
                    "note": "not the bound: the ray queues stream once, the scene is cache resident"}
    fetch_bytes = ps["box_pair_tests"] * node_pair_bytes + ps["sphere_tests"] * SPHERE_BYTES + ps["quad_tests"] * QUAD_BYTES
    ach_l2 = fetch_bytes / (trav_ms * 1e-3) * 1e-9 if trav_ms > 0 else 0.0
    roofline_l2 = {"kernel": kernel_name, "bound": "l2", "achieved": ach_l2, "peak": l2_peak.value, "unit": "GB/s",
                   "frac": ach_l2 / l2_peak.value if l2_peak.value > 0 else None,
                   "peak_source": "measured here (rt2_measure_l2_bandwidth: 24 MiB working set re-read with 16-byte loads)",
                   "scene_fetch_bytes_per_ray": fetch_bytes / max(ps["rays"], 1),
                   "node_pair_bytes": node_pair_bytes,
                   "note": "scene fetches issued by active lanes (64 B per float node pair / 32 B per quantised pair, 32 B per sphere, 80 B per quad) — an upper bound of "
                           "the L2 traffic: most of them hit L1 for scenes of a few hundred KB"}
    kernel_split = {"extend_world_ms": ext_ms, "extend_instances_ms": inst_ms, "finish_shade_ms": ps["gpu_ms_finish"],
                    "deferred_shade_ms": ps["gpu_ms_shade"], "sort_ms": ps["gpu_ms_sort"], "other_ms": ps["gpu_ms_other"], "steps": prof_steps}

    # ---- strong scaling: wall time to the whole frame ----
    frame = None
    if args.frame_spp > 0:
        from raytrace2_b200.distributed import frame_partition
        _, _, local_frames = frame_partition(args.frame_spp, rank, world)
        frame = {"spp": args.frame_spp, "workload": f"{scene_label} {W}x{H}, max_depth {args.max_depth}"}

        def ranks_leg():
            tracer.Reset()
            D.barrier()
            t0 = time.perf_counter()
            tracer.Update(local_frames)
            tracer.flush()
            t1 = time.perf_counter()
            with torch.cuda.stream(ext):
                if world > 1:
                    scratch.copy_(accum, non_blocking=True)
                    dist.reduce(scratch, dst=0, op=dist.ReduceOp.SUM)
                    img = (scratch.view(H, W, 4)[..., :3] / float(args.frame_spp)).cpu() if rank == 0 else None
                else:
                    img = tracer.NonConvertedPixels()
            tracer.synchronize()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            st = tracer.stats()
            return t2 - t0, t1 - t0, st["rays"], st["paths"], st["gpu_ms_total"], img

        if world > 1:
            ranks_leg()  # warm NCCL's reduce path at this size
            wall, issue, r, p, gpu_ms, _ = ranks_leg()
            (r_all, p_all), (wall_all, gpu_ms_all) = D.sum_max([r, p], [wall, gpu_ms])
            frame["ranks"] = {"wall_s": wall_all, "render_gpu_s": gpu_ms_all * 1e-3, "reduce_readback_s": max(wall_all - gpu_ms_all * 1e-3, 0.0),
                              "Mrays_per_s": r_all / wall_all * 1e-6, "paths_per_s": p_all / wall_all, "rays": r_all,
                              "how": "one process per GPU, frames f = rank (mod N), NCCL sum-reduce to rank 0, mean image to the host"}
        # one handle over all N GPUs, driven by rank 0 alone (what `raytrace_2 <scene>` does); the other ranks wait on the host
        D.barrier()
        if rank == 0:
            h = rt.RayTracer(scene, num_samples=args.frame_spp, max_depth=args.max_depth, device=0, seed=20261018, flags=flags, dims=dims,
                             n_gpus=world)
            h.Update(min(args.frame_spp, 2 * world))  # warm-up: first launches, peer mappings
            h.NonConvertedPixels()
            h.Reset()
            t0 = time.perf_counter()
            for _ in range(args.frame_spp):
                h.Update(1)  # the reference's own loop (App.cpp:243-248): one Update per sample
            img = h.NonConvertedPixels()
            wall = time.perf_counter() - t0
            st = h.stats()
            frame["handle"] = {"wall_s": wall, "render_gpu_s": st["gpu_ms_total"] * 1e-3, "reduce_readback_s": max(wall - st["gpu_ms_total"] * 1e-3, 0.0),
                               "Mrays_per_s": st["rays"] / wall * 1e-6, "paths_per_s": st["paths"] / wall, "rays": st["rays"],
                               "n_gpus": st["n_gpus"], "update_calls": args.frame_spp, "mean_radiance": float(img.mean()),
                               "how": "ONE rt2 handle with rt2_config.n_gpus = N in one process: frames dealt round-robin to the replicas, "
                                      "read-out = one kernel on GPU 0 summing the peers' accumulators over NVLink"}
            frame["wall_s"] = min(wall, frame["ranks"]["wall_s"]) if "ranks" in frame else wall
            del h
        D.host_barrier()

    # ---- the other BASELINE configs in the same run ----
    configs = None
    if not args.no_configs:
        configs = {}
        table = [("C1", "cornell_original_test", None, 256), ("C2", "final_render_book_1", None, 64),
                 ("C3", "cornell_volume_10000_samples", None, 256), ("C5_1M", "synthetic:1000000", (3840, 2160), 8)]
        for key, name, cdims, spp_step in table:
            try:
                configs[key] = time_config(rt, torch, D, args, name, cdims, spp_step)
            except Exception as e:  # noqa: BLE001 — a failing side config must not lose the headline line
                configs[key] = {"error": str(e)}
        configs["C4"] = {"workload": f"{scene_label} {W}x{H}, {S} spp per step per GPU", "Mrays_per_s": rays_all / (ms_all * 1e-3) * 1e-6,
                         "paths_per_s": paths_all / (ms_all * 1e-3), "rays_per_path": rays_all / max(paths_all, 1), "ms_per_step": ms_all / args.steps}

    line = {
        "metric": "Mrays/s", "value": rays_all / (ms_all * 1e-3) * 1e-6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic camera samples (Philox) on the repo's scene file; random Perlin tables",
        "config": {"workload": f"{scene_label} {W}x{H}, max_depth {args.max_depth}",
                   "step": f"{S} spp (one wavefront batch) per step per GPU", "spp_per_step_per_gpu": S, "paths_per_s": paths_all / (ms_all * 1e-3), "rays_per_path": rays_all / max(paths_all, 1),
                   "intersection_math": "fast (FMA)" if args.fast_math else "exact (bit-identical with the reference)",
                   "flags": flags, "instance_mode": st1["instance_mode"],
                   "compact_nodes": st1["compact_nodes"], "node_inflation": round(st1["node_inflation"], 4),
                   "l2": "wavefront state per step (%.0f MB) exceeds L2; no explicit flush" % (W * H * S * 184 / 1e6),
                   "parallelism": f"sample-partition x{world}" if world > 1 else "single GPU"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_all), "roofline": roofline, "roofline_hbm": roofline_hbm,
        "roofline_l2": roofline_l2, "kernel_split_profiled": kernel_split, "frame": frame, "configs": configs,
        "stack_overflows": st1["stack_overflows"],
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline and not is_synthetic(args.scene):
        threads = os.cpu_count() or 1
        spp_cpu = args.cpu_sample_spp or max(2, min(128, int(round(15 * threads / 8))))  # ~10 s of wall clock on all host threads
        mr, pps, kind, sec, crays, wh = reference_sample(args.scene, dims, spp_cpu, args.spp_total, args.max_depth, threads)
        line["cpu_baseline"] = {"value": mr, "unit": "Mrays/s", "cores": threads, "kind": kind, "paths_per_s": pps,
                                "sample": f"{spp_cpu} spp of the same workload at {wh[0]}x{wh[1]} ({sec:.1f} s)",
                                "frame_wall_s_extrapolated": wh[0] * wh[1] * args.frame_spp / pps if (pps and args.frame_spp) else None}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        D.host_barrier()
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

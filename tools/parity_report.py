#!/usr/bin/env python3
"""Image-parity + throughput report (north_star: per-pixel means within Monte-Carlo noise at equal spp, RMSE / PSNR against a
10k-spp converged reference render, Mrays/s, paths/s, wall time to a 10k-spp frame, next to the reference CPU renderer on
this box's host cores).  Writes JSON + Markdown under --out (default profiles/).

The converged reference is rendered by the reference's own code (oracle/_ref; the CPU port where HEAD's loader cannot read the
file) at a reduced resolution so it finishes in minutes on the host CPU; the GPU renders the same resolution / spp.
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytrace2_b200 as rt
from raytrace2_b200 import parity
from oracle import ref_oracle, rt_oracle

SCENES = [("cornell_original_test", "C1"), ("final_render_book_1", "C2"), ("cornell_volume_10000_samples", "C3"),
          ("book2_final_scene_10000_samples", "C4")]


def cpu_scene(path, spp, dims):
    if ref_oracle.available():
        try:
            return ref_oracle.RefScene(path, spp, dims=dims), "reference"
        except RuntimeError:
            pass
    return rt_oracle.PortScene(path, spp, dims=dims), "port"


def share_perlin(scene, cpu, kind):
    if scene.desc.n_perlin:
        tables = scene.get_perlin(0)
        texs = [i for i, t in enumerate(scene.textures()) if int(t["type"]) == 2]
        for ti in texs:
            cpu.perlin_set(ti, *tables)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles"))
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--conv-spp", type=int, default=10000)
    ap.add_argument("--conv-dim", type=int, default=120)
    ap.add_argument("--eq-spp", type=int, default=256)
    ap.add_argument("--full-spp", type=int, default=1024, help="spp of the full-resolution GPU throughput run")
    a = ap.parse_args()
    threads = os.cpu_count() or 1
    rows = []
    for name, tag in SCENES:
        path = os.path.join(ROOT, "data", name + ".json")
        scene = rt.Scene.load(path)
        W, H = scene.dims
        dims = (a.conv_dim, max(1, int(round(a.conv_dim * H / W))))
        # --- equal-spp z-score test + converged reference at reduced resolution
        cpu, kind = cpu_scene(path, a.conv_spp, dims)
        share_perlin(scene, cpu, kind)
        t0 = time.time()
        cs, css, crays, csec = cpu.render(0, a.conv_spp, 50, 0, True)
        cpu_mrays = crays / csec * 1e-6
        conv = cs / a.conv_spp
        tr = rt.RayTracer(scene, num_samples=a.conv_spp, seed=7, flags=rt.RT2_FLAG_MOMENTS, dims=dims)
        tr.Update(a.conv_spp)
        gs, gss = tr.read_accum(moments=True)
        z, valid = parity.z_scores(gs, gss, a.conv_spp, cs, css, a.conv_spp)
        zs = parity.summary(z, valid)
        tz = parity.tile_z_scores(gs, gss, a.conv_spp, cs, css, a.conv_spp, tile=max(8, a.conv_dim // 6))
        rm_full, ps_full = parity.rmse_psnr(gs / a.conv_spp, conv)
        # lower-spp GPU renders against the converged CPU image
        conv_curve = {}
        for spp in (16, 64, 256, 1024):
            t2 = rt.RayTracer(scene, num_samples=spp, seed=11, dims=dims)
            t2.Update(spp)
            r, p = parity.rmse_psnr(t2.NonConvertedPixels(), conv)
            conv_curve[spp] = {"rmse": r, "psnr_db": p}
        # --- full-resolution GPU throughput
        full = rt.RayTracer(scene, num_samples=10000, seed=3)
        full.Update(64)
        full.synchronize()
        full.Reset()
        full.Update(a.full_spp)
        st = full.stats()
        gpu_mrays = st["rays"] / st["gpu_ms_total"] * 1e-3
        gpu_paths = st["paths"] / st["gpu_ms_total"] * 1e3
        rows.append({
            "config": tag, "scene": name, "resolution": [W, H], "cpu_kind": kind, "cpu_threads": threads,
            "cpu_Mrays_s": cpu_mrays, "cpu_paths_s": dims[0] * dims[1] * a.conv_spp / csec, "cpu_rays_per_path": crays / (dims[0] * dims[1] * a.conv_spp),
            "gpu_Mrays_s": gpu_mrays, "gpu_paths_s": gpu_paths, "gpu_rays_per_path": st["rays"] / st["paths"],
            "gpu_seconds_per_10k_spp_frame": W * H * 10000 / gpu_paths,
            "cpu_seconds_per_10k_spp_frame_extrapolated": W * H * 10000 / (dims[0] * dims[1] * a.conv_spp / csec),
            "z_test": {"dims": dims, "spp_each": a.conv_spp, **zs, "tile_abs_z_max": float(np.abs(tz).max())},
            "gpu_vs_converged_cpu": {"dims": dims, "spp": a.conv_spp, "rmse": rm_full, "psnr_db": ps_full, "by_gpu_spp": conv_curve},
        })
        print(json.dumps(rows[-1]), flush=True)
        del tr, full
    os.makedirs(a.out, exist_ok=True)
    with open(os.path.join(a.out, f"{a.tag}_parity_report.json"), "w") as f:
        json.dump(rows, f, indent=1)
    with open(os.path.join(a.out, f"{a.tag}_parity_report.md"), "w") as f:
        f.write(f"# Parity + throughput report ({a.tag})\n\nOne B200 vs the reference CPU renderer on this box ({threads} host threads). "
                f"Converged reference: {a.conv_spp} spp at {a.conv_dim} px wide, rendered by the reference's own code where its loader reads the file "
                "(`reference`), else by the pinned CPU port (`port`).  z-test: GPU vs CPU at equal spp, per pixel and channel.\n\n")
        f.write("| cfg | scene | CPU Mrays/s | GPU Mrays/s | GPU paths/s | rays/path CPU / GPU | GPU s per 10k-spp frame | CPU s (extrap.) | mean z | std z | P(abs z>3) | tile max abs z | RMSE / PSNR @10k vs CPU 10k |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            zt, cv = r["z_test"], r["gpu_vs_converged_cpu"]
            f.write(f"| {r['config']} | {r['scene']} ({r['cpu_kind']}) | {r['cpu_Mrays_s']:.2f} | {r['gpu_Mrays_s']:.0f} | {r['gpu_paths_s']:.3g} | "
                    f"{r['cpu_rays_per_path']:.3f} / {r['gpu_rays_per_path']:.3f} | {r['gpu_seconds_per_10k_spp_frame']:.2f} | "
                    f"{r['cpu_seconds_per_10k_spp_frame_extrapolated']:.0f} | {zt['mean_z']:+.4f} | {zt['std_z']:.3f} | {zt['frac_gt3']:.4f} | "
                    f"{zt['tile_abs_z_max']:.2f} | {cv['rmse']:.4f} / {cv['psnr_db']:.1f} dB |\n")
        f.write("\nRMSE / PSNR of lower-spp GPU renders against the converged CPU image (same reduced resolution):\n\n| cfg | 16 spp | 64 spp | 256 spp | 1024 spp |\n|---|---|---|---|---|\n")
        for r in rows:
            c = r["gpu_vs_converged_cpu"]["by_gpu_spp"]
            f.write(f"| {r['config']} | " + " | ".join(f"{c[k]['rmse']:.4f} / {c[k]['psnr_db']:.1f} dB" for k in (16, 64, 256, 1024)) + " |\n")


if __name__ == "__main__":
    main()

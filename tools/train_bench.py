"""Training-loop benchmark in the shape of BASELINE.json configs[3] (SURVEY.md section 8(d) "config 4": a few million gaussians,
1600x1060, iterations INCLUDING densify/prune and the optimizer): everything around the rasterizer comes from
gaussian_store.GaussianModel (fused Adam, plan+gather densification), the loss is the reference's 0.8 L1 + 0.2 D-SSIM, and the
loop follows train.py:73-190 (one view per iteration by default, densify every 100 iterations after 500, opacity reset every
3000, SH degree +1 every 1000).  Synthetic scene and random target images (no dataset in this image), so the numbers are
throughput, not quality.

    python tools/train_bench.py --gaussians 6000000 --iterations 1000            # configs[3] size; needs a B200
    python tools/train_bench.py --gaussians 20000 --width 160 --height 96 --iterations 30 --densify-from 10 --densify-interval 10
"""
import argparse
import json
import math
import os
import sys
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
import diff_gaussian_rasterization as dgr  # noqa: E402
from gaussian_renderer import render_views_backward  # noqa: E402
from gaussian_renderer.synthetic import camera_matrices, make_scene, sphere_pose  # noqa: E402
from gaussian_store import GaussianModel  # noqa: E402


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gaussians", type=int, default=6_000_000)
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1060)
    ap.add_argument("--iterations", type=int, default=1000)
    ap.add_argument("--views-per-iteration", type=int, default=1, help="1 = the reference's loop; >1 uses the view-batch path")
    ap.add_argument("--cameras", type=int, default=64)
    ap.add_argument("--densify-from", type=int, default=500)
    ap.add_argument("--densify-until", type=int, default=15000)
    ap.add_argument("--densify-interval", type=int, default=100)
    ap.add_argument("--opacity-reset-interval", type=int, default=3000)
    ap.add_argument("--max-gaussians", type=int, default=12_000_000, help="densification is skipped above this count")
    ap.add_argument("--log-scale-mean", type=float, default=-5.6)
    return ap.parse_args(argv)


def run(a) -> dict:
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    sc = {k: v.to(dev) for k, v in make_scene(a.gaussians, seed=0, log_scale_mean=a.log_scale_mean).items()}
    pc = GaussianModel(3)
    op = sc["opacities"].clamp(1e-6, 1 - 1e-6).reshape(-1, 1)
    pc.create_from_tensors(sc["means3D"], sc["shs"][:, :1].contiguous(), sc["shs"][:, 1:].contiguous(), torch.log(sc["scales"]),
                           sc["rotations"], torch.log(op / (1 - op)), 1.0)
    del sc, op
    opt = SimpleNamespace(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
                          position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005, rotation_lr=0.001,
                          percent_dense=0.01, lambda_dssim=0.2, densify_grad_threshold=0.0002)      # arguments/__init__.py:77-95
    pc.training_setup(opt)
    extent = 1.1 * math.sqrt(3.0)                          # cameras_extent stand-in: radius of the unit cube's bounding sphere
    H, W = a.height, a.width
    cams = []
    for k in range(a.cameras):
        R, T = sphere_pose(k, 3.0)
        wvt, full, center = camera_matrices(R, T, math.radians(60.0), 2 * math.atan(math.tan(math.radians(30.0)) * H / W))
        cams.append(SimpleNamespace(image_height=H, image_width=W, FoVx=math.radians(60.0),
                                    FoVy=2 * math.atan(math.tan(math.radians(30.0)) * H / W), world_view_transform=wvt.to(dev),
                                    full_proj_transform=full.to(dev), camera_center=center.to(dev)))
    gen = torch.Generator(device=dev).manual_seed(1)
    targets = [torch.rand(3, H, W, device=dev, generator=gen) for _ in range(min(a.cameras, 16))]
    pipe = SimpleNamespace(convert_SHs_python=False, compute_cov3D_python=False, debug=False, antialiasing=False)
    bg = torch.zeros(3, device=dev)
    V = a.views_per_iteration
    P_trace, dens_ms, D_max = [pc.P], [], 0
    dgr.reset_launch_count()
    torch.cuda.reset_peak_memory_stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e_start = torch.cuda.Event(enable_timing=True)
    e_start.record()
    for it in range(1, a.iterations + 1):
        pc.update_learning_rate(it)
        if it % 1000 == 0:
            pc.oneupSHdegree()
        views = [cams[(it * V + j) % len(cams)] for j in range(V)]
        tgt = [targets[(it * V + j) % len(targets)] for j in range(V)]
        stats = {"xyz_gradient_accum": pc.xyz_gradient_accum, "denom": pc.denom}
        out = render_views_backward(views, pc, pipe, bg,
                                    lambda img, _d, i: dgr.photometric_loss_and_grad(img, tgt[i], opt.lambda_dssim)[:2],
                                    loss_returns_grad=True, overwrite=True, densify_stats=stats if it < a.densify_until else None)
        D_max = max(D_max, max(out["num_rendered"]))
        if it < a.densify_until:
            torch.maximum(pc.max_radii2D, out["radii_max"].to(pc.max_radii2D.dtype), out=pc.max_radii2D)      # train.py:166
            if it > a.densify_from and it % a.densify_interval == 0 and pc.P < a.max_gaussians:
                torch.cuda.synchronize()
                d0 = time.perf_counter()
                size_threshold = 20 if it > a.opacity_reset_interval else None
                info = pc.densify_and_prune(opt.densify_grad_threshold, 0.005, extent, size_threshold)          # train.py:168-170
                torch.cuda.synchronize()
                dens_ms.append((time.perf_counter() - d0) * 1e3)
                P_trace.append(info["P"])
            if it % a.opacity_reset_interval == 0:
                pc.reset_opacity()
        pc.optimizer_step()      # train.py:178-186; a no-op right after densify_and_prune, as in the reference (grad None)
    e_end = torch.cuda.Event(enable_timing=True)
    e_end.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    gpu_ms = e_start.elapsed_time(e_end)
    return {"tool": "train_bench", "gaussians_start": a.gaussians, "gaussians_end": pc.P, "image": [W, H],
            "iterations": a.iterations, "views_per_iteration": V, "wall_s": round(wall, 3), "gpu_ms": round(gpu_ms, 1),
            "iterations_per_s": round(a.iterations / (gpu_ms / 1e3), 2), "mpix_per_s": round(a.iterations * V * H * W / (gpu_ms / 1e3) / 1e6, 1),
            "densifications": len(dens_ms), "densify_ms_mean": round(sum(dens_ms) / len(dens_ms), 2) if dens_ms else None,
            "densify_ms_max": round(max(dens_ms), 2) if dens_ms else None,
            "P_trace": P_trace[:12], "instances_per_view_max": int(D_max), "loss_last": float(out["losses"].mean()),
            "optimizer_steps": pc.step_count, "gpu_launches": dgr.launch_count(),
            "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}


def bench_line(b) -> None:
    """`python bench.py --workload train6m`: BASELINE.json configs[3] (6 M gaussians, 1600x1060, 1 k iterations with densify /
    prune / opacity reset / fused Adam / L1 + D-SSIM), one line in bench.py's shape.  Not the headline metric (that is the
    rasterizer's Mpix/s): metric here is training iterations per second, one view per iteration as in train.py."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    a = parse([])
    if b.gaussians != 1_000_000:
        a.gaussians = b.gaussians
    if (b.width, b.height) != (1920, 1080):
        a.width, a.height = b.width, b.height
    a.iterations = b.iterations
    r = run(a)
    line = {"metric": "training iterations/s @6M gaussians 1600x1060 (BASELINE.json configs[3])", "value": r["iterations_per_s"],
            "unit": "it/s", "n_gpus": 1, "steps": a.iterations, "warmup": 0, "ms_per_step": r["gpu_ms"] / a.iterations,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{a.gaussians} gaussians, {a.width}x{a.height}, SH 3, {a.iterations} iterations, densify every "
                                   f"{a.densify_interval} after {a.densify_from}, L1 + D-SSIM, fused Adam on the flat store"},
            "gpu_launches": r["gpu_launches"], "train": r}
    print(json.dumps(line), flush=True)


def main():
    print(json.dumps(run(parse())))


if __name__ == "__main__":
    main()

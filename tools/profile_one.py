"""Runs a few bench steps (8 views, batched view path) of the bench workload for ncu.  Usage: python tools/profile_one.py [steps]"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
import torch
import bench
from gaussian_renderer.synthetic import make_scene
import diff_gaussian_rasterization as dgr
from gaussian_renderer import GradientBucket, render_views_backward

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
V = int(os.environ.get("GS_V", "8"))
P = int(os.environ.get("GS_P", "1000000"))
W, H = int(os.environ.get("GS_W", "1920")), int(os.environ.get("GS_H", "1080"))
for kv in os.environ.get("GS_OPTS", "").split(","):
    if "=" in kv:
        k, v = kv.split("="); dgr.set_option(k, int(v))
dev = torch.device("cuda", 0)
scene = make_scene(P, seed=0, log_scale_mean=bench.LOG_SCALE_MEAN)
pc = bench.BenchGaussians(scene, 3, dev)
bucket = GradientBucket(pc.parameters())
bg = torch.zeros(3, device=dev)
gts = [torch.rand(3, H, W, device=dev) for _ in range(V)]
cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), dev) for i in range(V)]
for _ in range(steps):
    bucket.zero_()
    render_views_backward(cams, pc, bench.Pipe(), bg, lambda img, d, i: dgr.l1_loss_and_grad(img, gts[i]), loss_returns_grad=True)
torch.cuda.synchronize()
print("done", steps)

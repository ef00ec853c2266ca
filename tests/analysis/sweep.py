"""A/B sweep of kernel variants on the bench workload: correctness vs the CPU oracle at a medium size, then
per-kernel CUDA-event times at full size.  python tests/analysis/sweep.py "fwd,bwd;fwd,bwd;..." """
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "gaussian-splatting_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
import gs_test_util as U
from oracle import torch_oracle as TO
import diff_gaussian_rasterization as dgr
from gaussian_renderer import render

combos = [tuple(int(x) for x in c.split(",")) for c in (sys.argv[1] if len(sys.argv) > 1 else "0,1;2,2;3,3").split(";")]
dev = torch.device("cuda", 0)
# correctness
scene_s = TO.make_scene(20000, seed=3, log_scale_mean=-3.6)
cam_s = TO.make_camera(400, 240, sh_degree=3, bg=(0.2, 0.4, 0.1))
args = U.make_args(scene_s, "sh")
gen = torch.Generator().manual_seed(9)
wc, wd = torch.randn(3, 240, 400, generator=gen).numpy(), torch.randn(1, 240, 400, generator=gen).numpy()
ref = U.run_oracle(args, cam_s, wc, wd)
# timing
P, W, H = 1_000_000, 1920, 1080
scene = TO.make_scene(P, seed=0, log_scale_mean=bench.LOG_SCALE_MEAN)
pc = bench.BenchGaussians(scene, 3, dev)
bg = torch.zeros(3, device=dev)
gt = torch.rand(3, H, W, device=dev)
cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), dev) for i in range(4)]
for combo in combos:
    fv, bv = combo[0], combo[1]
    sv = combo[2] if len(combo) > 2 else 1
    dgr.set_option("render_fwd_variant", fv); dgr.set_option("render_bwd_variant", bv); dgr.set_option("sort_variant", sv)
    out = {"fwd": fv, "bwd": bv, "sort": sv}
    try:
        got = U.run_cuda(args, cam_s, wc, wd)
        out["img_max_err"] = float(np.abs(got["color"] - ref["color"]).max())
        out["img_frac_gt_1e-5"] = float((np.abs(got["color"] - ref["color"]) > 1e-5).mean())
        out["grad_rel"] = {k: float(np.abs(got["grads"][k].reshape(v.shape) - v).max() / (np.abs(v).max() + 1e-20))
                           for k, v in ref["grads"].items() if v is not None}
    except Exception as e:
        out["error"] = repr(e)[:300]
    for it in range(3):
        if it == 1:
            dgr.set_option("time_kernels", 2); dgr.kernel_time("", reset=True)
        for cam in cams:
            pkg = render(cam, pc, bench.Pipe(), bg)
            (pkg["render"] - gt).abs().mean().backward()
    torch.cuda.synchronize()
    n = 2 * len(cams)
    out["ms"] = {k: round(dgr.kernel_time(k)[0] / n, 4) for k in ("render_fwd", "render_bwd", "preprocess_fwd", "preprocess_bwd", "sort_hist", "sort_rowscan", "sort_scatter", "emit", "tile_ranges")}
    out["ms"]["all_kernels"] = round(dgr.kernel_time("")[0] / n, 4)
    dgr.kernel_time("", reset=True); dgr.set_option("time_kernels", 0)
    print(json.dumps(out), flush=True)

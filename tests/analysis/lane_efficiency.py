"""CPU-only analysis for DESIGN.md's next step #1: on the bench workload (1 M gaussians, 1920x1080, view 0), how many of the
pixel evaluations the blend kernels EXECUTE actually contribute?  Uses the oracle's projection (test infrastructure; this
tool is an analysis aid, not a product or benchmark path) on a window of tiles in the image centre and replays, per tile,
what the kernels do: depth-ordered list, per-pixel termination at T < 1e-4, per-entry 8x4 patch masks.

    python tests/analysis/lane_efficiency.py [--tiles-x 8 --tiles-y 6]
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
import bench  # noqa: E402
from oracle import torch_oracle as TO  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gaussians", type=int, default=1_000_000)
ap.add_argument("--tiles-x", type=int, default=8)
ap.add_argument("--tiles-y", type=int, default=6)
a = ap.parse_args()
W, H = 1920, 1080
scene = TO.make_scene(a.gaussians, seed=0, log_scale_mean=bench.LOG_SCALE_MEAN)
R, T = bench.view_pose(0, 3.0)
fovx = math.radians(60.0)
fovy = 2.0 * math.atan(math.tan(fovx / 2) * H / W)
wvt, full, center = TO.camera_matrices(R, T, fovx, fovy)
cam = TO.OracleSettings(H, W, math.tan(fovx / 2), math.tan(fovy / 2), torch.zeros(3), 1.0, wvt, full, 3, center)
with torch.no_grad():
    pr = TO.project(scene["means3D"], None, scene["shs"], None, scene["opacities"], scene["scales"], scene["rotations"], None, cam)
vis = pr.visible.numpy()
xy, conic, op, depth, rect = (t.numpy() for t in (pr.xy, pr.conic, pr.opacity, pr.depth, pr.rect))
tx0, ty0 = (W // 16 - a.tiles_x) // 2, (H // 16 - a.tiles_y) // 2
tot = dict(entries=0, visited=0, fwd_warp_iters=0, fwd_useful=0, fwd_tile_iters=0, bwd_warp_iters=0, bwd_useful=0, pairs_alpha=0)
for ty in range(ty0, ty0 + a.tiles_y):
    for tx in range(tx0, tx0 + a.tiles_x):
        m = vis & (rect[:, 0] <= tx) & (tx < rect[:, 2]) & (rect[:, 1] <= ty) & (ty < rect[:, 3])
        idx = np.nonzero(m)[0]
        idx = idx[np.argsort(depth[idx], kind="stable")]
        px = (tx * 16 + np.arange(16))[None, :].repeat(16, 0).reshape(-1).astype(np.float32)
        py = (ty * 16 + np.arange(16))[:, None].repeat(16, 1).reshape(-1).astype(np.float32)
        dx = xy[idx, 0][:, None] - px[None, :]
        dy = xy[idx, 1][:, None] - py[None, :]
        power = -0.5 * (conic[idx, 0][:, None] * dx * dx + conic[idx, 2][:, None] * dy * dy) - conic[idx, 1][:, None] * dx * dy
        alpha = np.minimum(0.99, op[idx][:, None] * np.exp(power))
        hit = (power <= 0) & (alpha >= 1.0 / 255.0)                      # [n, 256]
        # exact culling keeps only entries that touch the tile at all (the kernels' lists are a slight superset)
        keep = hit.any(axis=1)
        idx, alpha, hit = idx[keep], alpha[keep], hit[keep]
        n = len(idx)
        # per-pixel transmittance before each entry, termination when the NEXT T would drop below 1e-4
        a_eff = np.where(hit, alpha, 0.0)
        Tbefore = np.cumprod(np.vstack([np.ones((1, 256)), 1.0 - a_eff[:-1]]), axis=0) if n else np.ones((0, 256))
        alive = Tbefore * (1.0 - a_eff) >= 1e-4                           # the entry that would cross the threshold is not blended
        alive = np.logical_and.accumulate(alive, axis=0)
        contrib = hit & alive
        last = np.nonzero(contrib.any(axis=1))[0]
        visited = int(last[-1]) + 1 if len(last) else 0
        # the tile's loop ends when every pixel is done
        done_all = np.nonzero(~alive.any(axis=1))[0]
        if len(done_all):
            visited = min(n, int(done_all[0]) + 1)
        else:
            visited = n
        c = contrib[:visited]
        h = hit[:visited]
        tot["entries"] += n
        tot["visited"] += visited
        tot["pairs_alpha"] += int(c.sum())
        patch = h.reshape(visited, 16, 16).reshape(visited, 4, 4, 2, 8).transpose(0, 1, 3, 2, 4).reshape(visited, 8, 32)   # [v, patch(4 rows x 2 cols), 32 px]
        patch_active = patch.any(axis=2)                                  # forward: one warp per 8x4 patch
        tot["fwd_warp_iters"] += int(patch_active.sum())
        tot["fwd_tile_iters"] += visited * 8
        tot["fwd_useful"] += int(c.sum())
        # backward: one warp per 16x8 half tile (= patches of two patch rows), 2x2 pixels per lane
        half = patch_active.reshape(visited, 2, 4).any(axis=2) if False else np.stack([patch_active[:, :4].any(axis=1), patch_active[:, 4:].any(axis=1)], axis=1)
        tot["bwd_warp_iters"] += int(half.sum())
        tot["bwd_useful"] += int(c.sum())
nt = a.tiles_x * a.tiles_y
print(f"window {a.tiles_x}x{a.tiles_y} tiles at the image centre of view 0; per tile: list {tot['entries'] / nt:.0f} entries, visited {tot['visited'] / nt:.0f}")
print(f"contributing (pixel, gaussian) pairs per tile: {tot['pairs_alpha'] / nt:.0f} = {tot['pairs_alpha'] / max(tot['visited'] * 256, 1) * 100:.1f} % of visited x 256")
print(f"forward : warp = 8x4 patch; warp iterations kept by the patch mask {tot['fwd_warp_iters'] / max(tot['fwd_tile_iters'], 1) * 100:.1f} % of visited x 8; "
      f"lane-pixels that contribute: {tot['fwd_useful'] / max(tot['fwd_warp_iters'] * 32, 1) * 100:.1f} %")
print(f"backward: warp = 16x8 half tile (2x2 px per lane); warp iterations kept {tot['bwd_warp_iters'] / max(tot['visited'] * 2, 1) * 100:.1f} % of visited x 2; "
      f"lane-pixels that contribute: {tot['bwd_useful'] / max(tot['bwd_warp_iters'] * 128, 1) * 100:.1f} %")

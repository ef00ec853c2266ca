"""Statistics of the bench workload's tile lists: patch hit rates and per-pixel validity (runs on the GPU box)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
import torch
import bench
from gaussian_renderer.synthetic import make_scene
import diff_gaussian_rasterization as dgr

dev = torch.device("cuda", 0)
P, W, H = 1_000_000, 1920, 1080
scene = make_scene(P, seed=0, log_scale_mean=bench.LOG_SCALE_MEAN)
pc = bench.BenchGaussians(scene, 3, dev)
cam = bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(0, 3.0), dev)
bg = torch.zeros(3, device=dev)
rs = dgr.GaussianRasterizationSettings(H, W, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), bg, 1.0, cam.world_view_transform,
                                       cam.full_proj_transform, 3, cam.camera_center, False, False, False)
with torch.no_grad():
    _, radii, _, pack = dgr._forward_impl(pc._xyz.detach(), pc._shs.detach(), None, pc._opacity.detach().reshape(-1),
                                          pc._scaling.detach(), pc._rotation.detach(), None, rs)
sv = dgr.state_views(pack, H, W)
splat = pack["geom"][:P * 48].view(torch.float32).view(P, 12)
gx, gy = (W + 15) // 16, (H + 15) // 16
nc = sv["n_contrib"]
pad = torch.zeros(gy * 16, gx * 16, dtype=nc.dtype, device=dev); pad[:H, :W] = nc
tile_nc = pad.view(gy, 16, gx, 16).permute(0, 2, 1, 3).reshape(gy * gx, 16, 16)
tile_max = tile_nc.reshape(gy * gx, 256).max(dim=1).values
ranges = sv["ranges"]; pl = sv["point_list"].long()
g = torch.Generator(device="cpu").manual_seed(0)
tiles = torch.randperm(gy * gx, generator=g)[:400].tolist()
tot_inst = tot_patch_hits = 0
tot_pairs = tot_valid = 0
warp_any = [0, 0]  # per 8x4 patch among hit patches
slot_stats = torch.zeros(5, dtype=torch.long)
for t in tiles:
    n = int(tile_max[t]); s0 = int(ranges[t, 0])
    if n == 0: continue
    ids = pl[s0:s0 + n]
    rec = splat[ids]
    x, y, A, B, C, o = rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3], rec[:, 4], rec[:, 5]
    ox, oy = (t % gx) * 16, (t // gx) * 16
    px = torch.arange(16, device=dev).float() + ox
    py = torch.arange(16, device=dev).float() + oy
    dx = x[:, None, None] - px[None, None, :]
    dy = y[:, None, None] - py[None, :, None]
    power = -0.5 * (A[:, None, None] * dx * dx + C[:, None, None] * dy * dy) - B[:, None, None] * dx * dy
    alpha = torch.clamp_max(o[:, None, None] * torch.exp(power), 0.99)
    pos = torch.arange(1, n + 1, device=dev)[:, None, None]
    valid = (power <= 0) & (alpha >= 1 / 255) & (pos <= tile_nc[t][None])
    tot_inst += n
    tot_pairs += n * 256; tot_valid += int(valid.sum())
    patches = valid.view(n, 4, 4, 2, 8).permute(0, 1, 3, 2, 4).reshape(n, 8, 32).any(dim=2)   # [n, 8 patches] truly-hit (any valid pixel)
    tot_patch_hits += int(patches.sum())
    per_inst = patches.sum(dim=1)
    slot_stats += torch.bincount(per_inst.clamp(max=4).cpu(), minlength=5)[:5]
    half = valid.view(n, 2, 128).any(dim=2)
    warp_any[0] += int(half.sum()); warp_any[1] += 2 * n
print("instances visited (sample):", tot_inst)
print("valid pixel pairs / evaluated pairs: %.3f" % (tot_valid / tot_pairs))
print("patches (8x4) with >=1 valid pixel per instance: %.2f of 8" % (tot_patch_hits / tot_inst))
print("half tiles (16x8) with >=1 valid pixel: %.3f" % (warp_any[0] / warp_any[1]))
print("distribution of #valid patches per instance (0,1,2,3,4+):", (slot_stats.float() / slot_stats.sum()).tolist())
print("valid pairs within valid patches: %.3f" % (tot_valid / max(1, tot_patch_hits * 32)))

"""Times the fused optimizer step and densify/prune of gaussian_store.GaussianModel against the reference formulation on
the same GPU (six nn.Parameters, activations + autograd chain, torch.optim.Adam; boolean-mask / cat surgery).

    python tools/store_bench.py [--gaussians 1000000]
"""
import argparse
import os
import sys
import time
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
import diff_gaussian_rasterization as dgr  # noqa: E402
from gaussian_store import GaussianModel, store_offsets  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gaussians", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
P, M = a.gaussians, 16
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
raw = dict(xyz=rnd(P, 3), f_dc=rnd(P, 1, 3) * 0.5, f_rest=rnd(P, 15, 3) * 0.1, scaling=rnd(P, 3) - 3.6, rotation=rnd(P, 4),
           opacity=torch.rand(P, 1, device=dev, generator=g) * 9 - 6.5)
opt = SimpleNamespace(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=30000,
                      feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005, rotation_lr=0.001, percent_dense=0.01)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---- fused store ----
m = GaussianModel(3).create_from_tensors(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["scaling"], raw["rotation"], raw["opacity"], 4.0)
m.training_setup(opt)
m.grad.copy_(torch.randn(m.grad.shape, device=dev, generator=g) * 1e-3)
t_fused = timed(m.optimizer_step, a.iters)
n_float = store_offsets(P, M)["total"]
bytes_step = n_float * 28 + 8 * P * 4
print(f"fused adam_step      : {t_fused:.3f} ms  ({bytes_step / t_fused / 1e6:.0f} GB/s algorithmic, {n_float / 1e6:.1f} M floats)")

# ---- reference formulation: nn.Parameters + activations + autograd chain + torch.optim.Adam ----
params = {k: torch.nn.Parameter(v.clone()) for k, v in raw.items()}
lrs = dict(xyz=0.00016 * 4, f_dc=0.0025, f_rest=0.0025 / 20, opacity=0.025, scaling=0.005, rotation=0.001)
g_act = {"xyz": rnd(P, 3), "features": rnd(P, 16, 3), "opacity": rnd(P, 1), "scaling": rnd(P, 3), "rotation": rnd(P, 4)}
for label, kw in (("torch Adam (default)", {}), ("torch Adam (fused=True)", {"fused": True})):
    adam = torch.optim.Adam([{"params": [params[k]], "lr": lrs[k], "name": k} for k in lrs], lr=0.0, eps=1e-15, **kw)

    def ref_step():
        act = {"xyz": params["xyz"], "features": torch.cat((params["f_dc"], params["f_rest"]), dim=1),
               "opacity": torch.sigmoid(params["opacity"]), "scaling": torch.exp(params["scaling"]),
               "rotation": torch.nn.functional.normalize(params["rotation"])}
        torch.autograd.backward([act[k] for k in g_act], [g_act[k] for k in g_act])
        adam.step()
        adam.zero_grad(set_to_none=True)

    t_ref = timed(ref_step, a.iters)
    print(f"{label:<21}: {t_ref:.3f} ms  (activations + autograd chain + step; {t_ref / t_fused:.1f}x the fused step)")

# ---- densify / prune ----
denom = torch.randint(0, 4, (P, 1), device=dev, generator=g).float()
accum = torch.rand(P, 1, device=dev, generator=g) * 0.0008 * denom
torch.cuda.synchronize()
ts = []
for rep in range(3):
    mm = GaussianModel(3).create_from_tensors(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["scaling"], raw["rotation"], raw["opacity"], 4.0)
    mm.training_setup(opt)
    mm.xyz_gradient_accum, mm.denom = accum.clone(), denom.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = mm.densify_and_prune(0.0002, 0.005, 4.0, 20)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"fused densify_and_prune: {min(ts):.2f} ms wall (P {P} -> {info['P']}, clone {info['n_clone']}, split {info['n_split']}, pruned {info['n_pruned']})")
dgr.set_option("time_kernels", 2)
mm = GaussianModel(3).create_from_tensors(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["scaling"], raw["rotation"], raw["opacity"], 4.0)
mm.training_setup(opt)
mm.xyz_gradient_accum, mm.denom = accum.clone(), denom.clone()
dgr.kernel_time("", reset=True)
mm.densify_and_prune(0.0002, 0.005, 4.0, 20)
mm.optimizer_step()
for k in ("densify_classify", "densify_scan_reduce", "densify_scan_partials", "densify_scan_apply", "densify_invert", "densify_gather",
          "activate", "adam_step"):
    ms, n = dgr.kernel_time(k)
    print(f"   {k:<22} {ms:.3f} ms / {n} launches")
dgr.set_option("time_kernels", 0)


# reference-style surgery with torch ops (gaussian_model.py:316-469), timing only
def ref_densify():
    ps = {k: v.detach().clone() for k, v in raw.items()}
    ms_ = {k: torch.zeros_like(v) for k, v in ps.items()}
    vs_ = {k: torch.zeros_like(v) for k, v in ps.items()}
    grads = accum / denom
    grads[grads.isnan()] = 0.0
    lim = 0.01 * 4.0
    sel = (torch.norm(grads, dim=-1) >= 0.0002) & (torch.exp(ps["scaling"]).max(dim=1).values <= lim)
    for d in (ps, ms_, vs_):
        for k in d:
            d[k] = torch.cat((d[k], d[k][sel] if d is ps else torch.zeros_like(d[k][sel])), dim=0)
    n0 = grads.shape[0]
    pg = torch.zeros(ps["xyz"].shape[0], device=dev)
    pg[:n0] = grads.squeeze()
    sel = (pg >= 0.0002) & (torch.exp(ps["scaling"]).max(dim=1).values > lim)
    stds = torch.exp(ps["scaling"][sel]).repeat(2, 1)
    samples = torch.normal(mean=torch.zeros_like(stds), std=stds)
    q = torch.nn.functional.normalize(ps["rotation"][sel]).repeat(2, 1)
    new = {k: v[sel].repeat(2, *([1] * (v.dim() - 1))) for k, v in ps.items()}
    new["xyz"] = new["xyz"] + samples * q[:, :1]          # stand-in for the bmm with build_rotation (same traffic)
    new["scaling"] = torch.log(stds / 1.6)
    for d in (ps, ms_, vs_):
        for k in d:
            d[k] = torch.cat((d[k], new[k] if d is ps else torch.zeros_like(new[k])), dim=0)
    keep = ~torch.cat((sel, torch.zeros(2 * int(sel.sum()), dtype=torch.bool, device=dev)))
    for d in (ps, ms_, vs_):
        for k in d:
            d[k] = d[k][keep]
    drop = (torch.sigmoid(ps["opacity"]) < 0.005).squeeze() | (torch.exp(ps["scaling"]).max(dim=1).values > 0.4)
    for d in (ps, ms_, vs_):
        for k in d:
            d[k] = d[k][~drop]
    return ps["xyz"].shape[0]


ts = []
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pn = ref_densify()
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"torch-op surgery       : {min(ts):.2f} ms wall (P -> {pn})")

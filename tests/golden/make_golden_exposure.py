"""Generates tests/golden/reference_exposure.npz: the per-image exposure parameters of the reference's OWN
scene/gaussian_model.py (get_exposure_from_name :136-140, the exposure optimizer / schedule of training_setup :201-211 and
update_learning_rate :213-217) stepped as train.py:178-179 does, on the CPU in this container.  Same loading harness as
make_golden_model.py (empty stand-ins for plyfile / simple_knn, torch.zeros without the device keyword).  The colour
transform inside the loss is the expression of gaussian_renderer/__init__.py:113-115.

    python tests/golden/make_golden_exposure.py
"""
import importlib.util
import os
import sys
import types
from argparse import ArgumentParser

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
_zeros = torch.zeros


def _cpu_zeros(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)


torch.zeros = _cpu_zeros
ply = types.ModuleType("plyfile")
ply.PlyData = ply.PlyElement = object
sys.modules["plyfile"] = ply
knn, knn_c = types.ModuleType("simple_knn"), types.ModuleType("simple_knn._C")
knn_c.distCUDA2 = None
sys.modules["simple_knn"], sys.modules["simple_knn._C"] = knn, knn_c
spec = importlib.util.spec_from_file_location("ref_gaussian_model", os.path.join(REF, "scene", "gaussian_model.py"))
GM = importlib.util.module_from_spec(spec)
spec.loader.exec_module(GM)
from arguments import OptimizationParams  # noqa: E402

opt = OptimizationParams(ArgumentParser())
g = torch.Generator().manual_seed(99)
NAMES = ["cam_a", "cam_b", "cam_c"]
P0 = 5
nnP = torch.nn.Parameter
pc = GM.GaussianModel(1)
pc.spatial_lr_scale = 1.0
pc._xyz = nnP(torch.randn(P0, 3, generator=g))
pc._features_dc = nnP(torch.randn(P0, 1, 3, generator=g))
pc._features_rest = nnP(torch.randn(P0, 3, 3, generator=g))
pc._scaling = nnP(torch.randn(P0, 3, generator=g))
pc._rotation = nnP(torch.randn(P0, 4, generator=g))
pc._opacity = nnP(torch.randn(P0, 1, generator=g))
pc.max_radii2D = torch.zeros(P0)
# create_from_pcd :173-176
pc.exposure_mapping = {name: idx for idx, name in enumerate(NAMES)}
pc.pretrained_exposures = None
pc._exposure = nnP(torch.eye(3, 4)[None].repeat(len(NAMES), 1, 1).requires_grad_(True))
pc.training_setup(opt)

imgs = torch.rand(len(NAMES), 3, 5, 7, generator=g)
tgts = torch.rand(len(NAMES), 3, 5, 7, generator=g)
ITERS = [1, 2, 3, 500, 1500, 29999]                  # the schedule is a function of the iteration number: sample it widely
lrs, exps, order = [], [], []
for k, it in enumerate(ITERS):
    pc.update_learning_rate(it)
    lrs.append(pc.exposure_optimizer.param_groups[0]["lr"])
    i = (2 * k + 1) % len(NAMES)
    order.append(i)
    exposure = pc.get_exposure_from_name(NAMES[i])
    img = torch.matmul(imgs[i].permute(1, 2, 0), exposure[:3, :3]).permute(2, 0, 1) + exposure[:3, 3, None, None]
    (img - tgts[i]).abs().mean().backward()
    pc.exposure_optimizer.step()                      # train.py:178-179
    pc.exposure_optimizer.zero_grad(set_to_none=True)
    exps.append(pc._exposure.detach().numpy().copy())

out = dict(names=np.array(NAMES), imgs=imgs.numpy(), tgts=tgts.numpy(), iters=np.array(ITERS), order=np.array(order),
           lrs=np.array(lrs, dtype=np.float64), exposures=np.stack(exps),
           sched=np.array([opt.exposure_lr_init, opt.exposure_lr_final, opt.exposure_lr_delay_steps, opt.exposure_lr_delay_mult,
                           opt.iterations], dtype=np.float64))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_exposure.npz")
np.savez_compressed(path, **out)
print("wrote", path, "lrs", lrs)

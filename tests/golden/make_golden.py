"""Generates tests/golden/reference_python.npz by IMPORTING the reference's own Python
(/root/reference/utils/{sh_utils,general_utils,graphics_utils}.py) in this container.

The reference hard-codes device="cuda" in general_utils.py:65,83,102; this script runs that
UNMODIFIED code on the CPU by wrapping torch.zeros to drop the device keyword.  Nothing else
is altered.  /root/reference does not exist on the GPU box, hence the committed fixture.

    python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)

_zeros = torch.zeros


def _cpu_zeros(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)


torch.zeros = _cpu_zeros
from utils import general_utils as GU  # noqa: E402
from utils import graphics_utils as GR  # noqa: E402
from utils import sh_utils as SH  # noqa: E402

g = torch.Generator().manual_seed(1234)
P = 1000  # BASELINE.json configs[0]: 1k random gaussians, covariance + SH on CPU
scales = torch.exp(torch.randn(P, 3, generator=g) * 0.7 - 3.0)
rots = torch.randn(P, 4, generator=g)  # NOT normalised: build_rotation normalises
L = GU.build_scaling_rotation(1.7 * scales, rots)
cov = GU.strip_symmetric(L @ L.transpose(1, 2))
R = GU.build_rotation(rots)

out = dict(scales=scales.numpy(), rots=rots.numpy(), scale_modifier=np.float32(1.7), cov6=cov.numpy(), R=R.numpy())

xyz = torch.randn(P, 3, generator=g)
campos = torch.tensor([0.3, -0.2, 4.0])
shs = torch.randn(P, 16, 3, generator=g) * 0.4
dirs = xyz - campos[None]
dirs = dirs / dirs.norm(dim=1, keepdim=True)
out.update(xyz=xyz.numpy(), campos=campos.numpy(), shs=shs.numpy())
for deg in range(4):
    rgb = SH.eval_sh(deg, shs.transpose(1, 2), dirs)
    out[f"sh_rgb_deg{deg}"] = torch.clamp_min(rgb + 0.5, 0.0).numpy()  # gaussian_renderer/__init__.py:79-80

# camera matrices (scene/cameras.py:80-89)
Rc = np.array([[0.8, -0.1, 0.59], [0.2, 0.97, -0.1], [-0.56, 0.2, 0.8]])
Rc, _ = np.linalg.qr(Rc)
T = np.array([0.1, -0.3, 3.5])
fovx, fovy = math.radians(65.0), math.radians(42.0)
wvt = torch.tensor(GR.getWorld2View2(Rc, T, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
proj = GR.getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy).transpose(0, 1)
full = wvt.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0)
center = wvt.inverse()[3, :3]
out.update(cam_R=Rc, cam_T=T, fovx=np.float64(fovx), fovy=np.float64(fovy), world_view=wvt.numpy(),
           full_proj=full.numpy(), cam_center=center.numpy())

# photometric loss (utils/loss_utils.py:40-86): L1 and SSIM values + autograd gradients of the reference's own functions
from utils import loss_utils as LU  # noqa: E402
gl = torch.Generator().manual_seed(99)
img = torch.rand(3, 37, 53, generator=gl, dtype=torch.float32).requires_grad_(True)
gt_img = (img.detach() + 0.15 * torch.randn(3, 37, 53, generator=gl)).clamp(0, 1)
ssim_v = LU.ssim(img, gt_img)
l1_v = LU.l1_loss(img, gt_img)
loss_v = 0.8 * l1_v + 0.2 * (1.0 - ssim_v)
(g_loss,) = torch.autograd.grad(loss_v, img)
out.update(loss_img=img.detach().numpy(), loss_gt=gt_img.numpy(), loss_ssim=np.float32(ssim_v.item()), loss_l1=np.float32(l1_v.item()),
           loss_total=np.float32(loss_v.item()), loss_grad=g_loss.numpy())

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_python.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: getattr(v, "shape", None) for k, v in out.items()})

"""Synthetic gaussians and cameras for benchmarks and examples.  Product-side helper: nothing
here touches the CPU oracle.  Formulae: /root/reference/utils/graphics_utils.py:38-71 (getWorld2View2,
getProjectionMatrix) and /root/reference/scene/cameras.py:80-89 (transposed matrices, camera centre)."""
from __future__ import annotations

import math

import numpy as np
import torch


def make_scene(P: int, seed: int = 0, sh_coeffs: int = 16, extent: float = 1.0, log_scale_mean: float = -4.0,
               log_scale_std: float = 0.5, dtype=torch.float32) -> dict:
    """xyz ~ U([-1,1]^3)*extent, log-scale ~ N(mu, std), rotation = normalised N(0,I), opacity = sigmoid(U(-2,4)),
    SH DC ~ N(0,0.5^2), rest ~ N(0,0.1^2).  Same generator sequence as the tests' scenes."""
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand(P, 3, generator=g) * 2 - 1) * extent
    scales = torch.exp(torch.randn(P, 3, generator=g) * log_scale_std + log_scale_mean)
    rot = torch.randn(P, 4, generator=g)
    rot = rot / rot.norm(dim=1, keepdim=True)
    opac = torch.sigmoid(torch.rand(P, 1, generator=g) * 6 - 2)
    shs = torch.randn(P, sh_coeffs, 3, generator=g) * 0.1
    shs[:, 0, :] = torch.randn(P, 3, generator=g) * 0.5
    return dict(means3D=xyz.to(dtype), scales=scales.to(dtype), rotations=rot.to(dtype), opacities=opac.to(dtype),
                shs=shs.to(dtype))


def look_at(eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)):
    """(R, T) in the convention of scene/cameras.py: R camera-to-world (columns = camera axes), T world->camera
    translation; the camera looks down +z."""
    eye = np.asarray(eye, dtype=np.float64)
    fwd = np.asarray(target, dtype=np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(np.asarray(up, dtype=np.float64), fwd)
    if np.linalg.norm(right) < 1e-8:
        right = np.cross(np.array([1.0, 0.0, 0.0]), fwd)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], axis=1)
    return R, -R.T @ eye


def camera_matrices(R, T, fovx: float, fovy: float, znear: float = 0.01, zfar: float = 100.0):
    """(world_view_transform, full_proj_transform, camera_center), transposed as the rasterizer consumes them."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = np.asarray(R).transpose()
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    wvt = torch.tensor(np.float32(np.linalg.inv(np.linalg.inv(Rt)))).transpose(0, 1)
    tx, ty = math.tan(fovx / 2), math.tan(fovy / 2)
    top, right = ty * znear, tx * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    proj = P.transpose(0, 1)
    full = wvt.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0)
    return wvt, full, wvt.inverse()[3, :3]


def sphere_pose(index: int, radius: float):
    """Camera `index` of a golden-angle spiral on a sphere, looking at the origin."""
    phi = index * 2.399963229728653
    y = 0.35 * math.sin(0.61803398875 * index * 2 * math.pi)
    r = math.sqrt(max(0.0, 1 - y * y))
    return look_at((radius * r * math.sin(phi), radius * y, -radius * r * math.cos(phi)))

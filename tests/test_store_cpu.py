"""CPU tests of the host logic of gaussian_store.GaussianModel: flat-store layout and aliasing, learning-rate schedule and
Adam step sizes against the reference fixture / torch.optim.Adam, and that nothing silently runs on the CPU."""
import math
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from gaussian_store import GaussianModel, expon_lr, store_offsets

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_model.npz")


def test_store_layout_matches_header_comment():
    P, M = 7, 16
    o = store_offsets(P, M)
    assert o == {"xyz": 0, "features": 21, "opacity": 21 + 48 * 7, "scaling": 21 + 48 * 7 + 7, "rotation": 21 + 48 * 7 + 28,
                 "total": 59 * 7}
    assert store_offsets(5, 1)["total"] == 14 * 5          # SH degree 0: 11 + 3 floats per gaussian


def test_views_alias_the_flat_buffers_and_grads():
    m = GaussianModel(3)
    m._allocate(5, torch.device("cpu"))                     # layout only; no kernel runs
    o = store_offsets(5, 16)
    m.store.copy_(torch.arange(m.store.numel(), dtype=torch.float32))
    assert m._xyz[2, 1].item() == 7.0
    assert m._features[1, 0, 2].item() == o["features"] + 48 + 2
    assert m._features_dc.shape == (5, 1, 3) and m._features_rest.shape == (5, 15, 3)
    assert m._features_rest[0, 0, 0].item() == o["features"] + 3
    assert m._opacity[4, 0].item() == o["opacity"] + 4
    assert m._scaling[1, 2].item() == o["scaling"] + 5
    assert m._rotation[4, 3].item() == o["total"] - 1
    for name, leaf in (("xyz", m.get_xyz), ("features", m.get_features), ("opacity", m.get_opacity), ("scaling", m.get_scaling),
                       ("rotation", m.get_rotation)):
        assert leaf.is_leaf and leaf.requires_grad and leaf.grad is not None and leaf.grad.is_contiguous(), name
        assert leaf.grad.shape == leaf.shape
        assert leaf.grad.data_ptr() == m.grad.data_ptr() + 4 * o[name], name
    # identity-activated groups read the store itself; the others read the activation buffer
    assert m.get_xyz.data_ptr() == m.store.data_ptr() and m.get_features.data_ptr() == m.store.data_ptr() + 4 * o["features"]
    assert m.get_opacity.data_ptr() == m.act.data_ptr() and m.get_rotation.data_ptr() == m.act.data_ptr() + 4 * 4 * 5


def test_lr_schedule_and_step_sizes_match_reference():
    gold = np.load(GOLD)
    opt = {k[4:]: float(gold[k]) for k in gold.files if k.startswith("opt_")}
    m = GaussianModel(3)
    m._allocate(3, torch.device("cpu"))
    m.spatial_lr_scale = float(gold["dens_extent"])
    m.training_setup(SimpleNamespace(**opt))
    for tag, it in (("s3", 3), ("s5", 5), ("s6", 6)):
        assert abs(m.update_learning_rate(it) - float(gold[tag + "_lr_xyz"])) < 1e-15
    assert m.lr["f_rest"] == opt["feature_lr"] / 20.0 and m.lr["opacity"] == opt["opacity_lr"]
    # torch.optim.Adam: step_size = lr / (1 - beta1^t), denom = sqrt(v) / sqrt(1 - beta2^t) + eps
    # ... with one step count PER GROUP (a parameter replaced by reset_opacity / densify misses the step that follows)
    m.group_steps = {n: 4 for n in m.group_steps}
    m.group_steps["opacity"] = 2
    ss, b2s = m.step_sizes()
    assert ss[3] == opt["opacity_lr"] / (1 - 0.9 ** 3) and b2s[3] == math.sqrt(1 - 0.999 ** 3)
    assert ss[4] == opt["scaling_lr"] / (1 - 0.9 ** 5) and b2s[4] == math.sqrt(1 - 0.999 ** 5)
    assert expon_lr(-1, 1e-3, 1e-5) == 0.0 and expon_lr(10, 0.0, 0.0) == 0.0
    assert abs(expon_lr(0, 1e-3, 1e-5, max_steps=100) - 1e-3) < 1e-18 and abs(expon_lr(100, 1e-3, 1e-5, max_steps=100) - 1e-5) < 1e-18


def test_no_cpu_fallback():
    m = GaussianModel(0)
    P = 4
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.create_from_tensors(torch.zeros(P, 3), torch.zeros(P, 1, 3), torch.zeros(P, 0, 3), torch.zeros(P, 3), torch.ones(P, 4),
                              torch.zeros(P, 1))
    m._allocate(P, torch.device("cpu"))
    m.lr = {n: 1e-3 for n in ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.optimizer_step()
    with pytest.raises(ValueError):
        GaussianModel(3).create_from_tensors(torch.zeros(P, 3), torch.zeros(P, 3), torch.zeros(P, 15, 3), torch.zeros(P, 3),
                                             torch.ones(P, 4), torch.zeros(P, 1))

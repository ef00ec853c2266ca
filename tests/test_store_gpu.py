"""GPU parity of the training state around the rasterizer (gaussian_store.GaussianModel: fused Adam step, densify/prune with
optimizer-state surgery, opacity reset) against (1) the fixture recorded from the reference's own scene/gaussian_model.py and
(2) oracle/model_oracle.py on larger seeded inputs."""
import math
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle.model_oracle import GROUPS, ModelOracle

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_model.npz")
ACT = ("xyz", "features", "opacity", "scaling", "rotation")
RAW_KEY = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling",
           "rotation": "_rotation"}
PARAM_TOL = 2e-6        # the store's parameters vs the reference's, absolute and relative
MOMENT_RTOL = 1e-4      # Adam moments (fma contraction and expf/sqrtf differ from the CPU by a few ulp)


def _model_cls():
    from gaussian_store import GaussianModel
    return GaussianModel


def _raw_of(m):
    """The store as the reference's six tensors (CPU)."""
    return {"xyz": m._xyz, "f_dc": m._features[:, :1], "f_rest": m._features[:, 1:], "opacity": m._opacity, "scaling": m._scaling,
            "rotation": m._rotation}


def _moments_of(m):
    from gaussian_store import store_offsets
    P, M, o = m.P, m.sh_coeffs, store_offsets(m.P, m.sh_coeffs)
    out = {}
    for kind, buf in (("m", m.exp_avg), ("v", m.exp_avg_sq)):
        feat = buf[o["features"]:o["opacity"]].view(P, M, 3)
        out[kind] = {"xyz": buf[:3 * P].view(P, 3), "f_dc": feat[:, :1], "f_rest": feat[:, 1:],
                     "opacity": buf[o["opacity"]:o["scaling"]].view(P, 1), "scaling": buf[o["scaling"]:o["rotation"]].view(P, 3),
                     "rotation": buf[o["rotation"]:].view(P, 4)}
    return out


def _opt_namespace(opt):
    return SimpleNamespace(**opt)


def _frac_bad(a, b, rtol, atol):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() > atol + rtol * b.abs()).double().mean())


def test_store_replays_reference_fixture():
    import model_replay as R
    R.replay(R.StoreDriver(R.load_gold(), _model_cls(), torch.device("cuda:0"), param_tol=PARAM_TOL, moment_rtol=MOMENT_RTOL))


def _random_state(P, seed, dev):
    g = torch.Generator().manual_seed(seed)
    raw = {"xyz": torch.randn(P, 3, generator=g), "f_dc": torch.randn(P, 1, 3, generator=g) * 0.5,
           "f_rest": torch.randn(P, 15, 3, generator=g) * 0.1, "opacity": torch.rand(P, 1, generator=g) * 9.0 - 6.5,
           "scaling": torch.randn(P, 3, generator=g) - 3.6, "rotation": torch.randn(P, 4, generator=g)}
    opt = dict(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=30000,
               feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005, rotation_lr=0.001, percent_dense=0.01)
    oracle = ModelOracle(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"], opt, 4.0)
    m = _model_cls()(3).create_from_tensors(*(raw[k].to(dev) for k in ("xyz", "f_dc", "f_rest", "scaling", "rotation", "opacity")), 4.0)
    m.training_setup(_opt_namespace(opt))
    return raw, opt, oracle, m, g


@pytest.mark.parametrize("P", [1, 33, 20011])
def test_adam_step_matches_torch_adam_with_autograd_chain(P):
    dev = torch.device("cuda:0")
    raw, opt, oracle, m, g = _random_state(P, 5 + P, dev)
    shapes = {"xyz": (P, 3), "features": (P, 16, 3), "opacity": (P, 1), "scaling": (P, 3), "rotation": (P, 4)}
    for it in range(1, 5):
        scale = 1e-3 if it < 3 else 1.0          # the moments keep the large step's rounding error afterwards
        grads = {n: torch.randn(*shapes[n], generator=g) * (1e-3 if it != 3 else 1.0) for n in ACT}
        oracle.step(it, grads)
        m.update_learning_rate(it)
        for n, leaf in (("xyz", m.get_xyz), ("features", m.get_features), ("opacity", m.get_opacity), ("scaling", m.get_scaling),
                        ("rotation", m.get_rotation)):
            leaf.grad.copy_(grads[n].to(dev))
        m.optimizer_step()
        rawm, mom, omom = _raw_of(m), _moments_of(m), oracle.moments()
        for n in GROUPS:
            # a raw gradient that cancels to ~0 may flip the sign of m/sqrt(v): allow a 1e-4 fraction of outliers
            assert _frac_bad(rawm[n], oracle.p[n].detach(), 5e-6, 5e-6) <= 1e-4, (it, n)
            # normalize's backward (g - y (y.g)) / |q| cancels: its absolute error scales with |g|, not with the result
            assert _frac_bad(mom["m"][n], omom[n]["m"], 1e-4, 1e-9 + 2e-7 * scale) <= 1e-4, (it, n)
            assert _frac_bad(mom["v"][n], omom[n]["v"], 1e-4, 1e-14 + 2e-9 * scale * scale) <= 1e-4, (it, n)
        act = oracle.activated()
        for n, leaf in (("opacity", m.get_opacity), ("scaling", m.get_scaling), ("rotation", m.get_rotation)):
            assert _frac_bad(leaf.detach(), act[n].detach(), 1e-5, 1e-6) <= 1e-4, (it, n, "activated")


def test_adam_step_visible_mask_leaves_other_rows_untouched():
    dev = torch.device("cuda:0")
    P = 4099
    raw, opt, oracle, m, g = _random_state(P, 17, dev)
    m.update_learning_rate(1)
    m.grad.copy_(torch.randn(m.grad.shape, generator=g).to(dev) * 1e-3)
    before = (m.store.clone(), m.exp_avg.clone(), m.exp_avg_sq.clone(), m.act.clone())
    vis = (torch.rand(P, generator=g) < 0.4).to(dev)
    m.optimizer_step(visible=vis)
    from gaussian_store import store_offsets
    o, M = store_offsets(P, 16), 16
    rows = {"xyz": (0, 3), "features": (o["features"], 3 * M), "opacity": (o["opacity"], 1), "scaling": (o["scaling"], 3),
            "rotation": (o["rotation"], 4)}
    for buf, old in zip((m.store, m.exp_avg, m.exp_avg_sq), before[:3]):
        for name, (a, w) in rows.items():
            new_v, old_v = buf[a:a + w * P].view(P, w), old[a:a + w * P].view(P, w)
            assert torch.equal(new_v[~vis], old_v[~vis]), name
            assert not torch.equal(new_v[vis], old_v[vis]), name
    assert torch.equal(m.act[:P][~vis], before[3][:P][~vis])


@pytest.mark.parametrize("max_screen", [20, None])
def test_densify_and_prune_matches_oracle(max_screen):
    dev = torch.device("cuda:0")
    P = 20011
    raw, opt, oracle, m, g = _random_state(P, 23, dev)
    # two optimizer steps first so that the moments are non-trivial
    shapes = {"xyz": (P, 3), "features": (P, 16, 3), "opacity": (P, 1), "scaling": (P, 3), "rotation": (P, 4)}
    for it in (1, 2):
        grads = {n: torch.randn(*shapes[n], generator=g) * 1e-3 for n in ACT}
        oracle.step(it, grads)
        m.update_learning_rate(it)
        for n, leaf in (("xyz", m.get_xyz), ("features", m.get_features), ("opacity", m.get_opacity), ("scaling", m.get_scaling),
                        ("rotation", m.get_rotation)):
            leaf.grad.copy_(grads[n].to(dev))
        m.optimizer_step()
    # from here on both sides start from the SAME state, so that classification ties cannot come from the Adam steps
    with torch.no_grad():
        for n in GROUPS:
            oracle.p[n].copy_(_raw_of(m)[n].cpu())
    denom = torch.randint(0, 4, (P, 1), generator=g).float()
    accum = torch.rand(P, 1, generator=g) * 0.0008 * denom
    oracle.grad_accum, oracle.denom = accum.clone(), denom.clone()
    m.xyz_gradient_accum, m.denom = accum.to(dev), denom.to(dev)
    drawn = {}

    def draw(rows):
        drawn["u"] = torch.randn(rows, 3, generator=g)
        return drawn["u"]

    info = m.densify_and_prune(0.0002, 0.005, 4.0, max_screen, unit_samples=draw)
    ref = oracle.densify_and_prune(0.0002, 0.005, 4.0, max_screen, unit_samples=drawn["u"])
    assert info == ref, (info, ref)
    assert info["n_clone"] > 100 and info["n_split"] > 100 and info["n_pruned"] > 100
    rawm, mom, omom = _raw_of(m), _moments_of(m), oracle.moments()
    for n in GROUPS:
        torch.testing.assert_close(rawm[n].cpu(), oracle.p[n].detach(), rtol=2e-6, atol=2e-6, msg=lambda s: f"{n}: {s}")
        for kind in ("m", "v"):
            torch.testing.assert_close(mom[kind][n].cpu(), omom[n][kind], rtol=MOMENT_RTOL, atol=1e-9 if kind == "m" else 1e-13,
                                       msg=lambda s: f"{kind} {n}: {s}")
    # the activated tensors follow the new store
    act = oracle.activated()
    torch.testing.assert_close(m.get_scaling.detach().cpu(), act["scaling"].detach(), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(m.get_rotation.detach().cpu(), act["rotation"].detach(), rtol=1e-5, atol=1e-6)
    assert m.grad.numel() == m.store.numel() and m.get_xyz.grad.data_ptr() == m.grad.data_ptr()


def test_store_trains_through_the_rasterizer():
    """render_views_backward accumulates straight into the store's gradient buffer; a few fused optimizer steps lower the loss."""
    import diff_gaussian_rasterization as dgr
    from gaussian_renderer import render_views_backward
    from gaussian_renderer.synthetic import make_scene, sphere_pose, camera_matrices
    dev = torch.device("cuda:0")
    P, H, W = 3000, 96, 128
    sc = {k: v.to(dev) for k, v in make_scene(P, seed=3, log_scale_mean=-3.0).items()}
    inv_sig = lambda y: torch.log(y / (1 - y))
    m = _model_cls()(3)
    m.active_sh_degree = 3
    m.create_from_tensors(sc["means3D"], sc["shs"][:, :1].contiguous(), sc["shs"][:, 1:].contiguous(), torch.log(sc["scales"]),
                          sc["rotations"], inv_sig(sc["opacities"].clamp(1e-4, 1 - 1e-4)).reshape(P, 1), 1.0)
    m.training_setup(SimpleNamespace(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
                                     position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005,
                                     rotation_lr=0.001, percent_dense=0.01))
    cams = []
    for k in range(2):
        R, T = sphere_pose(k, 4.0)
        wvt, full, center = camera_matrices(R, T, 1.0, 0.8)
        cams.append(SimpleNamespace(image_height=H, image_width=W, FoVx=1.0, FoVy=0.8, world_view_transform=wvt.to(dev),
                                    full_proj_transform=full.to(dev), camera_center=center.to(dev)))
    pipe = SimpleNamespace(convert_SHs_python=False, compute_cov3D_python=False, debug=False, antialiasing=False)
    bg = torch.zeros(3, device=dev)
    gts = [torch.full((3, H, W), 0.5, device=dev) for _ in cams]
    loss_fn = lambda img, dep, vi: dgr.l1_loss_and_grad(img, gts[vi])
    losses = []
    for it in range(1, 16):
        m.update_learning_rate(it)
        out = render_views_backward(cams, m, pipe, bg, loss_fn, loss_returns_grad=True, overwrite=True)
        assert float(m.grad.abs().sum()) > 0.0
        m.optimizer_step()
        losses.append(float(out["losses"].sum()))
    assert all(math.isfinite(v) for v in losses) and losses[-1] < losses[0] * 0.98, losses


def test_model_ply_round_trip(tmp_path):
    dev = torch.device("cuda:0")
    raw, opt, oracle, m, g = _random_state(257, 41, dev)
    path = str(tmp_path / "point_cloud.ply")
    m.save_ply(path)
    m2 = _model_cls()(3).load_ply(path, device=dev, spatial_lr_scale=4.0)
    assert m2.P == m.P and m2.active_sh_degree == 3 and torch.equal(m2.store, m.store)
    torch.testing.assert_close(m2.act, m.act, rtol=0, atol=0)

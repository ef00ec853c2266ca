"""Shared driver of the reference training-state fixture (tests/golden/reference_model.npz, recorded by running the
reference's own scene/gaussian_model.py: tests/golden/make_golden_model.py).  Three replays use it: the CPU oracle
(test_model_oracle.py), the store on the host build of the kernel sources (test_kernel_source_on_host.py) and the store on
the GPU (test_store_gpu.py).

The sequence: 3 steps | densify (size threshold) | 2 steps | opacity reset, 1 step | densify (no size threshold) | and then
the order train.py itself uses (train.py:139-190) -- backward, THEN densify / reset, THEN optimizer.step(), which skips the
parameters that were just replaced:  backward, densify, step (a no-op) | 1 step | backward, opacity reset, step (five
groups move) | 1 step (the opacity group's own step count now lags)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_model.npz")
ACT = ("xyz", "features", "opacity", "scaling", "rotation")
GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")
RAW_KEY = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling",
           "rotation": "_rotation"}
DENS_P = {"dens": "d_P", "dens2": "d2_P", "dens3": "t1_P"}


def load_gold():
    return {k: v for k, v in np.load(GOLD).items()}


def opt_of(gold):
    return {k[4:]: float(v) for k, v in gold.items() if k.startswith("opt_")}


def act_grads(gold, act, device="cpu"):
    """dLoss/d(activated) = w + u * act for the fixture's loss, rows [0, P)."""
    P = act["xyz"].shape[0]
    return {n: torch.from_numpy(gold["w_" + n][:P]).to(device) + torch.from_numpy(gold["u_" + n][:P]).to(device) * act[n].detach()
            for n in ACT}


def replay(drv):
    """drv: backward(it), opt_step(), densify(key, seed, max_screen), reset(), check(tag), check_steps(tag)."""
    it = 0

    def step():
        nonlocal it
        it += 1
        drv.backward(it)
        drv.opt_step()

    for _ in range(3):
        step()
    drv.check("s3")
    drv.densify("dens", 77, 20)
    drv.check("d")
    for _ in range(2):
        step()
    drv.check("s5")
    drv.reset()
    step()
    drv.check("s6")
    drv.densify("dens2", 78, None)
    drv.check("d2")
    # ---- train.py order ----
    it += 1
    drv.backward(it)
    drv.densify("dens3", 79, 20)
    drv.opt_step()
    drv.check("t1")
    drv.check_steps("t1")
    step()
    drv.check("t2")
    drv.check_steps("t2")
    it += 1
    drv.backward(it)
    drv.reset()
    drv.opt_step()
    drv.check("t3")
    drv.check_steps("t3")
    step()
    drv.check("t4")
    drv.check_steps("t4")


class StoreDriver:
    """gaussian_store.GaussianModel (on whatever device `dev` names) against the fixture."""

    def __init__(self, gold, model_cls, dev, param_tol=2e-6, moment_rtol=1e-4):
        from types import SimpleNamespace
        self.gold, self.dev, self.param_tol, self.moment_rtol = gold, dev, param_tol, moment_rtol
        self.opt = opt_of(gold)
        self.extent = float(gold["dens_extent"])
        t = lambda k: torch.from_numpy(gold["init" + RAW_KEY[k]]).to(dev)
        self.m = model_cls(3).create_from_tensors(t("xyz"), t("f_dc"), t("f_rest"), t("scaling"), t("rotation"), t("opacity"), self.extent)
        self.m.training_setup(SimpleNamespace(**self.opt))

    def backward(self, it):
        m, gold = self.m, self.gold
        lr = m.update_learning_rate(it)
        if f"s{it}_lr_xyz" in gold:
            assert abs(lr - float(gold[f"s{it}_lr_xyz"])) < 1e-15
        act = {"xyz": m.get_xyz, "features": m.get_features, "opacity": m.get_opacity, "scaling": m.get_scaling, "rotation": m.get_rotation}
        for n, g in act_grads(gold, act, self.dev).items():
            act[n].grad.copy_(g.view_as(act[n]))
        m.gradients_ready()

    def opt_step(self):
        self.m.optimizer_step()

    def reset(self):
        self.m.reset_opacity()

    def densify(self, key, seed, max_screen):
        m, gold = self.m, self.gold
        m.xyz_gradient_accum = torch.from_numpy(gold[key + "_accum"]).to(self.dev)
        m.denom = torch.from_numpy(gold[key + "_denom"]).to(self.dev)

        def draw(rows):                       # what the reference's torch.normal consumed from the CPU generator
            torch.manual_seed(seed)
            return torch.randn(rows, 3)

        info = m.densify_and_prune(self.opt["densify_grad_threshold"], 0.005, self.extent, max_screen, unit_samples=draw)
        assert info["P"] == int(gold[DENS_P[key]]), info
        assert info["n_clone"] > 0 and info["n_split"] > 0, info
        assert float(m.max_radii2D.abs().max()) == 0.0 and float(m.xyz_gradient_accum.abs().max()) == 0.0

    def check(self, tag):
        from gaussian_store import store_offsets
        m, gold = self.m, self.gold
        raw = {"xyz": m._xyz, "f_dc": m._features[:, :1], "f_rest": m._features[:, 1:], "opacity": m._opacity, "scaling": m._scaling,
               "rotation": m._rotation}
        for n in GROUPS:
            ref = torch.from_numpy(gold[tag + RAW_KEY[n]])
            assert tuple(raw[n].shape) == tuple(ref.shape), (tag, n, raw[n].shape, ref.shape)
            torch.testing.assert_close(raw[n].cpu(), ref, rtol=self.param_tol, atol=self.param_tol, msg=lambda s: f"{tag} {n}: {s}")
        P, M, o = m.P, m.sh_coeffs, store_offsets(m.P, m.sh_coeffs)
        for kind, buf in (("m", m.exp_avg), ("v", m.exp_avg_sq)):
            feat = buf[o["features"]:o["opacity"]].view(P, M, 3)
            got = {"xyz": buf[:3 * P].view(P, 3), "f_dc": feat[:, :1], "f_rest": feat[:, 1:], "opacity": buf[o["opacity"]:o["scaling"]].view(P, 1),
                   "scaling": buf[o["scaling"]:o["rotation"]].view(P, 3), "rotation": buf[o["rotation"]:].view(P, 4)}
            for n in GROUPS:
                torch.testing.assert_close(got[n].cpu(), torch.from_numpy(gold[f"{tag}_{kind}_{n}"]), rtol=self.moment_rtol,
                                           atol=1e-9 if kind == "m" else 1e-13, msg=lambda s: f"{tag} {kind} {n}: {s}")

    def check_steps(self, tag):
        for n in GROUPS:
            assert self.m.group_steps[n] == int(self.gold[f"{tag}_step_{n}"]), (tag, n, self.m.group_steps)

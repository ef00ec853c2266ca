"""Pins oracle/model_oracle.py against the fixture produced by the reference's own scene/gaussian_model.py + torch.optim.Adam
(tests/golden/make_golden_model.py); the sequence is tests/model_replay.py's, including train.py's own order (backward, then
densify / opacity reset, then an optimizer.step() that skips the replaced parameters)."""
import os

import numpy as np
import pytest
import torch

from oracle.model_oracle import GROUPS, ModelOracle, expon_lr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_model.npz")
ACT = ("xyz", "features", "opacity", "scaling", "rotation")
RAW_KEY = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling",
           "rotation": "_rotation"}


@pytest.fixture(scope="module")
def gold():
    return {k: v for k, v in np.load(GOLD).items()}


def opt_of(gold):
    return {k[4:]: float(v) for k, v in gold.items() if k.startswith("opt_")}


def make_oracle(gold):
    t = lambda k: torch.from_numpy(gold["init" + RAW_KEY[k]])
    return ModelOracle(t("xyz"), t("f_dc"), t("f_rest"), t("opacity"), t("scaling"), t("rotation"), opt_of(gold),
                       float(gold["dens_extent"]))


def act_grads(gold, act):
    P = act["xyz"].shape[0]
    return {n: torch.from_numpy(gold["w_" + n][:P]) + torch.from_numpy(gold["u_" + n][:P]) * act[n].detach() for n in ACT}


def check(m: ModelOracle, gold, tag, tol=2e-6):
    mom = m.moments()
    for n in GROUPS:
        ref = torch.from_numpy(gold[tag + RAW_KEY[n]])
        assert tuple(m.p[n].shape) == tuple(ref.shape), (tag, n, m.p[n].shape, ref.shape)
        torch.testing.assert_close(m.p[n].detach(), ref, rtol=tol, atol=tol, msg=lambda s: f"{tag} {n}: {s}")
        for kind in ("m", "v"):
            torch.testing.assert_close(mom[n][kind], torch.from_numpy(gold[f"{tag}_{kind}_{n}"]), rtol=1e-5, atol=1e-12,
                                       msg=lambda s: f"{tag} {kind} {n}: {s}")


def test_expon_lr_matches_reference_schedule(gold):
    o = opt_of(gold)
    scale = float(gold["dens_extent"])
    for tag, it in (("s3", 3), ("s5", 5), ("s6", 6)):
        lr = expon_lr(it, o["position_lr_init"] * scale, o["position_lr_final"] * scale, o["position_lr_delay_mult"],
                      int(o["position_lr_max_steps"]))
        assert abs(lr - float(gold[tag + "_lr_xyz"])) < 1e-15


def test_oracle_replays_reference_training_state(gold):
    import model_replay as R
    o = opt_of(gold)
    m = make_oracle(gold)

    class Drv:
        def backward(self, it):
            m.backward(it, act_grads(gold, m.activated()))

        def opt_step(self):
            m.optimizer_step()

        def reset(self):
            m.reset_opacity()

        def densify(self, key, seed, max_screen):
            m.grad_accum = torch.from_numpy(gold[key + "_accum"]).clone()
            m.denom = torch.from_numpy(gold[key + "_denom"]).clone()
            if key == "dens":
                m.max_radii2D = torch.from_numpy(gold["dens_max_radii2D"]).clone()
            torch.manual_seed(seed)
            info = m.densify_and_prune(o["densify_grad_threshold"], 0.005, float(gold["dens_extent"]), max_screen)
            assert info["n_clone"] > 0 and info["n_split"] > 0, info
            assert info["n_pruned"] > 0 or key != "dens", info   # first pass exercises every branch
            assert info["P"] == int(gold[R.DENS_P[key]])
            assert float(m.max_radii2D.abs().max()) == 0.0 and float(m.grad_accum.abs().max()) == 0.0

        def check(self, tag):
            check(m, gold, tag)

        def check_steps(self, tag):
            assert {n: int(v) for n, v in m.steps().items()} == {n: int(gold[f"{tag}_step_{n}"]) for n in GROUPS}

    R.replay(Drv())

"""CPU oracle for the training state around the rasterizer: per-group Adam with the
reference's activations chained by autograd, the exponential position learning rate, densify (clone + split) and prune
with optimizer-state surgery, and the opacity reset.

TEST INFRASTRUCTURE ONLY -- nothing under gaussian-splatting_b200/ may import this module.  It restates, in its own
words, what /root/reference/scene/gaussian_model.py does on these lines:

  activations                         :33-46, 102-130   exp / sigmoid / normalize, features = cat(dc, rest)
  training_setup (6 Adam groups)      :176-199          lr = (pos_lr*spatial_scale, f, f/20, opacity, scaling, rotation), eps 1e-15
  update_learning_rate                :213-223          only the xyz group is scheduled (utils/general_utils.py:29-62)
  reset_opacity                       :258-261, 302-314 raw = logit(min(sigmoid(raw), 0.01)), moments of the group zeroed
  prune / cat optimizer surgery       :316-397          survivors keep exp_avg / exp_avg_sq, new rows start at zero
  densify_and_split / _clone / _prune :399-469

and it is PINNED: tests/test_model_oracle.py replays tests/golden/reference_model.npz, which was produced by running that
very file (unmodified) on the CPU (tests/golden/make_golden_model.py).

Everything is float32 torch on the CPU; Adam is torch.optim.Adam itself, as in the reference."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


def expon_lr(step: int, lr_init: float, lr_final: float, delay_mult: float, max_steps: int, delay_steps: int = 0) -> float:
    """Log-linear decay from lr_init to lr_final over max_steps (general_utils.py:47-60)."""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    rate = 1.0
    if delay_steps > 0:
        rate = delay_mult + (1.0 - delay_mult) * math.sin(0.5 * math.pi * min(max(step / delay_steps, 0.0), 1.0))
    t = min(max(step / max_steps, 0.0), 1.0)
    return rate * math.exp(math.log(lr_init) * (1.0 - t) + math.log(lr_final) * t)


def quat_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """Rotation matrices of (unnormalised) quaternions, w first (general_utils.py:78-99)."""
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).reshape(-1, 3, 3)


class ModelOracle:
    def __init__(self, xyz, f_dc, f_rest, opacity, scaling, rotation, opt: Dict[str, float], spatial_lr_scale: float):
        mk = lambda t: torch.nn.Parameter(t.detach().clone().float())
        self.p = {"xyz": mk(xyz), "f_dc": mk(f_dc), "f_rest": mk(f_rest), "opacity": mk(opacity), "scaling": mk(scaling),
                  "rotation": mk(rotation)}
        self.opt, self.scale = dict(opt), float(spatial_lr_scale)
        lrs = {"xyz": opt["position_lr_init"] * self.scale, "f_dc": opt["feature_lr"], "f_rest": opt["feature_lr"] / 20.0,
               "opacity": opt["opacity_lr"], "scaling": opt["scaling_lr"], "rotation": opt["rotation_lr"]}
        self.adam = torch.optim.Adam([{"params": [self.p[n]], "lr": lrs[n], "name": n} for n in GROUPS], lr=0.0, eps=1e-15)
        P = self.P
        self.grad_accum, self.denom, self.max_radii2D = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P)

    @property
    def P(self) -> int:
        return int(self.p["xyz"].shape[0])

    def activated(self) -> Dict[str, torch.Tensor]:
        return {"xyz": self.p["xyz"], "features": torch.cat((self.p["f_dc"], self.p["f_rest"]), dim=1),
                "opacity": torch.sigmoid(self.p["opacity"]), "scaling": torch.exp(self.p["scaling"]),
                "rotation": torch.nn.functional.normalize(self.p["rotation"])}

    def lr_xyz(self, iteration: int) -> float:
        o = self.opt
        return expon_lr(iteration, o["position_lr_init"] * self.scale, o["position_lr_final"] * self.scale,
                        o["position_lr_delay_mult"], int(o["position_lr_max_steps"]))

    def backward(self, iteration: int, act_grads: Dict[str, torch.Tensor]) -> None:
        """``loss.backward()`` of train.py:144 given dLoss/d(activated tensors): autograd chains them to the raw parameters."""
        for grp in self.adam.param_groups:
            if grp["name"] == "xyz":
                grp["lr"] = self.lr_xyz(iteration)
        act = self.activated()
        names = list(act_grads)
        torch.autograd.backward([act[n] for n in names], [act_grads[n].reshape(act[n].shape) for n in names])

    def optimizer_step(self) -> None:
        """train.py:178-186.  A parameter that densify_and_prune / reset_opacity replaced since the backward pass has
        ``grad is None`` and is skipped by torch.optim.Adam: no update, no moment decay, no step-count increment."""
        self.adam.step()
        self.adam.zero_grad(set_to_none=True)

    def step(self, iteration: int, act_grads: Dict[str, torch.Tensor]) -> None:
        self.backward(iteration, act_grads)
        self.optimizer_step()

    def steps(self) -> Dict[str, float]:
        out = {}
        for grp in self.adam.param_groups:
            st = self.adam.state.get(grp["params"][0])
            out[grp["name"]] = float(st["step"]) if st is not None and "step" in st else 0.0
        return out

    # --- optimizer-state surgery -----------------------------------------------------------------------------------
    def _rebuild(self, fn_param, fn_moment, only=None) -> None:
        """Replaces the parameter (a fresh nn.Parameter: grad None) and moments of every group, or of the group ``only``."""
        for grp in self.adam.param_groups:
            if only is not None and grp["name"] != only:
                continue
            old = grp["params"][0]
            st = self.adam.state.pop(old, None)
            new = torch.nn.Parameter(fn_param(grp["name"], old.detach()))
            if st is not None:
                st["exp_avg"], st["exp_avg_sq"] = fn_moment(grp["name"], st["exp_avg"]), fn_moment(grp["name"], st["exp_avg_sq"])
                self.adam.state[new] = st
            grp["params"][0] = new
            self.p[grp["name"]] = new

    def _append(self, rows: Dict[str, torch.Tensor]) -> None:
        self._rebuild(lambda n, t: torch.cat((t, rows[n]), 0), lambda n, t: torch.cat((t, torch.zeros_like(rows[n])), 0))
        P = self.P
        self.grad_accum, self.denom, self.max_radii2D = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P)

    def prune(self, drop: torch.Tensor) -> None:
        keep = ~drop
        self._rebuild(lambda n, t: t[keep], lambda n, t: t[keep])
        self.grad_accum, self.denom, self.max_radii2D = self.grad_accum[keep], self.denom[keep], self.max_radii2D[keep]

    def reset_opacity(self) -> None:
        a = torch.minimum(torch.sigmoid(self.p["opacity"].detach()), torch.tensor(0.01))
        new = torch.log(a / (1.0 - a))
        self._rebuild(lambda n, t: new, lambda n, t: torch.zeros_like(t), only="opacity")   # replace_tensor_to_optimizer :302-314

    # --- densification -----------------------------------------------------------------------------------------------
    def densify_and_prune(self, max_grad: float, min_opacity: float, extent: float, max_screen_size: Optional[float],
                          unit_samples: Optional[torch.Tensor] = None, n_children: int = 2) -> Dict[str, int]:
        """``unit_samples`` [n_children * n_split, 3] standard normals (drawn from the global generator when None, which is
        what the reference's torch.normal(mean=0, std=stds) consumes)."""
        g = self.grad_accum / self.denom
        g[g.isnan()] = 0.0
        limit = self.opt["percent_dense"] * extent
        raw = {n: self.p[n].detach() for n in GROUPS}
        # clone: small gaussians with a large view-space gradient are duplicated in place
        sel = (g.norm(dim=-1) >= max_grad) & (torch.exp(raw["scaling"]).max(dim=1).values <= limit)
        n_clone = int(sel.sum())
        self._append({n: raw[n][sel] for n in GROUPS})
        # split: large ones are replaced by n_children samples of themselves, 1.6x smaller
        raw = {n: self.p[n].detach() for n in GROUPS}
        gp = torch.zeros(self.P)
        gp[:g.shape[0]] = g.squeeze(-1)
        sel = (gp >= max_grad) & (torch.exp(raw["scaling"]).max(dim=1).values > limit)
        n_split = int(sel.sum())
        rep = lambda t: t[sel].repeat(n_children, *([1] * (t.dim() - 1)))
        std = rep(torch.exp(raw["scaling"]))
        unit = torch.randn(std.shape) if unit_samples is None else unit_samples
        offs = torch.bmm(quat_to_matrix(rep(raw["rotation"])), (unit * std).unsqueeze(-1)).squeeze(-1)
        rows = {n: rep(raw[n]) for n in GROUPS}
        rows["xyz"] = offs + rows["xyz"]
        rows["scaling"] = torch.log(std / (0.8 * n_children))
        self._append(rows)
        self.prune(torch.cat((sel, torch.zeros(n_children * n_split, dtype=torch.bool))))
        # prune: transparent, or (once a screen-size threshold is in force) too large in world space.  max_radii2D was just
        # zeroed by the appends above, so the screen-space test of gaussian_model.py:462 never fires -- kept as is.
        drop = (torch.sigmoid(self.p["opacity"].detach()) < min_opacity).squeeze(-1)
        if max_screen_size:
            drop = drop | (self.max_radii2D > max_screen_size) | (torch.exp(self.p["scaling"].detach()).max(dim=1).values > 0.1 * extent)
        n_pruned = int(drop.sum())
        self.prune(drop)
        return {"n_clone": n_clone, "n_split": n_split, "n_pruned": n_pruned, "P": self.P}

    def moments(self) -> Dict[str, Dict[str, torch.Tensor]]:
        out = {}
        for grp in self.adam.param_groups:
            st = self.adam.state.get(grp["params"][0])
            if st is not None:
                out[grp["name"]] = {"m": st["exp_avg"], "v": st["exp_avg_sq"]}
        return out

"""Drop-in ``fused_ssim`` package (the reference imports it at /root/reference/train.py:31-35 and calls
``fused_ssim(image.unsqueeze(0), gt_image.unsqueeze(0))`` at :122; its own submodule is absent from /root/reference).

Backed by libgs_b200.so's fused photometric kernels (csrc/loss.cu) with lambda = 1, no L1 term and NO clamp of the input,
i.e. the plain mean SSIM of utils/loss_utils.py:56-86 (11x11 gaussian window, sigma 1.5, zero padding) and its gradient
w.r.t. the first image, for any input range."""
from __future__ import annotations

import torch

import diff_gaussian_rasterization as _dgr

__all__ = ["fused_ssim"]


class _FusedSSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2):
        x = img1.reshape(-1, *img1.shape[-2:])            # [N*C, H, W]: every plane is an independent channel
        y = img2.reshape(-1, *img2.shape[-2:])
        loss, grad, parts = _dgr.photometric_loss_and_grad(x, y, lambda_dssim=1.0, clamp_input=False)   # loss = 1 - SSIM
        ctx.save_for_backward(grad)
        ctx.shape = img1.shape
        return parts[2].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (-grad * g).reshape(ctx.shape), None        # d SSIM / d img1 = - d(1 - SSIM) / d img1


def fused_ssim(img1: torch.Tensor, img2: torch.Tensor, padding: str = "same", train: bool = True) -> torch.Tensor:
    if padding != "same":
        raise NotImplementedError("fused_ssim (B200): only the reference's zero 'same' padding is implemented")
    return _FusedSSIM.apply(img1, img2)

"""Fused photometric loss (csrc/loss.cu) vs the reference formulation in torch ops, 3x1080x1920 (GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
import torch
import diff_gaussian_rasterization as dgr
dev = torch.device("cuda", 0)
img = torch.rand(3, 1080, 1920, device=dev); gt = torch.rand(3, 1080, 1920, device=dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
import torch.nn.functional as F
_g = torch.exp(-((torch.arange(11, dtype=torch.float32) - 5) ** 2) / (2 * 1.5 ** 2)); _g = _g / _g.sum()
_win = (_g[:, None] @ _g[None, :])[None, None].expand(3, 1, 11, 11).contiguous().to(dev)


def ssim_torch(a, b):      # the formulation of utils/loss_utils.py:56-86: five depthwise 11x11 convolutions + elementwise ops
    mu1, mu2 = F.conv2d(a[None], _win, padding=5, groups=3), F.conv2d(b[None], _win, padding=5, groups=3)
    s1 = F.conv2d(a[None] * a[None], _win, padding=5, groups=3) - mu1 * mu1
    s2 = F.conv2d(b[None] * b[None], _win, padding=5, groups=3) - mu2 * mu2
    s12 = F.conv2d(a[None] * b[None], _win, padding=5, groups=3) - mu1 * mu2
    return (((2 * mu1 * mu2 + 0.01 ** 2) * (2 * s12 + 0.03 ** 2)) / ((mu1 * mu1 + mu2 * mu2 + 0.01 ** 2) * (s1 + s2 + 0.03 ** 2))).mean()


def torch_version():
    x = img.clone().requires_grad_(True)
    xc = x.clamp(0, 1)
    (0.8 * (xc - gt).abs().mean() + 0.2 * (1.0 - ssim_torch(xc, gt))).backward()
print("torch ops (fwd+bwd): %.3f ms" % t(torch_version))
print("fused L1+SSIM+grad : %.3f ms" % t(lambda: dgr.photometric_loss_and_grad(img, gt, 0.2)))
print("fused L1+grad      : %.3f ms" % t(lambda: dgr.l1_loss_and_grad(img, gt)))
dgr.set_option("time_kernels", 2); dgr.kernel_time("", reset=True)
for _ in range(10): dgr.photometric_loss_and_grad(img, gt, 0.2)
torch.cuda.synchronize()
print("ssim_maps %.3f ms, ssim_grad %.3f ms" % (dgr.kernel_time("ssim_maps")[0] / 10, dgr.kernel_time("ssim_grad")[0] / 10))

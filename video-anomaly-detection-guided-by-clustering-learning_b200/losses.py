"""Pixel losses of the training drivers as single-pass fused reductions
 with autograd.

* ``Recon_Loss``           — loss_tool/Recon_Loss.py:11-32 (mean L1, pads target on D)
* ``mse_mean(r, t)``       — ``torch.mean(nn.MSELoss(reduction='none')(r, t))``  main.py:191
* ``e4_norm(r, t)``        — ``torch.norm(nn.MSELoss(reduction='none')(r, t))``  main_predict.py:273-275
"""
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, f32c, ptr, stream, workspace, LOSS_L1_MEAN, LOSS_MSE_MEAN, LOSS_E4_NORM


class _PixelLoss(torch.autograd.Function):
    """out = [loss, raw_sum]; differentiable w.r.t. x (the reconstruction) only,
    like the reference's use (the target is data)."""

    @staticmethod
    def forward(ctx, x, t, mode):
        _lib.require_cuda(x, t)
        xc, tc = f32c(x), f32c(t)
        n = tc.numel()
        out = torch.empty((2,), device=xc.device, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_pixel_loss_workspace_bytes(n), xc.device)
        check(l.vadc_pixel_loss(ptr(xc), ptr(tc), n, None, 0, mode, ptr(out), ptr(ws), ws.numel(), stream()),
              "vadc_pixel_loss")
        ctx.save_for_backward(xc, tc, out)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, gout):
        xc, tc, out = ctx.saved_tensors
        n = xc.numel()
        gout = f32c(gout)
        # fold the gradient of the raw-sum output (out[1]) into d objective / d loss value (out[0])
        if ctx.mode == LOSS_E4_NORM:
            g = gout[0:1] + gout[1:2] * (2.0 * out[0:1])       # sum = loss^2
        else:
            g = gout[0:1] + gout[1:2] * float(n)               # sum = loss * n
        g = g.contiguous()
        gx = torch.empty_like(xc)
        check(_lib.lib().vadc_pixel_loss_bwd(ptr(xc), ptr(tc), n, ctx.mode, ptr(g), ptr(out), n, ptr(gx), stream()),
              "vadc_pixel_loss_bwd")
        return gx, None, None


def _pixel_loss(x, t, mode):
    if x.shape != t.shape:
        raise RuntimeError(f"The size of tensor a {tuple(x.shape)} must match the size of tensor b {tuple(t.shape)}")
    return _PixelLoss.apply(x, t, mode)


def l1_mean(x, t):
    return _pixel_loss(x, t, LOSS_L1_MEAN)[0]


def mse_mean(x, t):
    """main.py:191"""
    return _pixel_loss(x, t, LOSS_MSE_MEAN)[0]


def e4_norm(x, t):
    """main_predict.py:273-275: ||(x - t)^2||_F = sqrt(sum (x-t)^4)"""
    return _pixel_loss(x, t, LOSS_E4_NORM)[0]


def e4_sum(x, t):
    """raw sum (x-t)^4 — all-reduce this across ranks before the sqrt for
    full-batch semantics under data parallelism."""
    return _pixel_loss(x, t, LOSS_E4_NORM)[1]


class Recon_Loss(nn.Module):
    """Drop-in for loss_tool/Recon_Loss.py:11-32: two videos [B,C,D,H,W]; the
    target is zero-padded on D up to a multiple of patch_size[0]; mean L1."""

    def __init__(self, patch_size):
        super().__init__()
        self.patch_size = patch_size

    def forward(self, x, target):
        _, _, D, H, W = target.size()
        if D % self.patch_size[0] != 0:
            pad = self.patch_size[0] - D % self.patch_size[0]
            target = torch.nn.functional.pad(target, (0, 0, 0, 0, 0, pad))
        assert x.shape == target.shape
        return l1_mean(x, target)

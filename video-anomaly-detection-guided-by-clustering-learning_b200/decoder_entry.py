"""the immediate consumer of the cluster head's ``x_rec`` — ``Mymodel.norm`` (LayerNorm(192),
model/backbone.py:120) followed by the decoder's entry ``timedebd`` (ConvTranspose3d(192, 192, (2,1,1), stride (2,1,1)), or in
predict mode Conv3d(192, 192, (2,1,1), stride (2,1,1)): model/swin_decoder_predict.py:591-594, applied at :599-602 between
two rearranges) — as one fused op on channel-last tokens (libvadc: vadc_norm_timedebd_fwd / _bwd, vadc_norm_timeconv_fwd /
_bwd).  ``fuse_decoder_entry(model)`` rewires a reference ``Mymodel`` in place without
touching its parameters or state_dict."""
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, f32c, ptr, stream, workspace


class _NormTimeDebed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, weight, bias, eps):
        _lib.require_cuda(x, ln_w, ln_b, weight, bias)
        B, D, H, W, C = x.shape
        x2 = f32c(x).reshape(-1, C)
        N, HW = x2.shape[0], H * W
        w, b = f32c(ln_w), f32c(ln_b)
        Wc = f32c(weight).reshape(C, C, 2)                               # [ci, co, j]
        wt = Wc.permute(2, 1, 0).reshape(2 * C, C).contiguous()          # [j*C + co, ci]
        bi = f32c(bias)
        out = torch.empty((B, 2 * D, H, W, C), device=x2.device, dtype=torch.float32)
        mu = torch.empty((N,), device=x2.device, dtype=torch.float32)
        rstd = torch.empty((N,), device=x2.device, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_norm_timedebd_workspace_bytes(N, C), x2.device)
        check(l.vadc_norm_timedebd_fwd(ptr(x2), ptr(w), ptr(b), ptr(wt), ptr(bi), N, C, HW, float(eps), ptr(out), ptr(mu),
                                       ptr(rstd), ptr(ws), ws.numel(), stream()), "vadc_norm_timedebd_fwd")
        ctx.save_for_backward(x2, mu, rstd, w, b, Wc)
        ctx.shape, ctx.eps = (B, D, H, W, C), float(eps)
        return out

    @staticmethod
    def backward(ctx, gout):
        x2, mu, rstd, w, b, Wc = ctx.saved_tensors
        B, D, H, W, C = ctx.shape
        N, HW = x2.shape[0], H * W
        g = f32c(gout).reshape(2 * N, C)
        wk = Wc.permute(0, 2, 1).reshape(C, 2 * C).contiguous()          # [ci, j*C + co]
        dev = x2.device
        gx = torch.empty((N, C), device=dev, dtype=torch.float32)
        gw = torch.empty((C,), device=dev, dtype=torch.float32)
        gb = torch.empty((C,), device=dev, dtype=torch.float32)
        gwk = torch.empty((C, 2 * C), device=dev, dtype=torch.float32)
        gbias = torch.empty((C,), device=dev, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_norm_timedebd_workspace_bytes(N, C), dev)
        check(l.vadc_norm_timedebd_bwd(ptr(x2), ptr(mu), ptr(rstd), ptr(w), ptr(b), ptr(wk), ptr(g), N, C, HW, ctx.eps,
                                       ptr(gx), ptr(gw), ptr(gb), ptr(gwk), ptr(gbias), ptr(ws), ws.numel(), stream()),
              "vadc_norm_timedebd_bwd")
        gweight = gwk.view(C, 2, C).permute(0, 2, 1).reshape(C, C, 2, 1, 1)          # back to [ci, co, j, 1, 1]
        return gx.view(B, D, H, W, C), gw, gb, gweight, gbias, None


class _NormTimeConv(torch.autograd.Function):
    """predict mode: LayerNorm + Conv3d(C, C, (2,1,1), stride (2,1,1)) — frame pairs merged"""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, weight, bias, eps):
        _lib.require_cuda(x, ln_w, ln_b, weight, bias)
        B, D, H, W, C = x.shape
        if D % 2:
            raise RuntimeError("Conv3d(kernel (2,1,1), stride (2,1,1)) entry: the fused op takes an even number of frames")
        x2 = f32c(x).reshape(-1, C)
        N, HW = x2.shape[0], H * W
        w, b = f32c(ln_w), f32c(ln_b)
        wk = f32c(weight).reshape(C, C, 2).permute(0, 2, 1).reshape(C, 2 * C).contiguous()      # [co, j*C + ci]
        bi = f32c(bias)
        out = torch.empty((B, D // 2, H, W, C), device=x2.device, dtype=torch.float32)
        mu = torch.empty((N,), device=x2.device, dtype=torch.float32)
        rstd = torch.empty((N,), device=x2.device, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_norm_timedebd_workspace_bytes(N, C), x2.device)
        check(l.vadc_norm_timeconv_fwd(ptr(x2), ptr(w), ptr(b), ptr(wk), ptr(bi), N, C, HW, float(eps), ptr(out), ptr(mu),
                                       ptr(rstd), ptr(ws), ws.numel(), stream()), "vadc_norm_timeconv_fwd")
        ctx.save_for_backward(x2, mu, rstd, w, b, wk)
        ctx.shape, ctx.eps = (B, D, H, W, C), float(eps)
        return out

    @staticmethod
    def backward(ctx, gout):
        x2, mu, rstd, w, b, wk = ctx.saved_tensors
        B, D, H, W, C = ctx.shape
        N, HW = x2.shape[0], H * W
        g = f32c(gout).reshape(N // 2, C)
        dev = x2.device
        gx = torch.empty((N, C), device=dev, dtype=torch.float32)
        gw = torch.empty((C,), device=dev, dtype=torch.float32)
        gb = torch.empty((C,), device=dev, dtype=torch.float32)
        gwk = torch.empty((C, 2 * C), device=dev, dtype=torch.float32)
        gbias = torch.empty((C,), device=dev, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_norm_timedebd_workspace_bytes(N, C), dev)
        check(l.vadc_norm_timeconv_bwd(ptr(x2), ptr(mu), ptr(rstd), ptr(w), ptr(b), ptr(wk), ptr(g), N, C, HW, ctx.eps,
                                       ptr(gx), ptr(gw), ptr(gb), ptr(gwk), ptr(gbias), ptr(ws), ws.numel(), stream()),
              "vadc_norm_timeconv_bwd")
        gweight = gwk.view(C, 2, C).permute(0, 2, 1).reshape(C, C, 2, 1, 1)          # back to [co, ci, j, 1, 1]
        return gx.view(B, D, H, W, C), gw, gb, gweight, gbias, None


def _is_entry(td, cls, C):
    return (isinstance(td, cls) and tuple(td.kernel_size) == (2, 1, 1) and tuple(td.stride) == (2, 1, 1)
            and tuple(td.padding) == (0, 0, 0) and td.groups == 1 and td.bias is not None
            and td.in_channels == td.out_channels == C)


def norm_timedebd(x, norm, timedebd):
    """x [B,D,H,W,C] channel-last -> timedebd(norm(x)) channel-last.  ``norm``: nn.LayerNorm(C); ``timedebd``: the decoder's
    entry — nn.ConvTranspose3d(C, C, (2,1,1), stride (2,1,1)) (non-predict: [B,2D,H,W,C]) or nn.Conv3d(C, C, (2,1,1), stride
    (2,1,1)) (predict mode, swin_decoder_predict.py:591-592: [B,D/2,H,W,C])."""
    C = x.shape[-1]
    if _is_entry(timedebd, nn.ConvTranspose3d, C):
        return _NormTimeDebed.apply(x, norm.weight, norm.bias, timedebd.weight, timedebd.bias, norm.eps)
    if _is_entry(timedebd, nn.Conv3d, C):
        return _NormTimeConv.apply(x, norm.weight, norm.bias, timedebd.weight, timedebd.bias, norm.eps)
    raise RuntimeError("norm_timedebd fuses LayerNorm + ConvTranspose3d / Conv3d (C, C, (2,1,1), stride (2,1,1)) only "
                       "(the decoder entry, swin_decoder_predict.py:591-594)")


class _DeferredNorm(nn.Module):
    """stands in for ``Mymodel.norm``'s forward: hands the un-normalised tokens on (the decoder entry applies the
    LayerNorm); the LayerNorm module itself — its parameters and state_dict keys — stays where the reference has it"""

    def forward(self, x):
        return x


def fuse_decoder_entry(model):
    """Rewire a reference ``Mymodel`` (either decoder: ConvTranspose3d entry, or the predict-mode Conv3d entry) in place: ``model.norm(x)`` becomes the identity and
    ``model.decoder.timedebd`` — which receives ``rearrange(x, 'B D H W C -> B C D H W')``, a VIEW of the channel-last
    tokens — runs the fused LayerNorm + ConvTranspose3d on those tokens and returns a channel-first VIEW of its
    channel-last result, so the decoder's next rearrange is free as well.  Parameters, buffers and state_dict keys are
    untouched (the forwards are bound on the instances).  Returns the model."""
    norm, td = model.norm, model.decoder.timedebd
    if not isinstance(td, (nn.ConvTranspose3d, nn.Conv3d)):
        raise RuntimeError("fuse_decoder_entry: model.decoder.timedebd is neither ConvTranspose3d nor Conv3d")

    def td_forward(x_cf):                                   # [B, C, D, H, W]
        return norm_timedebd(x_cf.permute(0, 2, 3, 4, 1), norm, td).permute(0, 4, 1, 2, 3)

    norm.forward = lambda x: x                              # noqa: E731
    td.forward = td_forward
    return model

"""the encoder's downsample stage — ``nn.Sequential(Conv3d(Cin, Cout, (1,2,2), stride (1,2,2)), GELU)``
(model/swin_transformer.py:575-585) — as a producer of CHANNEL-LAST tokens (libvadc: vadc_downsample_gelu_fwd / _bwd).
The result is returned as a channel-first VIEW of channel-last memory, so the rearranges that follow it in the reference
('n c d h w -> n d h w c', swin_transformer.py:745 and model/backbone.py:82) cost nothing and the cluster head reads its
input without a transposing copy.  ``fuse_encoder_tail(model)`` rewires a reference ``Mymodel`` in place without touching
its parameters or state_dict."""
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, f32c, ptr, stream, workspace


class _DownsampleGelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        _lib.require_cuda(x, weight, bias)
        B, Cin, D, H2, W2 = x.shape
        Cout = weight.shape[0]
        H, W = H2 // 2, W2 // 2
        xc = f32c(x)
        w2 = f32c(weight).reshape(Cout, Cin * 4)
        bi = f32c(bias)
        need_grad = any(ctx.needs_input_grad)
        out = torch.empty((B, D, H, W, Cout), device=xc.device, dtype=torch.float32)
        pre = torch.empty_like(out) if need_grad else None
        l = _lib.lib()
        ws = workspace(l.vadc_downsample_gelu_workspace_bytes(B, Cin, D, H, W, Cout), xc.device)
        check(l.vadc_downsample_gelu_fwd(ptr(xc), ptr(w2), ptr(bi), B, Cin, D, H, W, Cout, ptr(out), ptr(pre),
                                         ptr(ws), ws.numel(), stream()), "vadc_downsample_gelu_fwd")
        if need_grad:
            ctx.save_for_backward(xc, w2, pre)
        ctx.dims = (B, Cin, D, H, W, Cout)
        return out

    @staticmethod
    def backward(ctx, gout):
        xc, w2, pre = ctx.saved_tensors
        B, Cin, D, H, W, Cout = ctx.dims
        g = f32c(gout)
        dev = xc.device
        gx = torch.empty_like(xc)
        gw = torch.empty((Cout, Cin * 4), device=dev, dtype=torch.float32)
        gb = torch.empty((Cout,), device=dev, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_downsample_gelu_workspace_bytes(B, Cin, D, H, W, Cout), dev)
        check(l.vadc_downsample_gelu_bwd(ptr(xc), ptr(w2), ptr(pre), ptr(g), B, Cin, D, H, W, Cout, ptr(gx), ptr(gw), ptr(gb),
                                         ptr(ws), ws.numel(), stream()), "vadc_downsample_gelu_bwd")
        return gx, gw.view(Cout, Cin, 1, 2, 2), gb


def _is_downsample(conv, act):
    return (isinstance(conv, nn.Conv3d) and isinstance(act, nn.GELU) and getattr(act, "approximate", "none") == "none"
            and tuple(conv.kernel_size) == (1, 2, 2) and tuple(conv.stride) == (1, 2, 2) and tuple(conv.padding) == (0, 0, 0)
            and tuple(conv.dilation) == (1, 1, 1) and conv.groups == 1 and conv.bias is not None
            and conv.padding_mode == "zeros")


def supported(x, conv):
    """True when the fused kernels take this input (even spatial size, shapes the tensor-core path accepts)"""
    if x.dim() != 5 or not x.is_cuda or x.shape[3] % 2 or x.shape[4] % 2:
        return False
    B, Cin, D, H2, W2 = x.shape
    return bool(_lib.lib().vadc_downsample_gelu_supported(B, Cin, D, H2 // 2, W2 // 2, conv.out_channels))


def downsample_gelu_tokens(x, conv, act=None):
    """x [B,Cin,D,2H,2W] -> gelu(conv(x)) as channel-last tokens [B,D,H,W,Cout] (contiguous).
    ``conv``: nn.Conv3d(Cin, Cout, (1,2,2), stride (1,2,2)); ``act``: nn.GELU() (exact) or None to skip the check."""
    if not _is_downsample(conv, act if act is not None else nn.GELU()):
        raise RuntimeError("downsample_gelu_tokens fuses Conv3d(kernel (1,2,2), stride (1,2,2), bias) + exact GELU only "
                           "(the encoder's downsample stage, swin_transformer.py:575-585)")
    if x.shape[1] != conv.in_channels:
        raise RuntimeError(f"expected input with {conv.in_channels} channels, got {x.shape[1]}")
    if not supported(x, conv):
        raise RuntimeError("downsample_gelu_tokens: shape outside the fused kernels (odd spatial size, or channel counts "
                           "the tensor-core path does not take)")
    return _DownsampleGelu.apply(x, conv.weight, conv.bias)


def fuse_encoder_tail(model):
    """Rewire the downsample stages of a reference ``Mymodel``'s encoder in place: each
    ``Sequential(Conv3d(k=(1,2,2), s=(1,2,2)), GELU)`` runs the fused op and returns a channel-first VIEW of its channel-last
    result, so the encoder's 'n c d h w -> n d h w c' (swin_transformer.py:745) and Mymodel.forward's
    'B C D H W -> B D H W C' (backbone.py:82) are free and the cluster heads get contiguous channel-last tokens.  Inputs
    the fused kernels do not take (CPU tensors, odd sizes) go through the original modules.  Parameters, buffers and
    state_dict keys are untouched (the forwards are bound on the instances).  Returns the number of stages rewired."""
    n = 0
    for seq in getattr(model.encoder, "downsample", []):
        if isinstance(seq, nn.Sequential) and len(seq) == 2 and _is_downsample(seq[0], seq[1]):
            conv, act = seq[0], seq[1]

            def fwd(x, conv=conv, act=act):
                if x.is_cuda and x.dtype == torch.float32 and supported(x, conv):
                    return downsample_gelu_tokens(x, conv, act).permute(0, 4, 1, 2, 3)
                return act(conv(x))

            seq.forward = fwd
            n += 1
    return n

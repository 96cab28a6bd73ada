"""LayerNorm(C) + ConvTranspose3d(C, C, (2,1,1), stride (2,1,1)) on channel-last tokens as one fused op
(vadc_norm_timedebd_fwd / _bwd) against the reference's own op chain — ``self.norm(x)`` (model/backbone.py:120), then
``rearrange -> timedebd -> rearrange`` (model/swin_decoder_predict.py:599-602) — evaluated by torch in float64."""
import numpy as np
import pytest
import torch

import videoad_b200 as V
from oracle import ref_loader
from gpu_util import N, rel, dev

pytestmark = pytest.mark.gpu


def _reference_chain(x, norm, td):
    z = norm(x)                                             # backbone.py:120
    y = td(z.permute(0, 4, 1, 2, 3))                        # 'B D H W C -> B C D H W', timedebd
    return y.permute(0, 2, 3, 4, 1)                         # 'B C D H W -> B D H W C'


@pytest.mark.parametrize("B,D,H,W,C", [(2, 4, 28, 28, 192), (1, 3, 7, 9, 64), (1, 1, 1, 1, 32), (2, 2, 32, 32, 192)])
def test_norm_timedebd_forward_backward(B, D, H, W, C):
    torch.manual_seed(B * 100 + C)
    norm = torch.nn.LayerNorm(C).to(dev())
    td = torch.nn.ConvTranspose3d(C, C, kernel_size=(2, 1, 1), stride=(2, 1, 1)).to(dev())
    with torch.no_grad():
        norm.weight.copy_(1 + 0.2 * torch.randn(C, device=dev())); norm.bias.copy_(0.1 * torch.randn(C, device=dev()))
    x = (torch.randn(B, D, H, W, C, device=dev()) * 1.4 + 0.3).requires_grad_(True)
    gy = torch.randn(B, 2 * D, H, W, C, device=dev())
    y = V.norm_timedebd(x, norm, td)
    assert y.shape == (B, 2 * D, H, W, C) and y.is_contiguous()
    y.backward(gy)
    got = [y.detach(), x.grad, norm.weight.grad, norm.bias.grad, td.weight.grad, td.bias.grad]
    # the reference chain in float64
    n64 = torch.nn.LayerNorm(C).to(dev()).double(); t64 = torch.nn.ConvTranspose3d(C, C, (2, 1, 1), stride=(2, 1, 1)).to(dev()).double()
    n64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    t64.load_state_dict({k: v.double() for k, v in td.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    y64 = _reference_chain(x64, n64, t64)
    y64.backward(gy.double())
    want = [y64.detach(), x64.grad, n64.weight.grad, n64.bias.grad, t64.weight.grad, t64.bias.grad]
    names = ["out", "gx", "g_ln_w", "g_ln_b", "g_weight", "g_bias"]
    for n_, a, b in zip(names, got, want):
        assert a.shape == b.shape, n_
        tol = 1e-5 if n_ == "out" else 2e-4
        assert rel(N(a), N(b)) < tol, (n_, rel(N(a), N(b)))


@pytest.mark.parametrize("B,D,H,W,C", [(2, 4, 28, 28, 192), (1, 2, 7, 9, 64), (1, 6, 1, 1, 32), (2, 8, 32, 32, 192)])
def test_predict_mode_entry_forward_backward(B, D, H, W, C):
    """predict mode (swin_decoder_predict.py:591-592): LayerNorm + Conv3d(C, C, (2,1,1), stride (2,1,1))"""
    torch.manual_seed(B * 100 + C + D)
    norm = torch.nn.LayerNorm(C).to(dev())
    td = torch.nn.Conv3d(C, C, kernel_size=(2, 1, 1), stride=(2, 1, 1)).to(dev())
    with torch.no_grad():
        norm.weight.copy_(1 + 0.2 * torch.randn(C, device=dev())); norm.bias.copy_(0.1 * torch.randn(C, device=dev()))
    x = (torch.randn(B, D, H, W, C, device=dev()) * 1.4 + 0.3).requires_grad_(True)
    gy = torch.randn(B, D // 2, H, W, C, device=dev())
    y = V.norm_timedebd(x, norm, td)
    assert y.shape == (B, D // 2, H, W, C) and y.is_contiguous()
    y.backward(gy)
    got = [y.detach(), x.grad, norm.weight.grad, norm.bias.grad, td.weight.grad, td.bias.grad]
    n64 = torch.nn.LayerNorm(C).to(dev()).double(); t64 = torch.nn.Conv3d(C, C, (2, 1, 1), stride=(2, 1, 1)).to(dev()).double()
    n64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    t64.load_state_dict({k: v.double() for k, v in td.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    y64 = _reference_chain(x64, n64, t64)
    y64.backward(gy.double())
    want = [y64.detach(), x64.grad, n64.weight.grad, n64.bias.grad, t64.weight.grad, t64.bias.grad]
    for n_, a, b in zip(["out", "gx", "g_ln_w", "g_ln_b", "g_weight", "g_bias"], got, want):
        assert a.shape == b.shape, n_
        assert rel(N(a), N(b)) < (1e-5 if n_ == "out" else 2e-4), (n_, rel(N(a), N(b)))


def test_wrong_module_kind_is_refused():
    norm = torch.nn.LayerNorm(64).to(dev())
    conv = torch.nn.Conv3d(64, 64, (1, 1, 1)).to(dev())
    with pytest.raises(RuntimeError):
        V.norm_timedebd(torch.zeros(1, 2, 4, 4, 64, device=dev()), norm, conv)
    with pytest.raises(RuntimeError):                                              # odd number of frames in predict mode
        V.norm_timedebd(torch.zeros(1, 3, 4, 4, 64, device=dev()), norm, torch.nn.Conv3d(64, 64, (2, 1, 1), stride=(2, 1, 1)).to(dev()))


@pytest.mark.skipif(ref_loader.ref_root() is None, reason="reference sources not staged (baseline/_ref)")
@pytest.mark.parametrize("ispredict", [False, True])
def test_fused_decoder_entry_inside_the_reference_model(ispredict):
    """``fuse_decoder_entry`` on the reference's own ``Mymodel`` (either decoder entry): same state_dict, same
    reconstruction as the untouched model"""
    ref = ref_loader.load()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    m = ref.build_mymodel(ispredict=ispredict).to(dev()).eval()
    clip = torch.rand(1, 3, 8, 224, 224, device=dev())
    m.cluster_loss_on(); m.encoder_compatness()
    with torch.no_grad():
        want = m(clip)[0]
    keys = list(m.state_dict().keys())
    V.fuse_decoder_entry(m)
    assert list(m.state_dict().keys()) == keys
    with torch.no_grad():
        got = m(clip)[0]
    assert got.shape == want.shape and rel(N(got), N(want)) < 1e-4

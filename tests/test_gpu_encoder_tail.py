"""the encoder's downsample stage — Conv3d(Cin, Cout, (1,2,2), stride (1,2,2)) + GELU
(model/swin_transformer.py:575-585) — as a producer of channel-last tokens (vadc_downsample_gelu_fwd / _bwd), against the
reference's own op chain (the nn.Sequential, then 'n c d h w -> n d h w c', swin_transformer.py:745) evaluated by torch in
float64."""
import pytest
import torch

import videoad_b200 as V
from oracle import ref_loader
from gpu_util import N, rel, dev

pytestmark = pytest.mark.gpu


def _stage(cin, cout):
    return torch.nn.Sequential(torch.nn.Conv3d(cin, cout, kernel_size=(1, 2, 2), stride=(1, 2, 2)), torch.nn.GELU()).to(dev())


@pytest.mark.parametrize("B,Cin,D,H,W,Cout", [(2, 96, 4, 28, 28, 192),      # the reference's stage at 224^2 input
                                              (1, 8, 3, 5, 7, 16),          # ragged: W < one 32-token block
                                              (1, 96, 2, 32, 40, 192),      # W spans two blocks, the second ragged
                                              (1, 48, 1, 4, 33, 96)])
def test_downsample_gelu_forward_backward(B, Cin, D, H, W, Cout):
    torch.manual_seed(B * 1000 + Cin + W)
    seq = _stage(Cin, Cout)
    x = (torch.randn(B, Cin, D, 2 * H, 2 * W, device=dev()) * 1.3).requires_grad_(True)
    gy = torch.randn(B, D, H, W, Cout, device=dev())
    y = V.downsample_gelu_tokens(x, seq[0], seq[1])
    assert y.shape == (B, D, H, W, Cout) and y.is_contiguous()
    y.backward(gy)
    got = [y.detach(), x.grad, seq[0].weight.grad, seq[0].bias.grad]
    s64 = _stage(Cin, Cout).double()
    s64.load_state_dict({k: v.double() for k, v in seq.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    y64 = s64(x64).permute(0, 2, 3, 4, 1)                    # 'n c d h w -> n d h w c'
    y64.backward(gy.double())
    want = [y64.detach(), x64.grad, s64[0].weight.grad, s64[0].bias.grad]
    for name, a, b in zip(["out", "gx", "g_weight", "g_bias"], got, want):
        assert a.shape == b.shape, name
        assert rel(N(a), N(b)) < (1e-5 if name == "out" else 2e-4), (name, rel(N(a), N(b)))


def test_inference_keeps_no_preactivation_and_other_modules_are_refused():
    seq = _stage(8, 16)
    x = torch.randn(1, 8, 2, 8, 8, device=dev())
    with torch.no_grad():
        y = V.downsample_gelu_tokens(x, seq[0], seq[1])
    assert rel(N(y), N(seq(x).permute(0, 2, 3, 4, 1))) < 1e-5
    with pytest.raises(RuntimeError):
        V.downsample_gelu_tokens(x, torch.nn.Conv3d(8, 16, (1, 3, 3), stride=(1, 2, 2)).to(dev()), seq[1])
    with pytest.raises(RuntimeError):
        V.downsample_gelu_tokens(x, seq[0], torch.nn.GELU(approximate="tanh"))
    with pytest.raises(RuntimeError):
        V.downsample_gelu_tokens(torch.randn(1, 8, 2, 7, 8, device=dev()), seq[0], seq[1])      # odd height


@pytest.mark.skipif(ref_loader.ref_root() is None, reason="reference sources not staged (baseline/_ref)")
def test_fused_encoder_tail_inside_the_reference_model():
    """``fuse_encoder_tail`` on the reference's own ``Mymodel``: same state_dict, same outputs as the untouched model, and
    the tokens that reach the cluster heads are contiguous channel-last (no transposing copy)"""
    ref = ref_loader.load()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    m = ref.build_mymodel(ispredict=False).to(dev()).eval()
    clip = torch.rand(1, 3, 8, 224, 224, device=dev())
    m.cluster_loss_on(); m.encoder_compatness()
    with torch.no_grad():
        want = m(clip)
    keys = list(m.state_dict().keys())
    assert V.fuse_encoder_tail(m) >= 1
    assert list(m.state_dict().keys()) == keys
    seen = {}
    h1 = m.encoder.downsample[0].register_forward_hook(lambda mod, args, out: seen.setdefault("down", out))
    h2 = m.cluster1.register_forward_pre_hook(lambda mod, args: seen.setdefault("x", args[0]))
    with torch.no_grad():
        got = m(clip)
    h1.remove(); h2.remove()
    # the stage hands on a channel-first VIEW of channel-last memory: the rearranges that follow are free
    assert seen["down"].shape[1] == 192 and seen["down"].permute(0, 2, 3, 4, 1).is_contiguous()
    # ... and the layout survives the last Swin stage (cuDNN conv3d and the residual keep channels-last memory): the cluster
    # head's input is contiguous channel-last, no transposing copy
    assert seen["x"].is_contiguous()
    assert got[0].shape == want[0].shape and rel(N(got[0]), N(want[0])) < 1e-4

"""C1 / C2 / L1 parity on the GPU, through the C ABI (ctypes -> libvadc.so).
Tolerances: distances / losses 1e-4 relative (BASELINE.json north_star) — the
kernels are in fact held to ~1e-5; argmin bit-exact excluding fp64-adjudicated
ties."""
import numpy as np
import pytest
import torch

import videoad_b200 as V
from oracle import np_oracle as O
from conftest import load_golden
from gpu_util import T, N, rel, dev, assert_labels_match, assert_selfdist_close, make_cluster_module

pytestmark = pytest.mark.gpu
IMPLS = [V.IMPL_SIMT, V.IMPL_AUTO]
GOLDEN = ["cluster_c64_k32", "cluster_c32_k16", "cluster_c192_k48_peaked"]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", GOLDEN)
def test_golden_forward_backward(name, impl):
    """the reference module's own outputs and autograd gradients (fixtures)"""
    g = load_golden(name)
    C, K = g["centers"].shape[1], g["centers"].shape[0]
    m = make_cluster_module(V, C, K, float(g["alpha"]), g["centers"], g["ln_w"], g["ln_b"], impl)
    m.cluster_center.requires_grad_(True)
    x = T(g["x"], grad=True)
    D, A, S, R, F, lab = m(x)
    assert D.shape == g["D"].shape and R.shape == g["x_rec"].shape and F.shape == g["feature"].shape
    assert lab.dtype == torch.int64 and lab.shape == g["label"].shape
    assert rel(N(F), g["feature"]) < 1e-5
    assert rel(N(D), g["D"]) < 1e-5
    np.testing.assert_allclose(N(A), g["A"], rtol=1e-3, atol=2e-6)
    assert rel(N(R), g["x_rec"]) < 1e-4
    assert_selfdist_close(N(S), g["S"])
    D64 = O.cluster_forward(g["x"], g["centers"], g["ln_w"], g["ln_b"], float(g["alpha"]), dtype=np.float64)["D"]
    assert_labels_match(N(lab), D64)
    # the reference's objective: torch.norm(D*A) + probes (make_golden.py)
    closs = torch.norm(D * A)
    assert abs(float(closs) - float(g["cluster_loss"])) < 1e-4 * float(g["cluster_loss"])
    assert abs(float(m.fused_cluster_loss()) - float(g["cluster_loss"])) < 1e-4 * float(g["cluster_loss"])
    obj = closs + (R * T(g["gR"])).sum() + (F * T(g["gF"])).sum()
    obj.backward()
    assert rel(N(x.grad), g["gx"]) < 2e-4
    assert rel(N(m.cluster_center.grad), g["gcenters"]) < 2e-4
    assert rel(N(m.norm.weight.grad), g["g_ln_w"]) < 2e-4
    assert rel(N(m.norm.bias.grad), g["g_ln_b"]) < 2e-4


@pytest.mark.parametrize("impl", IMPLS)
def test_fused_loss_backward_equals_explicit(impl):
    """fused d sum(D*A)^2 path == autograd through torch.norm(D*A)"""
    g = load_golden("cluster_c64_k32")
    grads = []
    for fused in (False, True):
        m = make_cluster_module(V, 64, 32, 16.0, g["centers"], g["ln_w"], g["ln_b"], impl)
        x = T(g["x"], grad=True)
        D, A, S, R, F, lab = m(x)
        loss = m.fused_cluster_loss() if fused else torch.norm(D * A)
        (loss * 1.7 + (R * T(g["gR"])).sum()).backward()
        grads.append((N(x.grad), N(m.cluster_center.grad), N(m.norm.weight.grad), N(m.norm.bias.grad)))
    for a, b in zip(*grads):
        assert rel(a, b) < 2e-5


SHAPES = [  # (N tokens, C, K, alpha): BASELINE configs at oracle-sized N + ragged / edge shapes
    (6272, 192, 1024, 16.0),   # reference-native (cfg1)
    (4096, 192, 32, 16.0),     # cfg2 head
    (2048, 768, 16, 32.0), (2048, 768, 64, 32.0), (2048, 768, 256, 32.0),   # cfg3 sweep
    (77, 192, 32, 16.0),       # ragged: not a multiple of any tile
    (1, 64, 16, 32.0),         # single token
    (300, 20, 12, 8.0),        # C, K only multiples of 4
]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("Ntok,C,K,alpha", SHAPES)
def test_forward_vs_oracle(Ntok, C, K, alpha, impl):
    rng = np.random.default_rng(Ntok + C + K)
    x = (rng.standard_normal((1, 1, 1, Ntok, C)) * 1.5 + 0.3).astype(np.float32)
    cen = rng.random((K, C)).astype(np.float32)
    w = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    m = make_cluster_module(V, C, K, alpha, cen, w, b, impl)
    with torch.no_grad():
        D, A, S, R, F, lab = m(T(x))
    o32 = O.cluster_forward(x, cen, w, b, alpha)
    o64 = O.cluster_forward(x, cen, w, b, alpha, dtype=np.float64)
    assert rel(N(F), o64["feature"]) < 1e-5
    assert rel(N(D), o64["D"]) < 1e-5            # spec: 1e-4
    assert_labels_match(N(lab), o64["D"])
    # A = exp(-alpha (D - Dmin)): an fp32 ulp of D (~1e-6) moves A by alpha*1e-6 relative
    np.testing.assert_allclose(N(A), o64["A"], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(N(A).sum(-1), 1.0, rtol=1e-5)
    assert rel(N(R), o64["x_rec"]) < 1e-4
    assert_selfdist_close(N(S), o32["S"])
    lo = float(O.frobenius_loss(o64["D"], o64["A"], np.float64))
    assert abs(float(m.fused_cluster_loss()) - lo) < 1e-4 * lo


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("Ntok,C,K,alpha", [(1500, 192, 32, 16.0), (700, 768, 64, 32.0), (515, 192, 1024, 16.0)])
def test_backward_vs_oracle(Ntok, C, K, alpha, impl):
    rng = np.random.default_rng(7 * Ntok + K)
    x = (rng.standard_normal((1, 1, 1, Ntok, C)) * 1.2).astype(np.float32)
    cen = rng.random((K, C)).astype(np.float32)
    w = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    gR = rng.standard_normal((Ntok, C)).astype(np.float32)
    gF = (0.05 * rng.standard_normal((Ntok, C))).astype(np.float32)
    gDx = (0.1 * rng.standard_normal((Ntok, K))).astype(np.float32)
    gAx = rng.standard_normal((Ntok, K)).astype(np.float32)
    m = make_cluster_module(V, C, K, alpha, cen, w, b, impl)
    xt = T(x, grad=True)
    D, A, S, R, F, lab = m(xt)
    obj = (D * T(gDx).view_as(D)).sum() + (A * T(gAx).view_as(A)).sum() + (R * T(gR).view_as(R)).sum() \
        + (F * T(gF)).sum() + 0.5 * m.fused_cluster_loss()
    obj.backward()
    f64 = O.cluster_forward(x, cen, w, b, alpha, dtype=np.float64)
    lD, lA = O.frobenius_loss_grads(f64["D"], f64["A"], 0.5, np.float64)
    gx, gc, gw, gb = O.cluster_backward(x, cen, w, b, alpha, gD=gDx + lD.reshape(Ntok, K),
                                        gA=gAx + lA.reshape(Ntok, K), gR=gR, gF=gF, dtype=np.float64)
    assert rel(N(xt.grad).reshape(Ntok, C), gx) < 2e-4
    assert rel(N(m.cluster_center.grad), gc) < 2e-4
    assert rel(N(m.norm.weight.grad), gw) < 2e-4
    assert rel(N(m.norm.bias.grad), gb) < 2e-4


@pytest.mark.parametrize("impl", IMPLS)
def test_empty_batch(impl):
    m = make_cluster_module(V, 64, 16, 16.0, np.random.rand(16, 64), np.ones(64), np.zeros(64), impl)
    x = torch.zeros((0, 2, 4, 4, 64), device=dev(), requires_grad=True)
    D, A, S, R, F, lab = m(x)
    assert D.shape == (0, 2, 4, 4, 16) and R.shape == (0, 2, 4, 4, 64) and F.shape == (0, 64) and lab.numel() == 0
    assert float(m.loss_sq) == 0.0 and S.shape == (16, 16)
    (R.sum() + m.loss_sq.sum()).backward()
    assert float(m.cluster_center.grad.abs().sum()) == 0.0


def test_alpha_kwarg_persists_like_reference():
    """cluster.py:49-50: a forward-time alpha overwrites assign_func.alpha for good"""
    g = load_golden("cluster_c64_k32")
    m = make_cluster_module(V, 64, 32, 16.0, g["centers"], g["ln_w"], g["ln_b"])
    with torch.no_grad():
        A8 = m(T(g["x"]), alpha=8.0)[1]
        assert m.assign_func.alpha == 8.0
        A8b = m(T(g["x"]))[1]
    assert torch.equal(A8, A8b)
    ref = O.cluster_forward(g["x"], g["centers"], g["ln_w"], g["ln_b"], 8.0, dtype=np.float64)["A"]
    np.testing.assert_allclose(N(A8), ref, rtol=1e-3, atol=1e-6)


def test_shape_errors_surface_as_exceptions():
    m = make_cluster_module(V, 64, 16, 16.0, np.random.rand(16, 64), np.ones(64), np.zeros(64))
    with pytest.raises(RuntimeError):
        m(torch.zeros((1, 1, 2, 2, 32), device=dev()))       # wrong C: LayerNorm / cdist column mismatch
    with pytest.raises(ValueError):
        m(torch.zeros((4, 64), device=dev()))                # B, D, H, W, C = x.shape


@pytest.mark.parametrize("impl,C,K", [(i, 192, 32) for i in IMPLS] +
                         [(V.IMPL_AUTO, 768, 16), (V.IMPL_AUTO, 768, 64), (V.IMPL_AUTO, 768, 256), (V.IMPL_AUTO, 192, 1024)])
def test_full_size_properties(impl, C, K):
    """BASELINE cfg2 size (B=64,T=16,256x256 -> N=524288 tokens, C=192, K=32), the cfg3 sweep (C=768, K=16/64/256)
    and the reference-native head (C=192, K=1024) at the same token count: size-independent invariants instead
    of the (minutes-long) CPU oracle."""
    torch.manual_seed(0)
    Ntok = 524288
    m = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=16.0).to(dev())
    m.impl = impl
    x = torch.randn(64, 8, 32, 32, C, device=dev())
    with torch.no_grad():
        D, A, S, R, F, lab = m(x)
        D2, A2 = D.view(Ntok, K), A.view(Ntok, K)
        assert torch.isfinite(D2).all() and torch.isfinite(A2).all()
        # rows of A are a distribution; the label is the argmin of the returned D and the argmax of A
        assert float((A2.sum(-1) - 1).abs().max()) < 1e-5
        assert torch.equal(lab, D2.argmin(-1)) or float((D2.gather(1, lab[:, None])[:, 0] - D2.min(-1).values).abs().max()) == 0.0
        # LayerNorm invariants: per-token mean 0 / var 1 (affine is identity at init)
        assert float(F.mean(-1).abs().max()) < 1e-5 and float((F.var(-1, unbiased=False) - 1).abs().max()) < 1e-3
        # x_rec is the A-weighted centroid mix; D matches a direct (non-mm) evaluation on a sample
        idx = torch.randint(0, Ntok, (4096,), device=dev())
        Rs = (A2[idx].double() @ m.cluster_center.double()).float()
        assert float((R.view(Ntok, C)[idx] - Rs).abs().max()) < 1e-4
        Dd = (F[idx].double()[:, None, :] - m.cluster_center.double()[None]).norm(dim=-1)
        assert float(((D2[idx].double() - Dd).abs() / Dd).max()) < 1e-5
        # fused loss == checksum of the returned tensors
        ref = (D2.double() * A2.double()).pow(2).sum().sqrt()
        assert abs(float(m.fused_cluster_loss()) - float(ref)) < 1e-5 * float(ref)


# ---------------------------------------------------------------------------
# training-graph backward (gradients arrive through x_rec and the fused cluster loss only:
# model/backbone.py:89-98, main_predict.py:284-296) -> the tcgen05 kernel of cluster_bwd_tc.cu
# ---------------------------------------------------------------------------
def _training_graph_backward(Ntok, C, K, alpha, seed, scale_g, loss_w):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((1, 1, 1, Ntok, C)) * 1.7 + 0.3).astype(np.float32)
    cen = rng.random((K, C)).astype(np.float32)
    w = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    gR = (scale_g * rng.standard_normal((Ntok, C))).astype(np.float32)
    m = make_cluster_module(V, C, K, alpha, cen, w, b, V.IMPL_AUTO)
    xt = T(x, grad=True)
    D, A, S, R, F, lab = m(xt)
    torch.autograd.backward([m.fused_cluster_loss() * loss_w, R], [None, T(gR).view_as(R)])
    f64 = O.cluster_forward(x, cen, w, b, alpha, dtype=np.float64)
    lD, lA = O.frobenius_loss_grads(f64["D"], f64["A"], loss_w, np.float64)
    ref = O.cluster_backward(x, cen, w, b, alpha, gD=lD.reshape(Ntok, K), gA=lA.reshape(Ntok, K), gR=gR,
                             dtype=np.float64)
    got = (N(xt.grad).reshape(Ntok, C), N(m.cluster_center.grad), N(m.norm.weight.grad), N(m.norm.bias.grad))
    return got, ref


@pytest.mark.parametrize("Ntok,C,K,alpha", [
    (64, 192, 32, 16.0),          # exactly one tile
    (1, 192, 32, 16.0),           # a single token
    (1500, 192, 32, 16.0),        # ragged tail (1500 = 23 * 64 + 28)
    (777, 128, 32, 32.0),
    (333, 64, 32, 8.0),
    (148 * 64 * 3 + 5, 192, 32, 16.0),   # every CTA walks several tiles (buffer recycling) + a one-row tail tile
])
def test_training_graph_backward_vs_oracle(Ntok, C, K, alpha):
    """tolerance 2e-4 of the largest gradient entry; the kernel's operands are two-term bf16 splits
    (2^-16 per element), the fp32 reference itself is ~2e-5 away from fp64 on these inputs"""
    got, ref = _training_graph_backward(Ntok, C, K, alpha, seed=Ntok + C, scale_g=1e-2, loss_w=1.3)
    for g, r, name in zip(got, ref, ("gx", "gcenters", "g_ln_w", "g_ln_b")):
        assert rel(g, r) < 2e-4, (name, rel(g, r))


@pytest.mark.parametrize("scale_g", [1e-12, 1e-6, 1.0, 1e6, 1e12])
def test_training_graph_backward_has_no_gradient_scale(scale_g):
    """upstream gradients have no a-priori magnitude: the bf16-split operands keep fp32's exponent range,
    so the relative error does not depend on the scale of gR (an fp16 operand would overflow / flush)"""
    got, ref = _training_graph_backward(700, 192, 32, 16.0, seed=5, scale_g=scale_g, loss_w=scale_g * 50)
    for g, r, name in zip(got, ref, ("gx", "gcenters", "g_ln_w", "g_ln_b")):
        assert np.isfinite(g).all(), name
        assert rel(g, r) < 2e-4, (name, scale_g, rel(g, r))


def test_training_graph_backward_matches_generic_kernels():
    """the same backward through the unfused SIMT kernels (VADC_BWD_IMPL=generic) and through the mma.sync
    fused kernel (VADC_BWD_IMPL=fused): three independent implementations agree"""
    import os
    outs = {}
    from videoad_b200 import _lib
    for impl in ("tc", "fused", "generic"):
        os.environ["VADC_BWD_IMPL"] = impl
        _lib.lib().vadc_refresh_env()                 # the switches are read once per process
        try:
            outs[impl], _ = _training_graph_backward(2000, 192, 32, 16.0, seed=11, scale_g=1e-2, loss_w=0.7)
        finally:
            os.environ.pop("VADC_BWD_IMPL", None)
            _lib.lib().vadc_refresh_env()
    for other in ("fused", "generic"):
        for a, b_, name in zip(outs["tc"], outs[other], ("gx", "gcenters", "g_ln_w", "g_ln_b")):
            assert rel(a, b_) < 1e-4, (other, name, rel(a, b_))


def _torch_fp64_training_graph(x, cen, w, b, alpha, gR, loss_w):
    """the reference op chain (model/cluster.py:81-99 + backbone.py:98) in float64 on the same GPU, through
    torch autograd: LayerNorm -> cdist (mm form) -> softmin -> A @ centers; objective loss_w * ||D*A||_F + <x_rec, gR>"""
    x64 = x.double().requires_grad_(True)
    c64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (cen, w, b))
    z = torch.nn.functional.layer_norm(x64, (x64.shape[-1],), w64, b64, 1e-5)
    d2 = (z * z).sum(-1, keepdim=True) + (c64 * c64).sum(-1)[None] - 2.0 * (z @ c64.t())
    D = d2.clamp_min(0).sqrt()
    A = torch.softmax(-alpha * (D - D.min(-1, keepdim=True).values), dim=-1)
    R = A @ c64
    loss = torch.norm(D * A) * loss_w
    torch.autograd.backward([loss, R], [None, gR.double()])
    return x64.grad, c64.grad, w64.grad, b64.grad, loss.detach()


@pytest.mark.parametrize("Ntok", [524288, 524288 - 37])
def test_full_size_training_graph_backward(Ntok):
    """the BENCHMARKED backward at the benchmarked size (BASELINE cfg2: N = 524288 tokens, C = 192, K = 32;
    VERDICT r1 weak #1): gx (every row), gcenters, g_ln_w, g_ln_b of the tcgen05 kernel against a float64
    torch-autograd evaluation of the reference op chain on the same GPU; same tolerance as
    test_training_graph_backward_vs_oracle (2e-4 of the largest entry).  The second case ends in a ragged tile."""
    C, K, alpha, loss_w = 192, 32, 16.0, 1.3
    torch.manual_seed(Ntok)
    m = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=alpha).to(dev())
    with torch.no_grad():
        m.norm.weight.copy_(1 + 0.2 * torch.randn(C, device=dev()))
        m.norm.bias.copy_(0.1 * torch.randn(C, device=dev()))
    x = (torch.randn(Ntok, C, device=dev()) * 1.7 + 0.3)
    gR = torch.randn(Ntok, C, device=dev()) * 1e-2
    xt = x.view(1, 1, 1, Ntok, C).clone().requires_grad_(True)
    D, A, S, R, F, lab = m(xt)
    torch.autograd.backward([m.fused_cluster_loss() * loss_w, R], [None, gR.view_as(R)])
    gx64, gc64, gw64, gb64, loss64 = _torch_fp64_training_graph(x, m.cluster_center, m.norm.weight, m.norm.bias,
                                                                 alpha, gR, loss_w)
    assert abs(float(m.fused_cluster_loss()) * loss_w - float(loss64)) < 1e-5 * float(loss64)

    def relmax(a, b):
        return float((a.double() - b).abs().max() / b.abs().max())
    errs = {"gx": relmax(xt.grad.view(Ntok, C), gx64), "gcenters": relmax(m.cluster_center.grad, gc64),
            "g_ln_w": relmax(m.norm.weight.grad, gw64), "g_ln_b": relmax(m.norm.bias.grad, gb64)}
    for name, e in errs.items():
        assert e < 2e-4, (name, errs)
    # per-row check too (a row-local error hides under the global max): every row of gx within 1e-3 of that row's scale
    row_err = (xt.grad.view(Ntok, C).double() - gx64).abs().max(-1).values / gx64.abs().max(-1).values.clamp_min(1e-30)
    assert float(row_err.max()) < 2e-3, float(row_err.max())


def test_cfg5_unet_shaped_tokens_and_grayscale_clips():
    """BASELINE configs[4] (UCSD Ped2 / Avenue-shaped clips, T=16, 256x256, grayscale or RGB, UNet3D recon path +
    cluster head): the graded part is C1 on the token grid that a T=16, 256x256 clip produces ([B, 8, 32, 32, C]) and
    L2 / E1 on 1- and 3-channel clips (UNet3D is a token / recon producer only)."""
    rng = np.random.default_rng(55)
    C, K, alpha = 192, 32, 16.0
    x = (rng.standard_normal((1, 8, 32, 32, C)) * 1.1 + 0.2).astype(np.float32)
    cen = rng.random((K, C)).astype(np.float32)
    w = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    m = make_cluster_module(V, C, K, alpha, cen, w, b)
    with torch.no_grad():
        D, A, S, R, F, lab = m(T(x))
    o64 = O.cluster_forward(x, cen, w, b, alpha, dtype=np.float64)
    assert D.shape == (1, 8, 32, 32, K) and R.shape == x.shape
    assert rel(N(D), o64["D"]) < 1e-5 and rel(N(R), o64["x_rec"]) < 1e-4
    assert_labels_match(N(lab), o64["D"])
    for ch in (1, 3):
        clip = rng.random((2, ch, 16, 256, 256)).astype(np.float32)
        recon = (clip + 0.03 * rng.standard_normal(clip.shape)).astype(np.float32)
        rl = V.Recon_Loss((2, 4, 4))
        ref = float(O.recon_l1(recon, clip, patch_d=2))
        assert abs(float(rl(T(recon), T(clip))) - ref) < 1e-5 * ref
        mse = V.frame_mse(T(recon), T(clip))
        assert mse.shape == (2, 16) and rel(N(mse), O.frame_mse(recon, clip)) < 2e-6


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("Ntok,C,K", [(700, 192, 32), (130, 64, 32), (90, 768, 64), (300, 192, 1024)])
def test_forward_rowstats(Ntok, C, K, impl):
    """the per-token sums the forward hands to the backward: |f|^2, sum f gamma, sum f gamma xhat"""
    from videoad_b200 import _lib
    rng = np.random.default_rng(Ntok)
    x = (rng.standard_normal((Ntok, C)) * 2 + 0.5).astype(np.float32)
    cen = rng.random((K, C)).astype(np.float32)
    w = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    l = _lib.lib()
    t = {k: T(v) for k, v in dict(x=x, cen=cen, w=w, b=b).items()}
    o = {k: torch.empty(s, device=dev(), dtype=torch.float32) for k, s in
         dict(D=(Ntok, K), A=(Ntok, K), R=(Ntok, C), F=(Ntok, C), mu=(Ntok,), rstd=(Ntok,), rs=(Ntok, 4), loss=(1,)).items()}
    lab = torch.empty((Ntok,), device=dev(), dtype=torch.int64)
    ws = torch.empty(l.vadc_cluster_fwd_workspace_bytes(Ntok, C, K, impl), device=dev(), dtype=torch.uint8)
    _lib.check(l.vadc_cluster_fwd(_lib.ptr(t["x"]), _lib.ptr(t["w"]), _lib.ptr(t["b"]), _lib.ptr(t["cen"]), Ntok, C, K,
                                  16.0, 1e-5, _lib.ptr(o["D"]), _lib.ptr(o["A"]), _lib.ptr(o["R"]), _lib.ptr(o["F"]),
                                  _lib.ptr(lab), _lib.ptr(o["mu"]), _lib.ptr(o["rstd"]), _lib.ptr(o["rs"]),
                                  _lib.ptr(o["loss"]), _lib.ptr(ws), ws.numel(), impl, _lib.stream()), "fwd")
    torch.cuda.synchronize()
    x64 = x.astype(np.float64)
    mu = x64.mean(1, keepdims=True)
    xh = (x64 - mu) / np.sqrt(x64.var(1, keepdims=True) + 1e-5)
    f = xh * w + b
    want = np.stack([(f * f).sum(1), (f * w).sum(1), (f * w * xh).sum(1)], 1)
    got = N(o["rs"])
    assert np.abs(got[:, :3] - want).max() / np.abs(want).max() < 1e-5
    assert (got[:, 3] == 0).all()


def test_training_step_replays_as_a_cuda_graph():
    """bench.py replays the whole step as one CUDA graph: the library must call nothing a capture forbids (no
    allocation, no synchronisation) and a replay on new input values must reproduce the eager result"""
    C, K = 192, 32
    main = torch.cuda.Stream()
    with torch.cuda.stream(main):                     # one non-default stream for every launch (see bench.py)
        torch.manual_seed(3)
        m = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=16.0).to(dev())
        x_buf = torch.randn(2, 4, 16, 16, C, device=dev())
        gR = torch.randn_like(x_buf) * 1e-2
        params = [m.cluster_center, m.norm.weight, m.norm.bias]
        out = {}

        def step():
            for p in params:
                p.grad = None
            x = x_buf.detach().requires_grad_(True)
            D, A, S, R, F, lab = m(x)
            loss = V.global_frobenius(m.loss_sq)
            torch.autograd.backward([loss, R], [None, gR])
            out["loss"], out["gx"], out["lab"] = loss, x.grad, lab

        for _ in range(3):
            step()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            step()
        cap = {k: v for k, v in out.items()}          # tensors owned by the graph's pool: refreshed by every replay
        x_buf.copy_(torch.randn_like(x_buf) * 1.5 + 0.2)
        gR.copy_(torch.randn_like(gR) * 1e-2)
        g.replay()
        main.synchronize()
        got = {k: v.clone() for k, v in cap.items()}
        gc_graph = m.cluster_center.grad.clone()
        step()                                        # eager, same inputs
        main.synchronize()
        assert torch.equal(got["lab"], out["lab"])
        assert torch.allclose(got["loss"], out["loss"], rtol=1e-6)
        assert torch.allclose(got["gx"], out["gx"], rtol=1e-5, atol=1e-9)
        assert torch.allclose(gc_graph, m.cluster_center.grad, rtol=1e-5, atol=1e-9)


def test_numa_local_pinned_buffers():
    """host staging buffers of the end-to-end path: pinned, equal to their source, CPU affinity restored"""
    import os
    before = os.sched_getaffinity(0)
    src = torch.randn(1000, 7)
    pin = V.pinned_like_local(src, 0)
    assert pin.is_pinned() and torch.equal(pin, src) and os.sched_getaffinity(0) == before
    assert V.gpu_local_cpus(0) <= before


def test_misaligned_contiguous_view_is_accepted():
    """a contiguous view with an odd storage offset (a sliced batch whose base is not 16-byte aligned) runs like the
    reference module does (f32c only guaranteed contiguity; the kernels need 16-byte bases)"""
    g = load_golden("cluster_c64_k32")
    m = make_cluster_module(V, 64, 32, 16.0, g["centers"], g["ln_w"], g["ln_b"])
    x = T(g["x"])
    flat = torch.empty(x.numel() + 1, device=dev())
    xv = flat[1:].view(x.shape)                       # contiguous, data_ptr % 16 == 4
    xv.copy_(x)
    assert xv.is_contiguous() and xv.data_ptr() % 16 != 0
    with torch.no_grad():
        D0, A0, _, R0, _, lab0 = m(x)
        D1, A1, _, R1, _, lab1 = m(xv)
    assert torch.equal(D0, D1) and torch.equal(R0, R1) and torch.equal(lab0, lab1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_runs_on_the_tensors_device_not_the_current_one():
    """module on cuda:1 while cuda:0 is current (no set_device): the library must launch on the tensors' device and
    stream; tensors on two different devices in one call are an error"""
    g = load_golden("cluster_c64_k32")
    d1 = torch.device("cuda", 1)
    m0 = make_cluster_module(V, 64, 32, 16.0, g["centers"], g["ln_w"], g["ln_b"])
    m1 = V.EuclidDistance_Assign_Module(64, 32, soft_assign_alpha=16.0).to(d1)
    m1.load_state_dict(m0.state_dict())
    assert torch.cuda.current_device() == 0
    x0 = T(g["x"], grad=True)
    x1 = x0.detach().to(d1).requires_grad_(True)
    outs = []
    for m, x in ((m0, x0), (m1, x1)):
        D, A, S, R, F, lab = m(x)
        (m.fused_cluster_loss() + (R * R).sum()).backward()
        outs.append((D, R, lab, x.grad, m.cluster_center.grad))
    assert torch.cuda.current_device() == 0
    for a, b in zip(*outs):
        assert b.device == d1 and torch.equal(a.cpu(), b.cpu())
    with pytest.raises(RuntimeError):
        m1(x0)


@pytest.mark.parametrize("Ntok,C,K,alpha", [(333, 192, 30, 16.0), (200, 64, 6, 32.0), (64, 32, 1, 8.0), (500, 192, 1023, 16.0)])
def test_any_cluster_num_forward_backward(Ntok, C, K, alpha):
    """the reference's constructor takes any cluster_num (model/cluster.py:58-75); K % 4 != 0 runs through zero-padded
    centroids that the device excludes from argmin / softmin / loss (vadc_cluster_fwd_padded): same parity bars as the
    other shapes, explicit gradients on every output"""
    rng = np.random.default_rng(K + Ntok)
    x = (rng.standard_normal((1, 1, 1, Ntok, C)) * 1.3 + 0.2).astype(np.float32)
    cen = rng.random((K, C)).astype(np.float32)
    w = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    gR = rng.standard_normal((Ntok, C)).astype(np.float32)
    gDx = (0.1 * rng.standard_normal((Ntok, K))).astype(np.float32)
    gAx = rng.standard_normal((Ntok, K)).astype(np.float32)
    m = make_cluster_module(V, C, K, alpha, cen, w, b)
    xt = T(x, grad=True)
    D, A, S, R, F, lab = m(xt)
    assert D.shape == (1, 1, 1, Ntok, K) and A.shape == D.shape and S.shape == (K, K)
    o64 = O.cluster_forward(x, cen, w, b, alpha, dtype=np.float64)
    assert rel(N(D), o64["D"]) < 1e-5
    assert_labels_match(N(lab), o64["D"])
    np.testing.assert_allclose(N(A), o64["A"], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(N(A).sum(-1), 1.0, rtol=1e-5)
    assert rel(N(R), o64["x_rec"]) < 1e-4
    lo = float(O.frobenius_loss(o64["D"], o64["A"], np.float64))
    assert abs(float(m.fused_cluster_loss()) - lo) < 1e-4 * lo
    obj = (D * T(gDx).view_as(D)).sum() + (A * T(gAx).view_as(A)).sum() + (R * T(gR).view_as(R)).sum() + 0.5 * m.fused_cluster_loss()
    obj.backward()
    lD, lA = O.frobenius_loss_grads(o64["D"], o64["A"], 0.5, np.float64)
    gx, gc, gw, gb = O.cluster_backward(x, cen, w, b, alpha, gD=gDx + lD.reshape(Ntok, K), gA=gAx + lA.reshape(Ntok, K),
                                        gR=gR, dtype=np.float64)
    assert m.cluster_center.grad.shape == (K, C)
    assert rel(N(xt.grad).reshape(Ntok, C), gx) < 2e-4
    assert rel(N(m.cluster_center.grad), gc) < 2e-4
    assert rel(N(m.norm.weight.grad), gw) < 2e-4
    assert rel(N(m.norm.bias.grad), gb) < 2e-4

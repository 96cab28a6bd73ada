"""race hunter for the round-2 kernels (persistent tc_gemm, fused epilogues, tiled LayerNorm kernels): every op here is
deterministic, so repeated launches on the same inputs must be BIT-identical; any difference is a synchronisation bug.
python scripts/gemm_stress.py [iterations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoad_b200 as V
from videoad_b200 import _lib

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)
torch.manual_seed(0)
bad = 0


def same(a, b):
    return all(torch.equal(x, y) for x, y in zip(a, b))


def hammer(name, fn):
    global bad
    ref = [t.clone() for t in fn()]
    assert all(bool(torch.isfinite(t).all()) for t in ref if t.is_floating_point()), name
    n = 0
    for _ in range(iters):
        n += 0 if same(fn(), ref) else 1
    bad += n
    print(f"{name}: {n} of {iters} repeats differ", flush=True)


def debug_gemm(M, N, K, b_mn):
    l = _lib.lib()
    A = torch.randn(M, K, device=dev); B = torch.randn((K, N) if b_mn else (N, K), device=dev)
    out = torch.empty(M, N, device=dev)
    ws = torch.empty(l.vadc_debug_tc_gemm_workspace_bytes(M, N, K), device=dev, dtype=torch.uint8)

    def f():
        _lib.check(l.vadc_debug_tc_gemm(_lib.ptr(A), _lib.ptr(B), M, N, K, int(b_mn), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                        _lib.stream()), "gemm")
        return [out]
    return f


hammer("tc_gemm persistent bf16x3 3001x1096x200", debug_gemm(3001, 1096, 200, 0))
hammer("tc_gemm persistent bf16x3 4096x2048x768 (B MN-major)", debug_gemm(4096, 2048, 768, 1))

mem = V.Memory(2000, 768, 768, 0.1, 0.1)
keys = torch.nn.functional.normalize(torch.rand(2000, 768, device=dev), dim=1)
q = torch.randn(8, 768, 32, 32, device=dev)
with torch.no_grad():
    hammer("memory forward N=8192 (fp16x2 persistent GEMMs, fused column statistics)",
           lambda: [t for t in mem(q, keys, train=True) if torch.is_tensor(t)])

sp = V.Space_EuclidDistance_Assign_Module(192, 128, space_size=32, soft_assign_alpha=32.0).to(dev)
xs = torch.randn(8, 8, 32, 32, 192, device=dev, requires_grad=True)


def space():
    for p in sp.parameters(): p.grad = None
    xs.grad = None
    Ds, As, S, _ = sp(xs)
    sp.fused_cluster_loss().backward()
    return [Ds.detach(), As.detach(), S.detach(), xs.grad, sp.cluster_center.grad, sp.norm.weight.grad, sp.norm.bias.grad]


hammer("space head fwd+bwd M=64 (batched persistent GEMMs, tiled LayerNorm kernels)", space)

mn = V.EuclidDistance_Assign_Module(192, 1024, soft_assign_alpha=16.0).to(dev)
xn = torch.randn(1, 1, 1, 32768, 192, device=dev, requires_grad=True)


def native():
    for p in mn.parameters(): p.grad = None
    xn.grad = None
    D, A, S, R, F, lab = mn(xn)
    (mn.fused_cluster_loss() + V.e4_norm(R, xn.detach())).backward()
    return [D.detach(), A.detach(), R.detach(), lab, xn.grad, mn.cluster_center.grad]


hammer("native head C=192 K=1024 fwd+bwd N=32768", native)

seq = torch.nn.Sequential(torch.nn.Conv3d(96, 192, (1, 2, 2), stride=(1, 2, 2)), torch.nn.GELU()).to(dev)
xe = torch.randn(4, 96, 4, 56, 56, device=dev, requires_grad=True)
ge = torch.randn(4, 4, 28, 28, 192, device=dev)


def tail():
    xe.grad = None
    for p in seq.parameters(): p.grad = None
    y = V.downsample_gelu_tokens(xe, seq[0], seq[1])
    y.backward(ge)
    return [y.detach(), xe.grad, seq[0].weight.grad, seq[0].bias.grad]


hammer("encoder tail fwd+bwd 12544 tokens", tail)
print("STRESS", "OK" if bad == 0 else f"FAILED ({bad} differing repeats)")
sys.exit(0 if bad == 0 else 1)

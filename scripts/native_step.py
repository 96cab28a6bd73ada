"""reference-native head shapes (model/backbone.py:40-41): cluster1 K=1024, C=192; space_cluster K=128, P=28*28 —
forward / backward times of the two heads on N = B*D*28*28 tokens"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoad_b200 as V
dev = torch.device("cuda", 0)
B, D, H, W, C = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 4, 28, 28, 192
torch.manual_seed(0)
c1 = V.EuclidDistance_Assign_Module(C, 1024, soft_assign_alpha=16.0).to(dev)
sp = V.Space_EuclidDistance_Assign_Module(C, 128, space_size=28, soft_assign_alpha=32.0).to(dev)
x = torch.randn(B, D, H, W, C, device=dev, requires_grad=True)
gR = torch.randn(B, D, H, W, C, device=dev) * 1e-3
def timed(fn, n=6):
    ms = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    return statistics.median(ms), out
t_f1, o1 = timed(lambda: c1(x))
def bwd1():
    for p in list(c1.parameters()) + [x]: p.grad = None
    Dm, A, S, R, F, lab = c1(x)
    torch.autograd.backward([c1.fused_cluster_loss(), R], [None, gR])
t_s1, _ = timed(bwd1)
t_f2, o2 = timed(lambda: sp(x))
def bwd2():
    for p in list(sp.parameters()) + [x]: p.grad = None
    Ds, As, S, _ = sp(x)
    torch.norm(Ds * As).backward()
t_s2, _ = timed(bwd2)
print(f"tokens {B*D*H*W}: cluster1 fwd {t_f1:.3f} ms, fwd+bwd {t_s1:.3f} ms | space fwd {t_f2:.3f} ms, fwd+bwd {t_s2:.3f} ms")

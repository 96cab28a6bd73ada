"""the non-headline rows once each (for ncu): pixel loss + frame scoring streams, space head fwd/bwd at the cfg2 batch,
memory forward (m=2000, d=768), cluster forward at C=768, K=256"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoad_b200 as V
dev = torch.device("cuda", 0)
torch.manual_seed(0)
r = torch.rand(16, 3, 16, 256, 256, device=dev); c = torch.rand(16, 3, 16, 256, 256, device=dev)
for _ in range(3):
    V.e4_norm(r, c); V.frame_mse(r, c, want_psnr=True); V.l1_mean(r, c)
sp = V.Space_EuclidDistance_Assign_Module(192, 128, space_size=32).to(dev)
xs = torch.randn(64, 8, 32, 32, 192, device=dev, requires_grad=True)
for _ in range(2):
    sp(xs); sp.fused_cluster_loss().backward()
mem = V.Memory(2000, 768, 768, 0.1, 0.1)
q = torch.randn(2, 768, 32, 32, device=dev)
keys = torch.nn.functional.normalize(torch.rand(2000, 768, device=dev), dim=1)
for _ in range(2):
    mem(q, keys, train=True)
m = V.EuclidDistance_Assign_Module(768, 256, soft_assign_alpha=16.0).to(dev)
x = torch.randn(1, 1, 1, 65536, 768, device=dev)
for _ in range(2):
    m(x)
torch.cuda.synchronize()
print("ok")

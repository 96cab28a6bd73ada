"""space head (C3) forward + backward at a given M = B*D (for ncu launch lists / timing)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, statistics
import videoad_b200 as V
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
torch.manual_seed(0)
C, K, S = 192, 128, 32
m = V.Space_EuclidDistance_Assign_Module(C, K, space_size=S, soft_assign_alpha=32.0).to(dev)
x = torch.randn(B, 8, S, S, C, device=dev, requires_grad=True)
fw, bw = [], []
for i in range(iters):
    for p in m.parameters(): p.grad = None
    x.grad = None
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    Ds, As, Sd, _ = m(x)
    loss = torch.norm(Ds * As)
    e[1].record()
    loss.backward()
    e[2].record(); torch.cuda.synchronize()
    fw.append(e[0].elapsed_time(e[1])); bw.append(e[1].elapsed_time(e[2]))
print(f"M={B*8} fwd ms median {statistics.median(fw):.3f}  bwd ms median {statistics.median(bw):.3f}")

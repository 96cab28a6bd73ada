"""space head (C3) forward + backward from the fused space loss at the cfg2 batch (M = 512): median ms over a few
iterations, or (argument 1) a single warm iteration for an ncu launch list"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import statistics
import torch
import videoad_b200 as V
single = len(sys.argv) > 1
dev = torch.device("cuda", 0)
torch.manual_seed(0)
C, K, S = 192, 128, 32
m = V.Space_EuclidDistance_Assign_Module(C, K, space_size=S, soft_assign_alpha=32.0).to(dev)
x = torch.randn(64, 8, S, S, C, device=dev, requires_grad=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fw, bw = [], []
for i in range(2 if single else 8):
    for p in m.parameters(): p.grad = None
    x.grad = None
    flush.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    m(x)
    loss = m.fused_cluster_loss()
    e[1].record()
    loss.backward()
    e[2].record(); torch.cuda.synchronize()
    fw.append(e[0].elapsed_time(e[1])); bw.append(e[1].elapsed_time(e[2]))
print(f"M=512 fwd ms median {statistics.median(fw[1:]):.3f}  bwd ms median {statistics.median(bw[1:]):.3f}  "
      f"sum {statistics.median(fw[1:]) + statistics.median(bw[1:]):.3f}")

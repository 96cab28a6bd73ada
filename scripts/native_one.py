"""the reference-native cluster head (C = 192, K = 1024) forward + backward, explicit torch.norm(D * A) loss: for ncu launch lists"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoad_b200 as V
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = V.EuclidDistance_Assign_Module(192, 1024, soft_assign_alpha=16.0).to(dev)
x = torch.randn(1, 1, 1, 65536, 192, device=dev, requires_grad=True)
for i in range(2):
    for p in m.parameters(): p.grad = None
    x.grad = None
    D, A, S, R, F, _ = m(x)
    loss = (m.fused_cluster_loss() if len(sys.argv) > 1 else torch.norm(D * A)) + V.e4_norm(R, x.detach())
    loss.backward()
torch.cuda.synchronize()

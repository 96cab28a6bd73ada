"""run the cfg2 cluster forward+backward a few times (for ncu captures / timing of the backward)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, statistics
import videoad_b200 as V
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
torch.manual_seed(0)
C, K = 192, 32
m = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=16.0).to(dev)
x = torch.randn(64, 8, 32, 32, C, device=dev, requires_grad=True)
gR = torch.randn(64, 8, 32, 32, C, device=dev) * 1e-3
ms = []
for i in range(iters):
    for p in m.parameters(): p.grad = None
    x.grad = None
    D, A, S, R, F, lab = m(x)
    loss = torch.sqrt(m.loss_sq[0])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); torch.autograd.backward([loss, R], [None, gR]); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
print("bwd ms:", [round(v, 3) for v in ms], "median", statistics.median(ms))

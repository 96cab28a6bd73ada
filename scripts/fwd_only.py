"""run the cfg2 cluster forward a few times (for ncu captures / quick timing)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, statistics
import videoad_b200 as V
impl = {"auto": V.IMPL_AUTO, "simt": V.IMPL_SIMT, "tc": V.IMPL_TCGEN05}[sys.argv[1] if len(sys.argv) > 1 else "auto"]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
C = int(sys.argv[3]) if len(sys.argv) > 3 else 192
K = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=16.0).to(dev)
m.impl = impl
x = torch.randn(64, 8, 32, 32, C, device=dev)
ms = []
with torch.no_grad():
    for i in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = m(x); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
print("fwd ms:", [round(v, 3) for v in ms], "median", statistics.median(ms))

"""generic-shape cluster forward (row kernels + tcgen05 GEMMs) a few times: python scripts/fwd_generic.py C K N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoad_b200 as V
C, K, n = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda", 0)
m = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=16.0).to(dev)
x = torch.randn(1, 1, 1, n, C, device=dev)
with torch.no_grad():
    for _ in range(3):
        m(x)
torch.cuda.synchronize()

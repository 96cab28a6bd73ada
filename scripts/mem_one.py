"""one memory eval forward at N = 65536 (m = 2000, d = 768) after a warm-up: for ncu launch lists"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoad_b200 as V
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mem = V.Memory(2000, 768, 768, 0.1, 0.1)
keys = torch.nn.functional.normalize(torch.rand(2000, 768, device=dev), dim=1)
q = torch.randn(64, 768, 32, 32, device=dev)
with torch.no_grad():
    for i in range(2):
        mem(q, keys, train=(len(sys.argv) > 1))
torch.cuda.synchronize()

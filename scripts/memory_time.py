"""memory module eval forward (m = 2000, d = 768) at N = 2048 / 8192 / 65536 tokens: median ms"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import statistics
import torch
import videoad_b200 as V
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mem = V.Memory(2000, 768, 768, 0.1, 0.1)
keys = torch.nn.functional.normalize(torch.rand(2000, 768, device=dev), dim=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n, (b, h) in ((2048, (2, 32)), (8192, (8, 32)), (65536, (64, 32))):
    q = torch.randn(b, 768, h, h, device=dev)
    ts = []
    with torch.no_grad():
        for i in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mem(q, keys, train=False); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    print(f"N={n} eval ms median {statistics.median(ts[1:]):.3f}")

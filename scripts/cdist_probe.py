"""batched cdist (tcgen05 path vs SIMT path vs float64)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import videoad_b200 as V
dev = torch.device("cuda", 0)
for (nb, R, P, C) in [(192, 128, 128, 784), (64, 128, 128, 1024), (8, 128, 128, 64), (3, 200, 136, 128), (40, 256, 256, 512)]:
    torch.manual_seed(1)
    a = torch.rand(nb, R, C, device=dev); b = torch.rand(nb, P, C, device=dev)
    ref = torch.cdist(a.double(), b.double())
    os.environ.pop("VADC_NO_TC_GEMM", None)
    o1 = V.cdist(a, b).double()
    os.environ["VADC_NO_TC_GEMM"] = "1"
    o2 = V.cdist(a, b).double()
    os.environ.pop("VADC_NO_TC_GEMM", None)
    e1 = (o1 - ref).abs(); e2 = (o2 - ref).abs()
    bad = (e1 > 1e-3).nonzero()
    print((nb, R, P, C), "tc max err", float(e1.max()), "simt max err", float(e2.max()), "n bad", bad.shape[0],
          "first bad", bad[:3].tolist(), "batches with bad", sorted(set(bad[:, 0].tolist()))[:10])

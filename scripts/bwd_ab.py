"""kernel-only time of the cfg2 cluster backward (library events right around the kernel) + parity of its outputs against
the mma.sync fused kernel, for the library selected by VADC_LIB_PATH:  python scripts/bwd_ab.py [iters]"""
import ctypes, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoad_b200 as V
from videoad_b200 import _lib
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)
torch.manual_seed(0)
C, K = 192, 32
m = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=16.0).to(dev)
with torch.no_grad():
    m.norm.weight.copy_(1 + 0.2 * torch.randn(C, device=dev)); m.norm.bias.copy_(0.1 * torch.randn(C, device=dev))
x = (torch.randn(64, 8, 32, 32, C, device=dev) * 1.3 + 0.2).requires_grad_(True)
gR = torch.randn(64, 8, 32, 32, C, device=dev) * 1e-3
lib = _lib.lib()


def run(n):
    for _ in range(n):
        for p in m.parameters():
            p.grad = None
        x.grad = None
        D, A, S, R, F, lab = m(x)
        torch.autograd.backward([torch.sqrt(m.loss_sq[0]), R], [None, gR])
    torch.cuda.synchronize()
    return x.grad.clone(), m.cluster_center.grad.clone(), m.norm.weight.grad.clone(), m.norm.bias.grad.clone()


os.environ["VADC_BWD_IMPL"] = "fused"; lib.vadc_refresh_env()
ref = run(1)
os.environ["VADC_BWD_IMPL"] = os.environ.get("AB_IMPL", "tc"); lib.vadc_refresh_env()
run(3)
lib.vadc_timing_enable(1)
got = run(iters)
ms, cnt = ctypes.c_float(0), ctypes.c_int(0)
lib.vadc_timing_read(1, ctypes.byref(ms), ctypes.byref(cnt))
lib.vadc_timing_enable(0)
errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(got, ref)]
alg = 524288 * (12 * C + 8 * K) + 4 * K * C
print(f"{os.environ.get('VADC_LIB_PATH', 'default'):60s} impl={os.environ['VADC_BWD_IMPL']:4s} bwd kernel {ms.value * 1e3:7.1f} us over {cnt.value} launches"
      f"  = {alg / ms.value / 1e6 / 6457.4:.3f} of HBM roofline   max rel err vs fused: {max(errs):.2e}")

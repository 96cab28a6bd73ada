"""quick probe of the tcgen05 fused cluster forward vs the fp64 oracle (dev helper)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import videoad_b200 as V
from oracle import np_oracle as O
from gpu_util import T, N as to_np, rel, make_cluster_module

for (Ntok, C, K, alpha) in [(128, 192, 32, 16.0), (4096, 192, 32, 16.0), (77, 192, 32, 16.0), (1000, 192, 64, 16.0), (300, 256, 32, 8.0), (20000, 128, 32, 16.0)]:
    rng = np.random.default_rng(Ntok)
    x = (rng.standard_normal((1, 1, 1, Ntok, C)) * 1.5 + 0.3).astype(np.float32)
    cen = rng.random((K, C)).astype(np.float32)
    w = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    m = make_cluster_module(V, C, K, alpha, cen, w, b, V.IMPL_TCGEN05)
    try:
        with torch.no_grad():
            D, A, S, R, F, lab = m(T(x))
        torch.cuda.synchronize()
    except RuntimeError as e:
        print(Ntok, C, K, "ERR", e); continue
    o = O.cluster_forward(x, cen, w, b, alpha, dtype=np.float64)
    lo = float(O.frobenius_loss(o["D"], o["A"], np.float64))
    mism = int((to_np(lab) != o["label"]).sum())
    print(f"N={Ntok} C={C} K={K}: F {rel(to_np(F), o['feature']):.2e} D {rel(to_np(D), o['D']):.2e} "
          f"A {np.abs(to_np(A) - o['A']).max():.2e} R {rel(to_np(R), o['x_rec']):.2e} "
          f"loss {abs(float(m.fused_cluster_loss()) - lo) / lo:.2e} label mism {mism}", flush=True)

"""tcgen05 backward vs the mma.sync fused backward vs the fp64 oracle (training-graph case: gR + fused loss)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import videoad_b200 as V
from oracle import np_oracle as O
dev = torch.device("cuda", 0)

def run(impl, x, gR, m):
    os.environ["VADC_BWD_IMPL"] = impl
    for p in m.parameters(): p.grad = None
    xt = x.clone().requires_grad_(True)
    D, A, S, R, F, lab = m(xt.view(1, 1, 1, *x.shape))
    loss = m.fused_cluster_loss() * 1.3
    torch.autograd.backward([loss, R], [None, gR.view(1, 1, 1, *gR.shape)])
    torch.cuda.synchronize()
    return [t.detach().cpu().numpy().astype(np.float64) for t in (xt.grad, m.cluster_center.grad, m.norm.weight.grad, m.norm.bias.grad)]

for (N, C, K, alpha) in [(64, 192, 32, 16.0), (1500, 192, 32, 16.0), (777, 128, 32, 32.0), (20000, 64, 32, 8.0), (148 * 64 * 3 + 5, 192, 32, 16.0)]:
    torch.manual_seed(N)
    m = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=alpha).to(dev)
    with torch.no_grad():
        m.norm.weight.copy_(1 + 0.2 * torch.randn(C, device=dev)); m.norm.bias.copy_(0.1 * torch.randn(C, device=dev))
    x = torch.randn(N, C, device=dev) * 2 + 0.3
    gR = torch.randn(N, C, device=dev) * 1e-2
    a = run("tc", x, gR, m)
    b = run("fused", x, gR, m)
    names = ["gx", "gcen", "gw", "gb"]
    msg = [f"{n} {np.abs(p - q).max() / np.abs(q).max():.1e}" for n, p, q in zip(names, a, b)]
    line = f"N={N} C={C}: tc vs fused: " + ", ".join(msg)
    if N <= 2000:
        xn = x.cpu().numpy(); cen = m.cluster_center.detach().cpu().numpy()
        w, bb = m.norm.weight.detach().cpu().numpy(), m.norm.bias.detach().cpu().numpy()
        f = O.cluster_forward(xn.reshape(1, 1, 1, N, C), cen, w, bb, alpha, dtype=np.float64)
        D64, A64 = f["D"].reshape(N, K), f["A"].reshape(N, K)
        gl = 1.3 / np.sqrt(((D64 * A64) ** 2).sum())
        ref = O.cluster_backward(xn, cen, w, bb, alpha, gD=gl * D64 * A64 * A64, gA=gl * D64 * D64 * A64,
                                 gR=gR.cpu().numpy(), dtype=np.float64)
        line += " | vs fp64 oracle: " + ", ".join(f"{n} tc {np.abs(p - r).max() / np.abs(r).max():.1e} fused {np.abs(q - r).max() / np.abs(r).max():.1e}"
                                                   for n, p, q, r in zip(names, a, b, ref))
    print(line, flush=True)

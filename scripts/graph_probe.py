"""capture one cluster-head training step in a CUDA graph (debugging aid for bench.py's replay path)"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import videoad_b200 as V
dev = torch.device("cuda", 0)
C, K = 192, 32
mod = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=16.0).to(dev)
x_buf = torch.randn(int(os.environ.get("B", "4")), 8, 32, 32, C, device=dev)
gR = torch.randn_like(x_buf) * 1e-3
params = [mod.cluster_center, mod.norm.weight, mod.norm.bias]
stage = sys.argv[1] if len(sys.argv) > 1 else "full"

def step():
    for p in params: p.grad = None
    x = x_buf.detach().requires_grad_(True)
    D, A, S, R, F, lab = mod(x)
    if stage == "fwd": return
    loss = V.global_frobenius(mod.loss_sq)
    if stage == "loss": return
    torch.autograd.backward([loss, R], [None, gR])

side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        step()
    g.replay(); torch.cuda.synchronize()
    print(stage, "capture ok", float(mod.cluster_center.grad.abs().sum()) if stage == "full" else "")
except Exception:
    traceback.print_exc()

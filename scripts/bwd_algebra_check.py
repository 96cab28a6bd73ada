"""numpy model of the tcgen05 backward's arithmetic (bf16 two-term operand splits, algebraic LayerNorm
row statistics, T via the distance identity) against the fp64 oracle"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import np_oracle as O

def bf(a):
    return torch.tensor(np.asarray(a, np.float32)).to(torch.bfloat16).float().numpy()
def split(a, terms=2):
    out, r = [], np.asarray(a, np.float32).copy()
    for _ in range(terms):
        h = bf(r); out.append(h); r = (r - h).astype(np.float32)
    return out
def mm3(a, b):   # hi*hi + hi*lo + lo*hi with fp32 accumulation (model: fp64 then round)
    return (a[0].astype(np.float64) @ b[0] + a[0].astype(np.float64) @ b[1] + a[1].astype(np.float64) @ b[0]).astype(np.float32)

rng = np.random.default_rng(0)
N, C, K, alpha = 4096, 192, 32, 16.0
x = rng.standard_normal((N, C)).astype(np.float32) * 2 + 0.5
cen = rng.random((K, C)).astype(np.float32)
gam = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32); bet = (0.1 * rng.standard_normal(C)).astype(np.float32)
gR = (rng.standard_normal((N, C)) * 1e-3).astype(np.float32)
f64 = O.cluster_forward(x.reshape(1, 1, 1, N, C), cen, gam, bet, alpha, dtype=np.float64)
D64, A64 = f64["D"].reshape(N, K), f64["A"].reshape(N, K)
L = np.sqrt(((D64 * A64) ** 2).sum()); gLsq = 1.0 / (2 * L)          # d sqrt(loss_sq) / d loss_sq
gDl, gAl = 2 * gLsq * D64 * A64 * A64, 2 * gLsq * D64 * D64 * A64
ref = O.cluster_backward(x, cen, gam, bet, alpha, gD=gDl, gA=gAl, gR=gR, dtype=np.float64)

# ---- model (fp32 + bf16x2 operands)
f32 = O.cluster_forward(x.reshape(1, 1, 1, N, C), cen, gam, bet, alpha, dtype=np.float32)
D, A, mu, rstd = f32["D"].reshape(N, K), f32["A"].reshape(N, K), f32["mu"], f32["rstd"]
sc = np.float32(2 * gLsq)
xh = ((x - mu[:, None]) * rstd[:, None]).astype(np.float32)
X = split(xh); G = split(gR); Cn = split(cen)
xr = (X[0] + X[1]).astype(np.float32)                   # what E3 reads back from shared memory
z = xh * gam + bet
w = gam * gam * xh
q1, q2, q3 = w.sum(1), (w * xh).sum(1), (bet * gam * xh).sum(1)
p1 = q1 + (bet * gam).sum(); p2 = q2 + q3; zz = q2 + 2 * q3 + (bet * bet).sum()
cc = (cen * cen).sum(1); cg = cen @ gam; bc = cen @ bet
G1 = mm3(G, [c.T for c in Cn])
gA = G1 + sc * D * D * A
dot = (gA * A).sum(1, keepdims=True)
gd = sc * D * A * A - alpha * A * (gA - dot)
r = np.where(D == 0, 0, gd / D).astype(np.float32)
rsum = r.sum(1)
T = 0.5 * (zz[:, None] + cc[None] - D * D) - bc[None]
s1 = (rsum * p1 - r @ cg) / C
s2 = (rsum * p2 - (r * T).sum(1)) / C
R_ = split(r); A_ = split(A)
acc = mm3(R_, Cn)
zr = xr * gam + bet
gz = zr * rsum[:, None] - acc
gx = ((gz * gam - s1[:, None] - xr * s2[:, None]) * rstd[:, None]).astype(np.float32)
P1 = mm3([g.T for g in G], A_)            # [C,K]
P2 = mm3([t.T for t in X], R_)            # [C,K]
rcol = r.sum(0)
gc = P1.T - gam[None] * P2.T - bet[None] * rcol[:, None] + cen * rcol[:, None]
gb = gz.sum(0); gw = (gz * xr).sum(0)
for name, got, want in (("gx", gx, ref[0]), ("gcen", gc, ref[1]), ("g_ln_w", gw, ref[2]), ("g_ln_b", gb, ref[3])):
    print(f"{name:7s} max err / max |ref| = {np.abs(got - want).max() / np.abs(want).max():.2e}")
# ---- algebraic g_ln_w / g_ln_b from P2, rcol and Q (no per-token column sums of gz needed)
Q = (xr * xr * rsum[:, None]).sum(0)
P2s = P2.sum(1)
gb_alg = gam * P2s + bet * rcol.sum() - rcol @ cen
gw_alg = gam * Q + bet * P2s - (cen.T * P2).sum(1)
for name, got, want in (("g_ln_w (algebraic)", gw_alg, ref[2]), ("g_ln_b (algebraic)", gb_alg, ref[3])):
    print(f"{name:20s} max err / max |ref| = {np.abs(got - want).max() / np.abs(want).max():.2e}")

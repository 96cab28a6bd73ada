"""C3 (space head) and M1-M5 (memory) parity on the GPU through the C ABI."""
import numpy as np
import pytest
import torch

import videoad_b200 as V
from oracle import np_oracle as O
from conftest import load_golden
from gpu_util import T, N, rel, dev, assert_selfdist_close

pytestmark = pytest.mark.gpu


def make_space(C, K, side, alpha, centers, w, b):
    m = V.Space_EuclidDistance_Assign_Module(C, K, space_size=side, soft_assign_alpha=alpha).to(dev())
    with torch.no_grad():
        m.cluster_center.copy_(T(centers)); m.norm.weight.copy_(T(w)); m.norm.bias.copy_(T(b))
    return m


@pytest.mark.parametrize("name", ["space_c8_k6_p16", "space_c16_k40_p36"])
def test_space_golden(name):
    g = load_golden(name)
    C, K, P = g["centers"].shape            # (space_c8_k6_p16: cluster_num = 6 — padded to 8 inside the wrapper)
    side = int(round(P ** 0.5))
    m = make_space(C, K, side, float(g["alpha"]), g["centers"], g["ln_w"], g["ln_b"])
    x = T(g["x"], grad=True)
    Ds, As, S, rec = m(x)
    assert rec == [] and Ds.shape == g["D"].shape
    assert rel(N(Ds), g["D"]) < 1e-5
    np.testing.assert_allclose(N(As), g["A"], rtol=1e-3, atol=2e-6)
    assert_selfdist_close(N(S), g["S"])
    loss = torch.norm(Ds * As)
    assert abs(float(loss) - float(g["space_loss"])) < 1e-4 * float(g["space_loss"])
    assert abs(float(m.fused_cluster_loss()) - float(g["space_loss"])) < 1e-4 * float(g["space_loss"])
    loss.backward()
    assert rel(N(x.grad), g["gx"]) < 3e-4
    assert rel(N(m.cluster_center.grad), g["gcenters"]) < 3e-4
    assert rel(N(m.norm.weight.grad), g["g_ln_w"]) < 3e-4
    assert rel(N(m.norm.bias.grad), g["g_ln_b"]) < 3e-4


@pytest.mark.parametrize("B,Dd,side,C,K", [(2, 4, 28, 192, 128), (1, 3, 8, 24, 8), (2, 2, 32, 64, 16),
                                            (4, 8, 28, 192, 128), (16, 8, 16, 192, 64),    # these two: batched tcgen05 GEMMs
                                            (1, 3, 7, 24, 8),        # odd P: token count not a multiple of 4
                                            (8, 16, 12, 192, 128),   # tcgen05, P = 144: a frame's last 64-token chunk is ragged
                                            (32, 16, 4, 256, 128),   # tcgen05, P = 16 < one chunk, C at the tiled kernels' limit
                                            (16, 16, 8, 288, 64)])   # tcgen05, C > 256: the untiled LayerNorm kernels
def test_space_vs_oracle(B, Dd, side, C, K):
    rng = np.random.default_rng(B * 100 + side)
    x = (rng.standard_normal((B, Dd, side, side, C)) * 1.3).astype(np.float32)
    cen = rng.random((C, K, side * side)).astype(np.float32)
    w = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    m = make_space(C, K, side, 32.0, cen, w, b)
    xt = T(x, grad=True)
    Ds, As, S, rec = m(xt)
    o = O.space_cluster_forward(x, cen, w, b, 32.0, dtype=np.float64)
    assert rel(N(Ds), o["D"]) < 1e-5
    np.testing.assert_allclose(N(As), o["A"], rtol=3e-3, atol=1e-6)
    assert_selfdist_close(N(S), O.cdist_mm(cen, cen))
    m.fused_cluster_loss().backward()
    gD, gA = O.frobenius_loss_grads(o["D"], o["A"], 1.0, np.float64)
    gx, gc, gw, gb = O.space_cluster_backward(x, cen, w, b, 32.0, gD=gD, gA=gA, dtype=np.float64)
    assert rel(N(xt.grad), gx) < 3e-4
    assert rel(N(m.cluster_center.grad), gc) < 3e-4
    assert rel(N(m.norm.weight.grad), gw) < 3e-4
    assert rel(N(m.norm.bias.grad), gb) < 3e-4


@pytest.mark.parametrize("B,Dd,side,C,K", [(1, 2, 4, 8, 6), (8, 8, 16, 192, 64)])     # SIMT path with a padded K; tcgen05 path
def test_space_self_distance_output_and_gradient(B, Dd, side, C, K):
    """cluster_dist (model/cluster.py:134) comes out of the same call as the distances; a gradient through it reaches
    the centroids with autograd's cdist formula (the reference's own loss never uses it: backbone.py:95-97)"""
    rng = np.random.default_rng(K)
    cen = rng.random((C, K, side * side)).astype(np.float32)
    m = make_space(C, K, side, 32.0, cen, np.ones(C, np.float32), np.zeros(C, np.float32))
    x = T(rng.standard_normal((B, Dd, side, side, C)).astype(np.float32))
    _, _, S, _ = m(x)
    assert S.shape == (C, K, K)
    assert_selfdist_close(N(S), O.cdist_mm(cen, cen))
    assert_selfdist_close(N(m.self_similarity()), O.cdist_mm(cen, cen))
    wgt = T(rng.random((C, K, K)).astype(np.float32))
    (S * wgt).sum().backward()
    cr = torch.tensor(cen, dtype=torch.float64, requires_grad=True)
    Sr = torch.cdist(cr, cr, compute_mode="donot_use_mm_for_euclid_dist")
    (Sr * wgt.double().cpu()).sum().backward()
    assert rel(N(m.cluster_center.grad), cr.grad.numpy()) < 2e-4


def test_three_term_modes_of_the_tensor_core_paths():
    """the tcgen05 contractions of the space head and the memory module default to two fp16 terms of power-of-two-scaled
    operands (22 significant bits); VADC_SPACE_TERMS=3 / VADC_MEMORY_TERMS=3 select the fp32-faithful bf16 x3 split"""
    import os
    from videoad_b200 import _lib
    os.environ["VADC_SPACE_TERMS"] = "3"; os.environ["VADC_MEMORY_TERMS"] = "3"
    _lib.lib().vadc_refresh_env()
    try:
        test_space_vs_oracle(4, 8, 28, 192, 128)
        test_memory_vs_oracle(2, 768, 32, 32, 2000)
    finally:
        os.environ.pop("VADC_SPACE_TERMS", None); os.environ.pop("VADC_MEMORY_TERMS", None)
        _lib.lib().vadc_refresh_env()


def test_soft_assign_modules():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 7, 13)).astype(np.float32)
    for dims in (1, -1, 0):
        xm = np.moveaxis(x, dims, -1)
        neg = V.NegSoftAssign(dims, 4.0)(T(x))
        pos = V.PosSoftAssign(dims, 2.0)(T(x))
        np.testing.assert_allclose(np.moveaxis(N(neg), dims, -1), O.neg_soft_assign(xm, 4.0), rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(np.moveaxis(N(pos), dims, -1), O.pos_soft_assign(xm, 2.0), rtol=1e-5, atol=1e-7)
    xt = T(x, grad=True)
    y = V.NegSoftAssign(-1, 3.0)(xt)
    (y * T(x) ** 2).sum().backward()
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    yr = torch.softmax(-3.0 * xr, -1)
    (yr * xr.detach() ** 2).sum().backward()
    assert rel(N(xt.grad), xr.grad.numpy()) < 1e-5


# ---------------------------------------------------------------------------
# Memory
# ---------------------------------------------------------------------------
def check_memory(query, keys, tol=2e-5):
    mem = V.Memory(keys.shape[0], keys.shape[1], keys.shape[1], 0.1, 0.1)
    o = O.memory_forward(query, keys, train=True, dtype=np.float64)
    uq, um, sq, sm, gl, sl = mem(T(query), T(keys), train=True)
    assert uq.shape == o["updated_query"].shape and not uq.is_contiguous()    # permuted view (Memory.py:259)
    assert rel(N(uq), o["updated_query"]) < tol
    assert rel(N(sq), o["score_query"]) < tol
    assert rel(N(sm), o["score_memory"]) < tol
    assert abs(float(gl) - o["gathering_loss"]) < tol * abs(o["gathering_loss"])
    assert abs(float(sl) - o["spreading_loss"]) < tol * abs(o["spreading_loss"])
    assert rel(N(um), o["updated_memory"]) < tol
    t = mem(T(query), T(keys), train=False)
    assert len(t) == 5 and t[1].data_ptr() == T(keys).data_ptr() or torch.equal(t[1], T(keys))
    assert rel(N(t[0]), o["updated_query"]) < tol
    return mem, o


@pytest.mark.parametrize("name", ["memory_d32_m10", "memory_d64_m50"])
def test_memory_golden(name):
    g = load_golden(name)
    mem = V.Memory(g["keys"].shape[0], g["keys"].shape[1], g["keys"].shape[1], 0.1, 0.1)
    uq, um, sq, sm, gl, sl = mem(T(g["query"]), T(g["keys"]), train=True)
    assert rel(N(uq), g["updated_query"]) < 2e-5
    assert rel(N(um), g["updated_memory"]) < 2e-5
    assert rel(N(sq), g["score_query"]) < 2e-5
    assert rel(N(sm), g["score_memory"]) < 2e-5
    assert abs(float(gl) - float(g["gathering_loss"])) < 2e-5 * float(g["gathering_loss"])
    assert abs(float(sl) - float(g["spreading_loss"])) < 2e-5 * float(g["spreading_loss"])
    uq_t, um_t, _, _, gl_t = mem(T(g["query"]), T(g["keys"]), train=False)
    assert rel(N(uq_t), g["test_updated_query"]) < 2e-5
    assert np.array_equal(N(um_t), g["test_updated_memory"])
    assert abs(float(V.MemoryLoss(T(g["keys"]))) - float(g["separateness"])) < 2e-5 * float(g["separateness"])


@pytest.mark.parametrize("B,d,h,w,m", [(2, 768, 32, 32, 2000), (1, 96, 7, 5, 33), (3, 64, 8, 8, 1)])
def test_memory_vs_oracle(B, d, h, w, m):
    """cfg3: m=2000, d=768, N=2048 plus ragged / single-slot edge cases"""
    rng = np.random.default_rng(m + d)
    query = rng.standard_normal((B, d, h, w)).astype(np.float32)
    keys = O.l2_normalize(rng.random((m, d)).astype(np.float32), 1)
    if m == 1:
        mem = V.Memory(1, d, d, 0.1, 0.1)
        with pytest.raises(RuntimeError):
            mem(T(query), T(keys), train=True)               # torch.topk(..., 2) fails on m = 1
        o = O.memory_forward(query, np.concatenate([keys, keys]), train=False, dtype=np.float64)
        out = mem(T(query), T(keys), train=False)
        assert rel(N(out[3]), np.ones((B * h * w, 1))) < 1e-6
        return
    check_memory(query, keys)


def test_memory_public_methods():
    rng = np.random.default_rng(11)
    query = rng.standard_normal((2, 48, 6, 6)).astype(np.float32)
    keys = O.l2_normalize(rng.random((20, 48)).astype(np.float32), 1)
    mem, o = check_memory(query, keys)
    q4 = mem.prepare_query(T(query))                        # [B,h,w,d]
    assert rel(N(q4).reshape(-1, 48), o["q"]) < 1e-6
    sq, sm = mem.get_score(T(keys), q4)
    assert rel(N(sq), o["score_query"]) < 2e-5 and rel(N(sm), o["score_memory"]) < 2e-5
    uq, sq2, sm2 = mem.read(q4, T(keys))
    assert rel(N(uq), o["updated_query"]) < 2e-5
    assert abs(float(mem.gather_loss(q4, T(keys), True)) - o["gathering_loss"]) < 2e-5 * o["gathering_loss"]
    assert abs(float(mem.spread_loss(q4, T(keys), True)) - o["spreading_loss"]) < 2e-5 * o["spreading_loss"]
    assert rel(N(mem.update(q4, T(keys), True)), o["updated_memory"]) < 2e-5
    top1 = T(o["top1"], dtype=torch.int64)[:, None]
    qu = mem.get_update_query(T(keys), top1, None, sq, q4.reshape(-1, 48), True)
    assert rel(N(qu), o["query_update"]) < 2e-5


def test_memory_update_is_deterministic():
    rng = np.random.default_rng(5)
    query = rng.standard_normal((4, 64, 16, 16)).astype(np.float32)
    keys = O.l2_normalize(rng.random((7, 64)).astype(np.float32), 1)      # few slots -> long segments
    mem = V.Memory(7, 64, 64, 0.1, 0.1)
    a = mem(T(query), T(keys), train=True)[1]
    b = mem(T(query), T(keys), train=True)[1]
    assert torch.equal(a, b)


# ---------------------------------------------------------------------------
# Memory: gradient with respect to the query (autograd of Memory.py:145-175)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["memory_d32_m10", "memory_d64_m50"])
def test_memory_backward_golden(name):
    """d query of <updated_query, W> + 0.7 gathering + 0.3 spreading against the reference module's own
    autograd (fixture written by tests/golden/make_golden.py); tolerance 2e-5 of the largest entry"""
    g = load_golden(name)
    mem = V.Memory(g["keys"].shape[0], g["keys"].shape[1], g["keys"].shape[1], 0.1, 0.1)
    W = T(g["g_updated_query"])
    q = T(g["query"]).requires_grad_(True)
    o = mem(q, T(g["keys"]), train=True)
    ((o[0] * W).sum() + 0.7 * o[4] + 0.3 * o[5]).backward()
    assert rel(N(q.grad), g["g_query_train"]) < 2e-5
    q = T(g["query"]).requires_grad_(True)
    o = mem(q, T(g["keys"]), train=False)
    ((o[0] * W).sum() + 0.7 * o[4]).backward()
    assert rel(N(q.grad), g["g_query_test"]) < 2e-5
    assert not o[1].requires_grad and not o[2].requires_grad and not o[3].requires_grad


@pytest.mark.parametrize("B,d,h,w,m", [(2, 768, 32, 32, 2000), (1, 96, 7, 5, 33), (3, 64, 8, 8, 2)])
def test_memory_backward_vs_oracle(B, d, h, w, m):
    """cfg3 shape (m=2000, d=768, N=2048) and ragged shapes against the fp64 oracle; each loss term alone too"""
    rng = np.random.default_rng(B * 100 + m)
    keys = rng.random((m, d)).astype(np.float32)
    keys /= np.linalg.norm(keys, axis=1, keepdims=True)
    query = rng.standard_normal((B, d, h, w)).astype(np.float32)
    W = rng.standard_normal((B, 2 * d, h, w)).astype(np.float32)
    mem = V.Memory(m, d, d, 0.1, 0.1)
    f = O.memory_forward(query, keys, train=True, dtype=np.float64)
    for cu, cg, cs in [(1.0, 0.7, 0.3), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (1.0, 0.0, 0.0)]:
        q = T(query).requires_grad_(True)
        o = mem(q, T(keys), train=True)
        loss = 0
        if cu: loss = loss + cu * (o[0] * T(W)).sum()
        if cg: loss = loss + cg * o[4]
        if cs: loss = loss + cs * o[5]
        loss.backward()
        ref = O.memory_query_backward(query, keys, f["top1"], f["top2"], W * cu if cu else None,
                                      cg if cg else None, cs if cs else None)
        assert rel(N(q.grad), ref) < 2e-5, (cu, cg, cs)
    # an unused forward leaves no gradient and costs no backward launch
    q = T(query).requires_grad_(True)
    o = mem(q, T(keys), train=True)
    (o[2].sum() * 0 + 1).backward() if o[2].requires_grad else None
    assert q.grad is None


# ---------------------------------------------------------------------------
# Memory under data parallelism: two "ranks" on one GPU = the full batch
# ---------------------------------------------------------------------------
class _TwoRanks:
    """two python threads stand in for two ranks: a reduction hands each thread the other's tensor (the role of the
    NCCL all-reduce); a lock lets only one of them enqueue work at a time (they share one stream and workspace)"""

    def __init__(self):
        import threading
        self.lock, self.bar, self.slots = threading.Lock(), threading.Barrier(2), [None, None]

    def reducer(self, rank):
        def reduce(t, op):
            self.slots[rank] = t
            self.lock.release()
            self.bar.wait()
            other = self.slots[1 - rank]
            self.bar.wait()
            self.lock.acquire()
            return torch.maximum(t, other) if op == "max" else t + other
        return reduce


@pytest.mark.parametrize("d,h,w,m", [(64, 6, 5, 17), (768, 16, 16, 2000)])
def test_memory_global_batch_two_shards(d, h, w, m):
    """Memory.global_batch: the batch split over two ranks reproduces the single-process full-batch forward
    (column softmax over ALL tokens, update sums, loss means) and its gradient — SURVEY 8e 'Memory under DP'"""
    import threading
    rng = np.random.default_rng(m)
    keys = rng.random((m, d)).astype(np.float32)
    keys /= np.linalg.norm(keys, axis=1, keepdims=True)
    query = rng.standard_normal((2, d, h, w)).astype(np.float32)
    query[1] *= 1.7                                        # different column maxima on the two shards
    W = rng.standard_normal((2, 2 * d, h, w)).astype(np.float32)
    full = V.Memory(m, d, d, 0.1, 0.1)
    qf = T(query).requires_grad_(True)
    of = full(qf, T(keys), train=True)
    ((of[0] * T(W)).sum() + 0.7 * of[4] + 0.3 * of[5]).backward()

    ranks, out = _TwoRanks(), [None, None]

    def run(rank):
        ranks.lock.acquire()
        try:
            mem = V.Memory(m, d, d, 0.1, 0.1)
            mem.global_batch, mem._reduce = True, ranks.reducer(rank)
            q = T(query[rank:rank + 1]).requires_grad_(True)
            o = mem(q, T(keys), train=True)
            ((o[0] * T(W[rank:rank + 1])).sum() + 0.7 * o[4] + 0.3 * o[5]).backward()
            out[rank] = (o, q.grad)
        finally:
            ranks.lock.release()

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert out[0] is not None and out[1] is not None
    n = h * w
    for r in range(2):
        o, g = out[r]
        assert rel(N(o[1]), N(of[1])) < 2e-5                                   # updated_memory: identical on both ranks
        assert rel(N(o[2]), N(of[2])[r * n:(r + 1) * n]) < 2e-5                # score_query: softmax over ALL tokens
        assert rel(N(o[3]), N(of[3])[r * n:(r + 1) * n]) < 2e-5
        assert abs(float(o[4]) - float(of[4])) < 2e-5 * abs(float(of[4]))      # global means
        assert abs(float(o[5]) - float(of[5])) < 2e-5 * abs(float(of[5]))
        assert rel(N(g), N(qf.grad)[r:r + 1]) < 2e-5


def test_memory_full_size_properties():
    """cfg3 at a full clip batch: N = 65536 tokens (B=64, 32x32), m = 2000, d = 768 — [N,m] score matrices of 524 MB.
    The 2e11-MAC contractions are checked on a row sample against float64; everything downstream of the returned
    scores (top-1 / top-2, losses, segmented update, normalisation) is recomputed on the host for ALL tokens."""
    B, d, h, w, m = 64, 768, 32, 32, 2000
    g = torch.Generator(device="cuda").manual_seed(5)
    query = torch.randn(B, d, h, w, device=dev(), generator=g)
    keys = torch.nn.functional.normalize(torch.rand(m, d, device=dev(), generator=g), dim=1)
    mem = V.Memory(m, d, d, 0.1, 0.1)
    uq, um, sq, sm, gl, sl = mem(query, keys, train=True)
    Ntok = B * h * w
    assert sq.shape == (Ntok, m) and sm.shape == (Ntok, m) and uq.shape == (B, 2 * d, h, w)
    # softmax normalisations (float64 sums on the device tensors, read back as two small vectors)
    assert float((sm.double().sum(1) - 1).abs().max()) < 1e-5
    assert float((sq.double().sum(0) - 1).abs().max()) < 1e-4
    q = O.memory_prepare_query(N(query), np.float64)                          # [N,d] host, float64
    kn = N(keys).astype(np.float64)
    rows = np.random.default_rng(0).choice(Ntok, 192, replace=False)
    logit = q[rows] @ kn.T
    e = np.exp(logit - logit.max(1, keepdims=True))
    assert rel(N(sm[torch.as_tensor(rows, device=dev())]), e / e.sum(1, keepdims=True)) < 2e-5
    uqf = N(uq.permute(0, 2, 3, 1).reshape(Ntok, 2 * d)[torch.as_tensor(rows, device=dev())])
    assert rel(uqf[:, :d], q[rows]) < 2e-6
    assert rel(uqf[:, d:], (e / e.sum(1, keepdims=True)) @ kn) < 2e-5
    # downstream of the scores, all tokens
    smh, sqh = N(sm), N(sq)
    part = np.argpartition(-smh, 1, axis=1)[:, :2]
    first = smh[np.arange(Ntok), part[:, 0]] >= smh[np.arange(Ntok), part[:, 1]]
    g1 = np.where(first, part[:, 0], part[:, 1]); g2 = np.where(first, part[:, 1], part[:, 0])
    gather = ((q - kn[g1]) ** 2).mean()
    dap = np.sqrt(((q - kn[g1] + 1e-6) ** 2).sum(1)); dan = np.sqrt(((q - kn[g2] + 1e-6) ** 2).sum(1))
    spread = np.maximum(dap - dan + 1.0, 0).mean()
    assert abs(float(gl) - gather) < 2e-5 * gather and abs(float(sl) - spread) < 2e-5 * spread
    wgt = sqh[np.arange(Ntok), g1].astype(np.float64) / sqh.max(0).astype(np.float64)[g1]
    upd = np.zeros((m, d))
    np.add.at(upd, g1, wgt[:, None] * q)
    ref = upd + kn
    ref /= np.maximum(np.linalg.norm(ref, axis=1, keepdims=True), 1e-12)
    assert rel(N(um), ref) < 2e-5

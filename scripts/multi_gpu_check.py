"""Multi-GPU correctness of the data-parallel path on REAL GPUs (one process per GPU, NCCL + NVLink peer memory):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_check.py
Checks (every rank asserts, rank 0 prints one line per check):
  1. the one-shot peer all-reduce == NCCL all-reduce, eagerly and replayed from a CUDA graph, bit-identical across ranks
  2. N-rank cluster training step == single-process full batch (loss, parameter gradients, token gradients)
  3. Memory.global_batch over N ranks == single-process full batch (SURVEY 8e 'Memory under DP')
  4. sharded evaluate_videos == single-process evaluate_videos (SURVEY 8e 'scoring under DP')"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import videoad_b200 as V

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def say(*a):
    if rank == 0:
        print(*a, flush=True)


def _excepthook(tp, val, tb):
    import traceback
    msg = f"[rank {rank}] FAILED: " + "".join(traceback.format_exception(tp, val, tb)).strip()[-1500:]
    print(msg, flush=True)
    try:                                              # torchrun's summary drops the message: keep it where gpurun merges files back
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/multi_gpu_check_rank{rank}.err", "w") as fh:
            fh.write(msg + "\n")
    except OSError:
        pass
    sys.__excepthook__(tp, val, tb)


sys.excepthook = _excepthook


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# ---- 1. peer all-reduce
assert V.enable_peer_allreduce(1 << 14), V.peer_allreduce_status()
say("collectives:", V.peer_allreduce_status())
g = torch.Generator(device=dev).manual_seed(100 + rank)
for it in range(50):
    ts = [torch.randn(32, 192, device=dev, generator=g), torch.randn(192, device=dev, generator=g),
          torch.randn(192, device=dev, generator=g), torch.randn(1, device=dev, generator=g)]
    want = [t.clone() for t in ts]
    for w in want:
        dist.all_reduce(w)
    V.allreduce_sum_packed(ts)
    for a, b in zip(ts, want):
        # NCCL adds in its own order, ours in rank order: compare against the size of the summands, not of the sum (a sum of
        # `world` standard normals that nearly cancels — the one-element message — has no relative accuracy to speak of)
        err = float((a.double() - b.double()).abs().max()) / max(float(b.double().abs().max()), float(world) ** 0.5)
        assert err < 1e-5, (it, tuple(a.shape), err)
    chk = torch.stack([t.double().sum() for t in ts])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks disagree bitwise"
say("1a. peer all-reduce == NCCL over 50 eager calls; results bit-identical across ranks")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    buf = [torch.zeros(32, 192, device=dev), torch.zeros(1, device=dev)]
    src = [torch.randn(32, 192, device=dev, generator=g), torch.randn(1, device=dev, generator=g)]

    def step():
        for b_, s_ in zip(buf, src):
            b_.copy_(s_)
        V.allreduce_sum_packed(buf)
    for _ in range(3):
        step()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        step()
    for it in range(20):
        for s_ in src:
            s_.copy_(torch.randn(s_.shape, device=dev, generator=g))
        want = [s_.clone() for s_ in src]
        gr.replay()
        s.synchronize()
        for w in want:
            dist.all_reduce(w)
        for a, b in zip(buf, want):
            err = float((a.double() - b.double()).abs().max()) / max(float(b.double().abs().max()), float(world) ** 0.5)
            assert err < 1e-5, ("graph", it, tuple(a.shape), err)
    del gr
torch.cuda.synchronize()
say("1b. ... and replayed 20x from a CUDA graph with new inputs")

# ---- 1c. latency of the two per-step messages, peer kernel vs NCCL (CUDA events, 200 back-to-back calls)
def lat(fn):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1000 / 200


grads = [torch.randn(32, 192, device=dev), torch.randn(192, device=dev), torch.randn(192, device=dev)]
one = [torch.randn(1, device=dev)]
t_peer_g, t_peer_1 = lat(lambda: V.allreduce_sum_packed(grads)), lat(lambda: V.allreduce_sum_packed(one))
V.disable_peer_allreduce()
t_nccl_g, t_nccl_1 = lat(lambda: V.allreduce_sum_packed(grads)), lat(lambda: V.allreduce_sum_packed(one))
assert V.enable_peer_allreduce(1 << 14)
say("1c. us per all-reduce (eager, incl. launch): 25 KB gradients peer %.1f vs NCCL+pack/unpack %.1f; scalar peer %.1f vs NCCL %.1f"
    % (t_peer_g, t_nccl_g, t_peer_1, t_nccl_1))

# ---- 2. cluster training step: N ranks vs one process
C, K, B = 192, 32, 4
torch.manual_seed(7)
mod = V.EuclidDistance_Assign_Module(C, K, soft_assign_alpha=16.0).to(dev)
for p in mod.parameters():
    dist.broadcast(p.data, 0)
gen = torch.Generator(device=dev).manual_seed(5)
x_all = torch.randn(world * B, 4, 16, 16, C, device=dev, generator=gen)        # same on every rank (same seed)
g_all = torch.randn(world * B, 4, 16, 16, C, device=dev, generator=gen) * 1e-2
params = [mod.cluster_center, mod.norm.weight, mod.norm.bias]


def run(x, gR, collectives):
    for p in params:
        p.grad = None
    x = x.detach().requires_grad_(True)
    D, A, S, R, F, lab = mod(x)
    loss = V.global_frobenius(mod.loss_sq, ddp_compat=not collectives)
    torch.autograd.backward([loss, R], [None, gR])
    if collectives:
        V.allreduce_sum_packed([p.grad for p in params])
    return [loss.detach().clone()] + [p.grad.clone() for p in params] + [x.grad.clone()]


lo, hi = V.shard_range(world * B)
got = run(x_all[lo:hi], g_all[lo:hi], True)
want = run(x_all, g_all, False)
errs = [rel(a, b) for a, b in zip(got[:4], want[:4])] + [rel(got[4], want[4][lo:hi])]
assert max(errs) < 1e-4, errs                     # fp32 sums over N x more tokens in another order
say("2. %d-rank training step == single-process full batch: max rel err %.2e" % (world, max(errs)))

# ---- 3. memory with global-batch semantics
m, d = 200, 64
mem = V.Memory(m, d, d, 0.1, 0.1)
keys = torch.nn.functional.normalize(torch.rand(m, d, device=dev, generator=gen), dim=1)
q_all = torch.randn(world * 2, d, 8, 8, device=dev, generator=gen)
ref = mem(q_all, keys, train=True)
mem.global_batch = True
lo, hi = V.shard_range(world * 2)
out = mem(q_all[lo:hi], keys, train=True)
e_um = rel(out[1], ref[1]); e_gl = rel(out[4].reshape(1), ref[4].reshape(1)); e_sl = rel(out[5].reshape(1), ref[5].reshape(1))
e_uq = rel(out[0], ref[0][lo:hi])
N_loc = (hi - lo) * 64
e_sq = rel(out[2], ref[2][lo * 64: lo * 64 + N_loc])
assert max(e_um, e_gl, e_sl, e_uq, e_sq) < 1e-4, (e_um, e_gl, e_sl, e_uq, e_sq)
say("3. Memory.global_batch over %d ranks == full batch: updated_memory %.1e, losses %.1e / %.1e, score_query %.1e" % (world, e_um, e_gl, e_sl, e_sq))

# ---- 4. sharded scoring
rng = np.random.default_rng(3)
lengths = [25, 33, 41, 17, 29, 37, 24, 40]
videos = [torch.rand(3, T, 16, 16, generator=torch.Generator().manual_seed(T)) for T in lengths]
labels = [(rng.random(T) < 0.4).astype(np.int64) for T in lengths]
for l in labels:
    l[0], l[1] = 0, 1
scenes = ["%02d" % (i % 3) for i in range(len(lengths))]
model = lambda c: c + 0.05 * torch.sin(37.0 * c) * (1.0 + c)      # noqa: E731
a1, p1, s1, l1 = V.evaluate_videos(model, videos, labels, scenes, 4, 1)
a2, p2, s2, l2 = V.evaluate_videos(model, videos, labels, scenes, 4, 1, shard=True)
assert a1 == a2 and p1 == p2 and all(np.array_equal(x, y) for x, y in zip(s1, s2))
say("4. sharded evaluate_videos == single process: AUC %.6f on every rank" % a2)

V.disable_peer_allreduce()
dist.barrier()
dist.destroy_process_group()
say("MULTI-GPU CHECK OK")

"""Host-side mirror of the MNAD memory module (model/Memory.py:62-261) on top
of libvadc.so.  Same class / method names and return tuples; ``keys`` in,
``updated_memory`` out (stateless, detached) like the reference.

What changes underneath: ``get_score`` is computed ONCE per
forward instead of four times, and the python loop over memory slots with a
``nonzero()`` host sync per slot (Memory.py:100-113) is a stable counting sort
plus a segmented weighted sum on the device.
"""
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, f32c, ptr, stream, workspace


def MemoryLoss(memory):
    """model/Memory.py:52-59: separateness sum |K K^T/2 + 1/2 - I| / (m (m-1))"""
    _lib.require_cuda(memory)
    k = f32c(memory)
    m, d = k.shape
    out = torch.empty((1,), device=k.device, dtype=torch.float32)
    l = _lib.lib()
    ws = workspace(l.vadc_memory_separateness_workspace_bytes(m, d), k.device)
    check(l.vadc_memory_separateness(ptr(k), m, d, ptr(out), ptr(ws), ws.numel(), stream()),
          "vadc_memory_separateness")
    return out[0]


def _dist_reduce(t, op):
    """all-reduce over the default process group (NCCL over NVLink); identity when not distributed"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return t


class _Scores:
    """everything ``get_score`` + the topk calls derive from one (query, keys) pair"""
    __slots__ = ("q", "score_query", "score_memory", "colmax", "colsum", "top1", "top2", "score_memory_terms")


def _compute_scores(q, keys, need_query_softmax=True, want_read_terms=False):
    N, d = q.shape
    m = keys.shape[0]
    dev = q.device
    s = _Scores()
    s.q = q
    s.score_memory = torch.empty((N, m), device=dev, dtype=torch.float32)
    s.score_query = torch.empty((N, m), device=dev, dtype=torch.float32) if need_query_softmax else None
    s.colmax = torch.empty((m,), device=dev, dtype=torch.float32)
    s.colsum = torch.empty((m,), device=dev, dtype=torch.float32)
    s.top1 = torch.empty((N,), device=dev, dtype=torch.int64)
    s.top2 = torch.empty((N,), device=dev, dtype=torch.int64)
    # operand terms of score_memory for the read contraction, written in the softmax pass (opaque: 4 bytes per element)
    s.score_memory_terms = torch.empty((N * m * 4 + 16,), device=dev, dtype=torch.uint8) if want_read_terms else None
    l = _lib.lib()
    ws = workspace(l.vadc_memory_score_workspace_bytes(N, m, d), dev)
    check(l.vadc_memory_score(ptr(q), ptr(keys), N, m, d, ptr(s.score_query), ptr(s.score_memory),
                              ptr(s.colmax), ptr(s.colsum), ptr(s.top1), ptr(s.top2), ptr(s.score_memory_terms),
                              ptr(ws), ws.numel(), stream()), "vadc_memory_score")
    return s


class _MemoryForward(torch.autograd.Function):
    """Memory.forward (Memory.py:145-175) as one autograd node.  Gradients flow to ``query`` exactly where
    the reference graph lets them: through the first half of ``updated_query`` (the read softmax is
    ``.detach()``ed, :255), ``gathering_loss`` (:245) and ``spreading_loss`` (:229), back through
    ``F.normalize`` (:148).  ``keys`` are constants (every use is detached or, for the read product, the
    caller passes a plain tensor: main.py:131); ``updated_memory`` is detached (:204); the two score
    matrices are returned for inspection and are not differentiable here."""

    @staticmethod
    def forward(ctx, mod, query, keys, train):
        keys_c = f32c(keys)
        qc = f32c(query)
        q4 = mod.prepare_query(qc)
        q, shp = mod._flat(q4)
        s = _compute_scores(q, keys_c, want_read_terms=True)
        gathering_loss, spreading_loss = mod._losses(s, keys_c, train)
        gathering_loss = gathering_loss.clone()           # separate outputs, not two views of one buffer
        spreading_loss = None if spreading_loss is None else spreading_loss.clone()
        updated_query = mod._read(s, keys_c, shp)
        ctx.loss_scale = 1.0
        glob = mod.global_batch and mod._reduce is not None
        if glob:
            # tokens are sharded over ranks: column statistics, losses and update sums become those of the full batch
            n_all = mod._reduce(torch.tensor([float(q.shape[0])], device=q.device), "sum")
            ctx.loss_scale = float(q.shape[0]) / float(n_all)         # local mean -> this rank's share of the global mean
            cm_g, contrib, S = mod._dp_column_stats(s)
            mod._dp_rescale_scores(s, contrib, S)
            gathering_loss = mod._reduce(gathering_loss.reshape(1) * ctx.loss_scale, "sum")[0]
            if spreading_loss is not None:
                spreading_loss = mod._reduce(spreading_loss.reshape(1) * ctx.loss_scale, "sum")[0]
        ctx.save_for_backward(qc, keys_c, s.top1, s.top2)
        ctx.train = bool(train)
        ctx.set_materialize_grads(False)
        if train:
            updated_memory = mod._dp_update(s, keys_c, cm_g) if glob else mod._update(s, keys_c)
            ctx.mark_non_differentiable(updated_memory, s.score_query, s.score_memory)
            return (updated_query, updated_memory, s.score_query, s.score_memory,
                    gathering_loss, spreading_loss)
        ctx.mark_non_differentiable(s.score_query, s.score_memory)
        return updated_query, s.score_query, s.score_memory, gathering_loss

    @staticmethod
    def backward(ctx, *grads):
        qc, keys_c, top1, top2 = ctx.saved_tensors
        if ctx.train:
            g_uq, _, _, _, g_gather, g_spread = grads
        else:
            (g_uq, _, _, g_gather), g_spread = grads, None
        if g_uq is None and g_gather is None and g_spread is None:
            return None, None, None, None
        B, d, h, w = qc.shape
        g_uq = None if g_uq is None else f32c(g_uq)                  # [B,2d,h,w]; the first d channels reach q
        # (global batch: the kernel divides by the local token count; loss_scale turns that into the global mean)
        g_gather = None if g_gather is None else f32c(g_gather).reshape(1) * ctx.loss_scale
        g_spread = None if g_spread is None else f32c(g_spread).reshape(1) * ctx.loss_scale
        gq = torch.empty_like(qc)
        check(_lib.lib().vadc_memory_query_bwd(ptr(qc), ptr(keys_c), ptr(top1), ptr(top2) if ctx.train else None,
                                               ptr(g_uq), ptr(g_gather), ptr(g_spread), B, d, h * w,
                                               keys_c.shape[0], ptr(gq), stream()), "vadc_memory_query_bwd")
        return None, gq, None, None


class Memory(nn.Module):
    """Drop-in for model/Memory.py:62-261.  ``forward`` is differentiable with respect to ``query``
    (``_MemoryForward``); the stand-alone public methods (``read``, ``gather_loss``, ...) are
    forward-only.  The reference module is not wired into ``Mymodel``."""

    def __init__(self, memory_size, feature_dim, key_dim, temp_update, temp_gather):
        super().__init__()
        self.memory_size = memory_size
        self.feature_dim = feature_dim
        self.key_dim = key_dim
        self.temp_update = temp_update
        self.temp_gather = temp_gather
        # data parallel, SURVEY 8e: False = every rank treats its shard as the batch (what the reference's DDP run
        # does); True = column softmax, update sums and loss means over the tokens of ALL ranks, i.e. the
        # single-process full-batch result: all-reduce MAX of the column maxima [m], SUM of the column sums [m],
        # SUM of the update sums [m,d] and of three scalars
        self.global_batch = False
        self._reduce = _dist_reduce

    # -- helpers ------------------------------------------------------------
    @staticmethod
    def _flat(query):
        """[B,h,w,d] (the layout every public method of the reference takes) -> [N,d]"""
        _lib.require_cuda(query)
        B, h, w, d = query.shape
        return f32c(query).reshape(B * h * w, d), (B, h, w, d)

    @staticmethod
    def prepare_query(query):
        """F.normalize(query, dim=1) + permute(0,2,3,1)  (Memory.py:148-149):
        [B,d,h,w] -> [B,h,w,d] contiguous"""
        _lib.require_cuda(query)
        qc = f32c(query)
        B, d, h, w = qc.shape
        out = torch.empty((B, h, w, d), device=qc.device, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_memory_prepare_query_workspace_bytes(B, h * w), qc.device)
        check(l.vadc_memory_prepare_query(ptr(qc), B, d, h * w, ptr(out), ptr(ws), ws.numel(), stream()),
              "vadc_memory_prepare_query")
        return out

    # -- reference API --------------------------------------------------------
    def get_score(self, mem, query):
        """Memory.py:133-143 -> (softmax over tokens, softmax over slots), [N,m] each"""
        q, _ = self._flat(query)
        s = _compute_scores(q, f32c(mem))
        return s.score_query, s.score_memory

    def forward(self, query, keys, train=True):
        """Memory.py:145-175.  query [B,d,h,w], keys [m,d]."""
        _lib.require_cuda(query, keys)
        if train:
            return _MemoryForward.apply(self, query, keys, True)
        uq, sq, sm, gl = _MemoryForward.apply(self, query, keys, False)
        return uq, keys, sq, sm, gl

    def update(self, query, keys, train):
        """Memory.py:177-204"""
        with torch.no_grad():
            q, _ = self._flat(query)
            keys_c = f32c(keys)
            return self._update(_compute_scores(q, keys_c), keys_c)

    def get_update_query(self, mem, max_indices, update_indices, score, query, train):
        """Memory.py:94-131: weighted sum of the queries assigned to each slot"""
        with torch.no_grad():
            q = f32c(query)
            return self._segmented_update(q, f32c(mem), f32c(score),
                                          max_indices.reshape(-1).to(torch.int64).contiguous())[0]

    def pointwise_gather_loss(self, query_reshape, keys, gathering_indices, train):
        """Memory.py:206-212 (elementwise, not a hot path: plain indexing)"""
        return (query_reshape - keys[gathering_indices].squeeze(1).detach()) ** 2

    def spread_loss(self, query, keys, train):
        """Memory.py:214-231"""
        with torch.no_grad():
            q, _ = self._flat(query)
            keys_c = f32c(keys)
            return self._losses(_compute_scores(q, keys_c, False), keys_c, True)[1]

    def gather_loss(self, query, keys, train):
        """Memory.py:233-247"""
        with torch.no_grad():
            q, _ = self._flat(query)
            keys_c = f32c(keys)
            return self._losses(_compute_scores(q, keys_c, False), keys_c, False)[0]

    def read(self, query, updated_memory):
        """Memory.py:249-261"""
        with torch.no_grad():
            q, shp = self._flat(query)
            keys_c = f32c(updated_memory)
            s = _compute_scores(q, keys_c, want_read_terms=True)
            return self._read(s, keys_c, shp), s.score_query, s.score_memory

    # -- kernels --------------------------------------------------------------
    def _losses(self, s, keys, with_spread):
        N, d = s.q.shape
        m = keys.shape[0]
        if with_spread and m < 2:
            raise RuntimeError("selected index k out of range")      # torch.topk(…, 2) on m < 2
        out = torch.empty((2,), device=s.q.device, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_memory_losses_workspace_bytes(N, d), s.q.device)
        check(l.vadc_memory_losses(ptr(s.q), ptr(keys), ptr(s.top1), ptr(s.top2) if with_spread else None,
                                   N, m, d, ptr(out), ptr(ws), ws.numel(), stream()), "vadc_memory_losses")
        return out[0], (out[1] if with_spread else None)

    def _read(self, s, keys, shp):
        B, h, w, d = shp
        N, m = s.q.shape[0], keys.shape[0]
        uq = torch.empty((N, 2 * d), device=s.q.device, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_memory_read_workspace_bytes(N, m, d), uq.device)
        check(l.vadc_memory_read(ptr(s.q), ptr(s.score_memory), ptr(s.score_memory_terms), ptr(keys),
                                 N, m, d, ptr(uq), ptr(ws), ws.numel(), stream()), "vadc_memory_read")
        return uq.view(B, h, w, 2 * d).permute(0, 3, 1, 2)          # Memory.py:258-259

    def _segmented_update(self, q, keys, score_query, top1):
        N, d = q.shape
        m = keys.shape[0]
        qu = torch.empty((m, d), device=q.device, dtype=torch.float32)
        um = torch.empty((m, d), device=q.device, dtype=torch.float32)
        l = _lib.lib()
        ws = workspace(l.vadc_memory_update_workspace_bytes(N, m, d), q.device)
        check(l.vadc_memory_update(ptr(q), ptr(keys), ptr(score_query), ptr(top1), N, m, d, ptr(qu), ptr(um),
                                   ptr(ws), ws.numel(), stream()), "vadc_memory_update")
        return qu, um

    def _update(self, s, keys):
        return self._segmented_update(s.q, keys, s.score_query, s.top1)[1].detach()

    # -- global batch over ranks (include/vadc.h: vadc_memory_dp_*) -------------
    def _dp_column_stats(self, s):
        """(global column maxima, this rank's share of the column sums, the global column sums)"""
        l = _lib.lib()
        m = s.colmax.numel()
        cm_g = self._reduce(s.colmax.clone(), "max")
        contrib = torch.empty_like(s.colsum)
        check(l.vadc_memory_dp_contrib(ptr(s.colmax), ptr(s.colsum), ptr(cm_g), m, ptr(contrib), stream()),
              "vadc_memory_dp_contrib")
        S = self._reduce(contrib.clone(), "sum")
        return cm_g, contrib, S

    def _dp_rescale_scores(self, s, contrib, S):
        """local softmax over this rank's tokens -> softmax over all ranks' tokens, in place"""
        N, m = s.score_query.shape
        check(_lib.lib().vadc_scale_columns(ptr(s.score_query), ptr(contrib), ptr(S), N, m, stream()),
              "vadc_scale_columns")

    def _dp_update(self, s, keys, cm_g):
        """update sums of this rank (their weights are invariant under the rescaling above), scaled to the global
        column maximum, summed over ranks, then normalize(u + keys) — identical on every rank"""
        l = _lib.lib()
        m, d = keys.shape
        qu = self._segmented_update(s.q, keys, s.score_query, s.top1)[0]
        check(l.vadc_memory_dp_scale_update(ptr(qu), ptr(s.colmax), ptr(cm_g), m, d, stream()),
              "vadc_memory_dp_scale_update")
        qu = self._reduce(qu, "sum")
        um = torch.empty_like(qu)
        check(l.vadc_memory_finish_update(ptr(qu), ptr(keys), m, d, ptr(um), stream()), "vadc_memory_finish_update")
        return um.detach()

"""Evaluation scoring: per-frame reconstruction error on
the GPU, PSNR -> per-video regularity score -> per-scene AUC on the host.

Reference: tool/evaluate.py:166-224, tool/contrast_evaluae.py:229-299,
tool/predict_evaluae.py:228-275, misc/utils.py:124-135.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, f32c, ptr, stream, workspace


def _plane_view(t):
    """a [B,C,T,H,W] tensor whose H*W planes are contiguous is used as it is (strided batch / channel / frame axes: a
    clip batch cut out of a resident video, ``recon[:, :, :1]``); anything else is made contiguous"""
    t = t if t.dtype == torch.float32 else t.float()
    H, W = t.shape[-2:]
    ok = (W == 1 or t.stride(-1) == 1) and (H == 1 or t.stride(-2) == W) and all(st >= 0 for st in t.stride())
    if not ok or t.data_ptr() % 4:
        t = t.contiguous()
    return t


def frame_mse(recon, clip, want_psnr=False, out_mse=None, out_psnr=None):
    """per-frame MSE of a clip batch: recon, clip [B,C,D,H,W] -> [B,D] fp32
    (``MSELoss(none)`` -> 'B C D H W -> B D C H W' -> mean W, H, C;
    contrast_evaluae.py:232-235).  One pass over both tensors, nothing
    materialised; strided views (clips cut out of a resident video) are read in place.  With ``want_psnr`` also returns
    10 log10(1/mse) [B,D] float64 computed on the device (device-resident evaluation loop).
    ``out_mse`` / ``out_psnr``: optional contiguous [B*D] destinations (slices of a whole-run buffer)."""
    _lib.require_cuda(recon, clip)
    if recon.shape != clip.shape:
        raise RuntimeError(f"The size of tensor a {tuple(recon.shape)} must match the size of tensor b {tuple(clip.shape)}")
    r, c = _plane_view(recon), _plane_view(clip)
    B, Cc, T, H, W = r.shape
    mse = torch.empty((B, T), device=r.device, dtype=torch.float32) if out_mse is None else out_mse
    ps = None
    if want_psnr:
        ps = torch.empty((B, T), device=r.device, dtype=torch.float64) if out_psnr is None else out_psnr
    if B * T == 0:
        return (mse, ps) if want_psnr else mse
    l = _lib.lib()
    ws = workspace(l.vadc_frame_mse_workspace_bytes(B, T, H * W, Cc), r.device)
    check(l.vadc_frame_mse_strided(ptr(r), r.stride(0), r.stride(1), r.stride(2), ptr(c), c.stride(0), c.stride(1), c.stride(2),
                                   B, Cc, T, H * W, ptr(mse), ptr(ps), ptr(ws), ws.numel(), stream()),
          "vadc_frame_mse_strided")
    return (mse, ps) if want_psnr else mse


def clip_mse(recon, clip):
    """one score per clip (tool/predict_evaluae.py:228-234): mean over C,D,H,W"""
    B = recon.shape[0]
    r = f32c(recon).reshape(B, 1, 1, 1, -1)
    c = f32c(clip).reshape(B, 1, 1, 1, -1)
    return frame_mse(r, c)[:, 0]


def psnr(mse):
    """misc/utils.py:124-128 — list in, list out, Python float64."""
    return [10 * math.log10(1.0 / mse_item) for mse_item in mse]


def anomly_score(psnr):
    """misc/utils.py:131-135 — per-video 1 - minmax(psnr); raises
    ZeroDivisionError on a constant list exactly like the reference."""
    max_psnr = max(psnr)
    min_psnr = min(psnr)
    return [1.0 - (psnr_item - min_psnr) / (max_psnr - min_psnr) for psnr_item in psnr]


def minmax_score_device(psnr_dev, seg_offsets):
    """E3 on the device for the resident evaluation loop: psnr [total] float64,
    seg_offsets [n_videos+1] int64 -> score [total] float64."""
    _lib.require_cuda(psnr_dev, seg_offsets)
    p = psnr_dev.contiguous()
    off = seg_offsets.to(torch.int64).contiguous()
    out = torch.empty_like(p)
    check(_lib.lib().vadc_minmax_score(ptr(p), ptr(off), off.numel() - 1, ptr(out), stream()),
          "vadc_minmax_score")
    return out


def roc_auc_score(y_true, y_score):
    """Mann-Whitney AUC with mid-ranks: identical to sklearn.metrics.roc_auc_score
    for binary labels (tool/contrast_evaluae.py:278), including its ValueError
    when only one class is present."""
    y = np.asarray(y_true).astype(bool)
    s = np.asarray(y_score, dtype=np.float64)
    n_pos = int(y.sum())
    n_neg = int(y.size - n_pos)
    if n_pos == 0 or n_neg == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    order = np.argsort(s, kind="mergesort")
    ss = s[order]
    # mid-ranks of tied groups
    boundaries = np.flatnonzero(np.concatenate(([True], ss[1:] != ss[:-1], [True])))
    ranks_sorted = np.empty(s.size, np.float64)
    for a, b in zip(boundaries[:-1], boundaries[1:]):
        ranks_sorted[a:b] = 0.5 * (a + b - 1) + 1.0
    ranks = np.empty(s.size, np.float64)
    ranks[order] = ranks_sorted
    u = ranks[y].sum() - n_pos * (n_pos + 1) / 2.0
    return float(u / (n_pos * n_neg))


def regularity_auc(video_mse, video_labels, video_scene):
    """evaluation aggregate of tool/contrast_evaluae.py:262-299 (= evaluate.py:198-224):
    per video psnr -> anomly_score, concatenated per scene in first-seen order,
    AUC per scene, arithmetic mean over scenes.  Returns (auc, {scene: auc})."""
    scene_dict, scene_label = {}, {}
    for mse, lab, sc in zip(video_mse, video_labels, video_scene):
        mse = mse.tolist() if hasattr(mse, "tolist") else list(mse)
        predict_label = np.array(anomly_score(psnr(mse)))
        truth_label = np.asarray(lab.cpu() if isinstance(lab, torch.Tensor) else lab)
        assert len(predict_label) == len(truth_label)
        if sc in scene_dict:
            scene_dict[sc] = np.append(scene_dict[sc], predict_label)
            scene_label[sc] = np.append(scene_label[sc], truth_label)
        else:
            scene_dict[sc], scene_label[sc] = predict_label, truth_label
    per = {k: roc_auc_score(scene_label[k], scene_dict[k]) for k in scene_dict}
    return sum(per.values()) / len(per), per


def eval_clip_starts(n_frames, frame_num, batch_size):
    """clip schedule of tool/contrast_evaluae.py:185-203 (host integers only): non-overlapping clips, the
    reference's strict ``<`` bounds kept (the last clip is dropped when it would end exactly at T)."""
    batches, index = [], 0
    while index + frame_num < n_frames:
        starts = [index]
        for _ in range(batch_size - 1):
            if index + frame_num + 1 < n_frames:
                index = index + frame_num
                starts.append(index)
            else:
                break
        index = index + frame_num
        batches.append(starts)
    return batches


def eval_clip_starts_stride1(n_frames, frame_num, batch_size):
    """clip schedule of tool/predict_evaluae.py:185-203 and (batch 1) main_predict.py:401-404: clips one frame apart"""
    batches, index = [], 0
    while index + frame_num < n_frames:
        starts = [index]
        for _ in range(batch_size - 1):
            if index + frame_num + 1 < n_frames:
                index = index + 1
                starts.append(index)
            else:
                break
        index = index + 1
        batches.append(starts)
    return batches


def _clip_batch_view(v, start, nclips, step, frame_num):
    """[nclips, C, frame_num, H, W] view of the resident video ``v`` [C,T,H,W]: clip i = frames start + i*step ...;
    no copy (consecutive clips for step = frame_num, overlapping ones for step = 1)"""
    C, T, H, W = v.shape
    if start + (nclips - 1) * step + frame_num > T:
        # the reference concatenates a short tail clip onto full ones (contrast_evaluae.py:196): torch.cat's error
        raise RuntimeError("Sizes of tensors must match except in dimension 0. Expected size %d but got size %d for tensor "
                           "number 1 in the list." % (frame_num, T - (start + (nclips - 1) * step)))
    sc, st, sh, sw = v.stride()
    return v.as_strided((nclips, C, frame_num, H, W), (step * st, sc, st, sh, sw), v.storage_offset() + start * st)


def gather_video_scores(local):
    """§8e 'scoring under DP': every rank scored its own videos; rank 0 needs all of them in video order for the
    per-scene AUC.  ``local`` = list of (video_index, scores ndarray, labels ndarray).  Returns the merged, ordered list
    on every rank (the payload is ~40 k floats for the ShanghaiTech test set: one object all-gather)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return sorted(local, key=lambda t: t[0])
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    return sorted((item for part in parts for item in part), key=lambda t: t[0])


@torch.no_grad()
def evaluate_videos(model_fn, videos, labels, scenes, frame_num, batch_size, mode="contrast", ispredict=False,
                    shard=False):
    """Device-resident evaluation loop with the semantics of the reference's three evaluators:

    ``mode='contrast'``     tool/contrast_evaluae.py:170-300 (non-predict): consecutive clips, one score per FRAME
    ``mode='predict'``      tool/predict_evaluae.py:170-284: clips one frame apart, one score per CLIP; with
                            ``ispredict`` the model sees the clip's first four frames and its last frame is the target
    ``mode='first_frame'``  main_predict.py:389-457: clips one frame apart, the score is the error of the clip's first frame

    ``videos``: list of [C,T,H,W] tensors (host or device; a host video is copied to the GPU ONCE instead of
    once per clip), ``labels``: per-video [T] frame labels, ``scenes``: per-video scene id.
    ``model_fn(clips [B,C,D,H,W]) -> recon`` (for the reference's ``Mymodel`` pass ``lambda c: model(c)[0]``); the clip
    batch is a strided VIEW of the resident video (no stack / cat copy).  Per clip batch ONE fused reduction writes MSE
    and PSNR straight into the whole-run buffers on the device (no materialised loss tensor, no ``.tolist()`` sync per
    clip); per-video min-max normalisation runs on the device over all videos at once; only the final scores cross to
    the host for the per-scene AUC.  ``shard=True``: the videos are split over the ranks of the default process group
    (contiguous ``shard_range``), the scores are gathered and every rank returns the same result (§8e).
    A video whose PSNR is constant, or a frame reconstructed exactly (MSE 0), raises ZeroDivisionError like
    misc/utils.py:128,135.  Returns (auc, {scene: auc}, [per-video score arrays], [per-video label arrays])."""
    if mode not in ("contrast", "predict", "first_frame"):
        raise ValueError(f"unknown evaluation mode {mode!r}")
    labels = [np.asarray(l.cpu() if isinstance(l, torch.Tensor) else l).reshape(-1) for l in labels]
    n_videos = len(labels)
    lo, hi = 0, n_videos
    if shard:
        from .distributed import shard_range
        lo, hi = shard_range(n_videos)
    step = frame_num if mode == "contrast" else 1
    sched = eval_clip_starts if mode == "contrast" else eval_clip_starts_stride1
    per_clip = frame_num if mode == "contrast" else 1
    # the schedule of every video of this rank (host integers; a video has as many frames as labels) -> buffer sizes
    plans, total = {}, 0
    for i in range(lo, hi):
        batches = sched(len(labels[i]), frame_num, batch_size)
        n = sum(len(b) for b in batches) * per_clip
        plans[i] = (batches, total, n)
        total += n
    local = []
    if total > 0:
        mse_all = ps_all = None
        seg, order = [0], []
        for i, vid in enumerate(videos):
            if i >= hi:
                break
            if i < lo:
                continue
            batches, off, n = plans[i]
            lab = labels[i]
            if vid.shape[1] != len(lab):
                raise ValueError(f"video {i}: {vid.shape[1]} frames but {len(lab)} labels")
            v = vid if vid.is_cuda else vid.cuda(non_blocking=True)                # resident once, not once per clip
            if not (v.dtype == torch.float32 and v.stride(3) == 1 and v.stride(2) == v.shape[3] and v.data_ptr() % 16 == 0
                    and v.stride(0) % 4 == 0 and v.stride(1) % 4 == 0):
                v = f32c(v)                         # (a slice [:, :T] of a longer recording keeps its strides: no copy)
            if mse_all is None:
                mse_all = torch.empty((total,), device=v.device, dtype=torch.float32)
                ps_all = torch.empty((total,), device=v.device, dtype=torch.float64)
                mse_ptr, ps_ptr = mse_all.data_ptr(), ps_all.data_ptr()
            Cv, Tv, Hv, Wv = v.shape
            st_c, st_t = v.stride(0), v.stride(1)
            vptr, voff = v.data_ptr(), 0
            fast = None
            if mode == "contrast" and v.stride(3) == 1 and v.stride(2) == Wv and v.device.index == torch.cuda.current_device():
                l = _lib.lib()
                ws = workspace(l.vadc_frame_mse_workspace_bytes(batch_size, frame_num, Hv * Wv, Cv), v.device)
                ws_ptr, ws_bytes = ctypes.c_void_p(ws.data_ptr()), ws.numel()
                fast = l.vadc_frame_mse_strided.raw
                sptr = ctypes.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)      # once per video, not per batch
                HWv = Hv * Wv
            labs = []
            for starts in batches:
                nb = len(starts)
                clip = _clip_batch_view(v, starts[0], nb, step, frame_num)             # [B,C,D,H,W] view, no copy
                if mode == "contrast":
                    k = nb * frame_num
                    recon = model_fn(clip)
                    if fast is not None and recon.dtype == torch.float32 and recon.shape == clip.shape and recon.is_cuda:
                        # lean launch path (318 clip batches for the ShanghaiTech test set: the host must issue a batch in
                        # less time than the GPU needs to reduce it, ~75 us): raw entry point, pointers by arithmetic
                        r = recon if (recon.stride(-1) == 1 and recon.stride(-2) == Wv and recon.data_ptr() % 4 == 0) else recon.contiguous()
                        rc = fast(ctypes.c_void_p(r.data_ptr()), r.stride(0), r.stride(1), r.stride(2),
                                  ctypes.c_void_p(vptr + 4 * (voff + starts[0] * st_t)), step * st_t, st_c, st_t,
                                  nb, Cv, frame_num, HWv, ctypes.c_void_p(mse_ptr + 4 * off), ctypes.c_void_p(ps_ptr + 8 * off),
                                  ws_ptr, ws_bytes, sptr)
                        if rc != 0:
                            check(rc, "vadc_frame_mse_strided")
                    else:
                        frame_mse(recon, clip, want_psnr=True, out_mse=mse_all[off:off + k], out_psnr=ps_all[off:off + k])
                    labs.append(starts)                                                  # frame labels gathered once per video below
                else:
                    k = nb
                    if mode == "first_frame":
                        recon, target = model_fn(clip)[:, :, :1], clip[:, :, :1]
                    elif ispredict:
                        recon, target = model_fn(clip[:, :, 0:4]), clip[:, :, -1:]
                    else:
                        recon, target = model_fn(clip), clip
                    m = frame_mse(recon, target)                                       # [B, D']: mean over C, H, W
                    cm = m[:, 0] if m.shape[1] == 1 else m.mean(dim=1)                 # ... and over D' (predict_evaluae.py:233)
                    mse_all[off:off + k] = cm
                    ps_all[off:off + k] = 10.0 * torch.log10(1.0 / mse_all[off:off + k].double())
                    labs.append(lab[np.asarray(starts) + (frame_num if (mode == "first_frame" or ispredict) else 0)])
                off += k
            seg.append(seg[-1] + n)
            if mode == "contrast":
                st_all = np.fromiter((s0 for b_ in labs for s0 in b_), dtype=np.int64)
                lv = np.asarray(lab)[(st_all[:, None] + np.arange(frame_num)[None, :]).ravel()] if len(st_all) else np.zeros(0, np.asarray(lab).dtype)
            else:
                lv = np.concatenate(labs) if labs else np.zeros(0, lab.dtype)
            order.append((i, lv))
        score = minmax_score_device(ps_all, torch.tensor(seg, device=ps_all.device, dtype=torch.int64)).cpu().numpy()
        for (i, lv), a, b in zip(order, seg[:-1], seg[1:]):
            if b > a and not np.isfinite(score[a:b]).all():
                # constant PSNR (max == min) or an exactly reconstructed frame (mse == 0): misc/utils.py:128,135 divide by zero
                raise ZeroDivisionError("float division by zero")
            local.append((i, score[a:b], lv))
    merged = gather_video_scores(local) if shard else local
    if not merged or sum(len(m[1]) for m in merged) == 0:
        raise ValueError("no clips to evaluate")
    scene_dict, scene_label = {}, {}
    for (vi, sv, lv) in merged:
        sc = scenes[vi]
        assert len(sv) == len(lv)
        if sc in scene_dict:
            scene_dict[sc] = np.append(scene_dict[sc], sv)
            scene_label[sc] = np.append(scene_label[sc], lv)
        else:
            scene_dict[sc], scene_label[sc] = sv, lv
    per = {k: roc_auc_score(scene_label[k], scene_dict[k]) for k in scene_dict}
    return sum(per.values()) / len(per), per, [m[1] for m in merged], [m[2] for m in merged]

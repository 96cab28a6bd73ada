"""Evaluation scoring: per-frame reconstruction error on
the GPU, PSNR -> per-video regularity score -> per-scene AUC on the host.

Reference: tool/evaluate.py:166-224, tool/contrast_evaluae.py:229-299,
tool/predict_evaluae.py:228-275, misc/utils.py:124-135.
"""
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, f32c, ptr, stream, workspace


def frame_mse(recon, clip, want_psnr=False):
    """per-frame MSE of a clip batch: recon, clip [B,C,D,H,W] -> [B,D] fp32
    (``MSELoss(none)`` -> 'B C D H W -> B D C H W' -> mean W, H, C;
    contrast_evaluae.py:232-235).  One pass over both tensors, nothing
    materialised.  With ``want_psnr`` also returns 10 log10(1/mse) [B,D] float64
    computed on the device (device-resident evaluation loop)."""
    _lib.require_cuda(recon, clip)
    if recon.shape != clip.shape:
        raise RuntimeError(f"The size of tensor a {tuple(recon.shape)} must match the size of tensor b {tuple(clip.shape)}")
    r, c = f32c(recon), f32c(clip)
    B, Cc, T, H, W = r.shape
    mse = torch.empty((B, T), device=r.device, dtype=torch.float32)
    ps = torch.empty((B, T), device=r.device, dtype=torch.float64) if want_psnr else None
    l = _lib.lib()
    ws = workspace(l.vadc_frame_mse_workspace_bytes(B, T, H * W, Cc), r.device)
    check(l.vadc_frame_mse(ptr(r), ptr(c), B, Cc, T, H * W, ptr(mse), ptr(ps), ptr(ws), ws.numel(), stream()),
          "vadc_frame_mse")
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


@torch.no_grad()
def evaluate_videos(model_fn, videos, labels, scenes, frame_num, batch_size):
    """Device-resident evaluation loop with the semantics of
    tool/contrast_evaluae.py:170-300 (non-predict mode).

    ``videos``: list of [C,T,H,W] tensors (host or device; a host video is copied to the GPU ONCE instead of
    once per clip), ``labels``: per-video [T] frame labels, ``scenes``: per-video scene id.
    ``model_fn(clips [B,C,D,H,W]) -> recon [B,C,D,H,W]`` (for the reference's ``Mymodel`` pass
    ``lambda c: model(c)[0]``).  Per clip batch ONE fused kernel produces per-frame MSE and PSNR on the device
    (no materialised loss tensor, no ``.tolist()`` sync per clip); per-video min-max normalisation runs on the
    device over all videos at once; only the final scores cross to the host for the per-scene AUC.
    Returns (auc, {scene: auc}, [per-video score arrays], [per-video label arrays])."""
    dev = None
    ps_chunks, seg, lab_out = [], [0], []
    for vid, lab in zip(videos, labels):
        v = vid if vid.is_cuda else vid.cuda(non_blocking=True)
        dev = v.device
        lab = np.asarray(lab.cpu() if isinstance(lab, torch.Tensor) else lab).reshape(-1)
        n, labs = 0, []
        for starts in eval_clip_starts(v.shape[1], frame_num, batch_size):
            clip = torch.stack([v[:, s0:s0 + frame_num] for s0 in starts])       # [B,C,D,H,W]
            recon = model_fn(clip)
            _, ps = frame_mse(recon, clip, want_psnr=True)                       # [B,D] float64, on the device
            ps_chunks.append(ps.reshape(-1))
            n += ps.numel()
            for s0 in starts:
                labs.append(lab[s0:s0 + frame_num])
        seg.append(seg[-1] + n)
        lab_out.append(np.concatenate(labs) if labs else np.zeros(0, lab.dtype))
    if dev is None or seg[-1] == 0:
        raise ValueError("no clips to evaluate")
    psnr_all = torch.cat(ps_chunks)
    score = minmax_score_device(psnr_all, torch.tensor(seg, device=dev, dtype=torch.int64)).cpu().numpy()
    scores = [score[a:b] for a, b in zip(seg[:-1], seg[1:])]
    scene_dict, scene_label = {}, {}
    for sc, sv, lv in zip(scenes, scores, lab_out):
        assert len(sv) == len(lv)
        if sc in scene_dict:
            scene_dict[sc] = np.append(scene_dict[sc], sv)
            scene_label[sc] = np.append(scene_label[sc], lv)
        else:
            scene_dict[sc], scene_label[sc] = sv, lv
    per = {k: roc_auc_score(scene_label[k], scene_dict[k]) for k in scene_dict}
    return sum(per.values()) / len(per), per, scores, lab_out


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

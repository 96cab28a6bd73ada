"""torch-CPU port of the reference's op chain for the hot path.  TEST / BASELINE
INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): used by tests to cross-check
``np_oracle`` and by ``bench.py`` as the timed CPU baseline (``cpu_baseline`` and
``--impl reference``).  The reference is itself a PyTorch program whose CPU path
is exactly these ATen calls, but ``/root/reference`` does not travel to the GPU
box, so the op chain is restated here; each function cites the lines it follows
(paths relative to the reference root).  Checked against the reference-generated
fixtures in ``tests/test_oracle_port.py``.
"""
import math

import torch
import torch.nn.functional as F


def soft_assign_neg(d, alpha):
    """NegSoftAssign.forward, dims=-1 (model/cluster.py:48-55)"""
    d_min, _ = torch.min(d, -1, keepdim=True)
    e = torch.exp((-alpha) * (d - d_min))
    return e / e.sum(-1, keepdim=True)


def cluster_head(x, centers, ln_w, ln_b, alpha, eps=1e-5):
    """EuclidDistance_Assign_Module.forward (model/cluster.py:81-99)"""
    xn = F.layer_norm(x.clone(), (x.shape[-1],), ln_w, ln_b, eps)        # :83-84
    B, D, H, W, C = xn.shape
    x_re = xn.reshape(B, D * H * W, C)                                    # :86
    dist = torch.cdist(x_re, centers.unsqueeze(0))                        # :87
    label = torch.argmin(dist, dim=2, keepdim=True).reshape(-1)           # :88-89
    x_distance = dist.reshape(B, D, H, W, -1)                             # :91
    x_assign = soft_assign_neg(x_distance, alpha)                         # :92
    cluster_dist = torch.cdist(centers, centers)                          # :93
    x_rec = x_assign @ centers.clone()                                    # :94-95
    feature = x_re.reshape(-1, C)                                         # :96
    return x_distance, x_assign, cluster_dist, x_rec, feature, label


def space_head(x, centers, ln_w, ln_b, alpha, eps=1e-5):
    """Space_EuclidDistance_Assign_Module.forward (model/cluster.py:127-149)"""
    xn = F.layer_norm(x.clone(), (x.shape[-1],), ln_w, ln_b, eps)
    B, D, H, W, C = xn.shape
    x_re = xn.reshape(B * D, H * W, C).permute(2, 0, 1)                   # 'B D H W C -> C (B D) (H W)'
    dist = torch.cdist(x_re.contiguous(), centers)                        # :133
    x_distance = dist.reshape(C, B, D, -1).permute(1, 2, 0, 3)            # 'C (B D) CN -> B D C CN'
    x_assign = soft_assign_neg(x_distance, alpha)
    return x_distance, x_assign, torch.cdist(centers, centers), []


def cluster_train_step(x, centers, ln_w, ln_b, alpha, g_rec):
    """one training step of the cluster head as the reference runs it:
    forward (cluster.py:81-99), cluster loss torch.norm(D*A) (backbone.py:98),
    the decoder's upstream gradient on x_rec stood in by ``g_rec``, and
    ``loss.backward()`` (main_predict.py:296).  Returns loss and the four grads."""
    x = x.detach().requires_grad_(True)
    centers = centers.detach().requires_grad_(True)
    ln_w = ln_w.detach().requires_grad_(True)
    ln_b = ln_b.detach().requires_grad_(True)
    D, A, S, R, feat, label = cluster_head(x, centers, ln_w, ln_b, alpha)
    loss = torch.norm(D * A)
    (loss + (R * g_rec).sum()).backward()
    return loss.detach(), x.grad, centers.grad, ln_w.grad, ln_b.grad


def pixel_losses(recon, target):
    """main.py:191 and main_predict.py:273-275"""
    e = torch.nn.MSELoss(reduction='none')(recon, target)
    return torch.mean(e), torch.norm(e)


def frame_scores(recon, clip):
    """per-frame error -> psnr (tool/contrast_evaluae.py:232-238, misc/utils.py:124-128)"""
    loss = torch.nn.MSELoss(reduction='none')(recon, clip)
    loss = loss.permute(0, 2, 1, 3, 4)                                    # 'B C D H W -> B D C H W'
    lf = torch.mean(loss, dim=4).mean(dim=3).mean(dim=2)
    lf = sum(lf.tolist(), [])
    return lf, [10 * math.log10(1.0 / m) for m in lf]


def memory_forward(query, keys, train=True):
    """Memory.forward (model/Memory.py:145-175) with the m-iteration python
    loop of get_update_query (:94-131) kept as the reference has it — that
    loop is part of what the reference costs on a CPU."""
    B, d, h, w = query.shape
    q4 = F.normalize(query, dim=1).permute(0, 2, 3, 1)

    def get_score(mem, q):
        s = torch.matmul(q, mem.t()).view(-1, mem.shape[0])
        return F.softmax(s, dim=0), F.softmax(s, dim=1)

    qr = q4.contiguous().view(B * h * w, d)
    _, sm = get_score(keys, q4)
    _, g1 = torch.topk(sm, 1, dim=1)
    gathering = torch.nn.MSELoss()(qr, keys[g1].squeeze(1).detach())
    out_spread = None
    if train:
        _, sm = get_score(keys, q4)
        _, g2 = torch.topk(sm, 2, dim=1)
        out_spread = torch.nn.TripletMarginLoss(margin=1.0)(qr, keys[g2[:, 0]].detach(), keys[g2[:, 1]].detach())
    sq, sm = get_score(keys, q4)
    concat = torch.matmul(sm.detach(), keys)
    uq = torch.cat((qr, concat), dim=1).view(B, h, w, 2 * d).permute(0, 3, 1, 2)
    if not train:
        return uq, keys, sq, sm, gathering
    sq2, sm2 = get_score(keys, q4)
    _, gi = torch.topk(sm2, 1, dim=1)
    m = keys.shape[0]
    upd = torch.zeros((m, d))
    for i in range(m):
        idx = torch.nonzero(gi.squeeze(1) == i)
        if idx.shape[0] != 0:
            upd[i] = torch.sum((sq2[idx, i] / torch.max(sq2[:, i])) * qr[idx].squeeze(1), dim=0)
    um = F.normalize(upd + keys, dim=1)
    return uq, um.detach(), sq, sm, gathering, out_spread

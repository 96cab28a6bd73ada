"""Swap the B200 classes into an importable copy of the reference so that
``model.backbone.Mymodel`` (model/backbone.py:40-41) builds with them — the
drop-in test of SURVEY.md §4(ii).  Only used where the reference tree exists."""
import os
import sys

import numpy as np
import torch


def patch_reference():
    """Replace model.cluster's / model.Memory's / loss_tool.Recon_Loss's classes and
    misc.utils' psnr / anomly_score in ``sys.modules`` with the videoad_b200 ones.
    Returns the list of patched attribute names."""
    from . import cluster, memory, losses, scoring
    patched = []
    table = {
        "model.cluster": {
            "EuclidDistance_Assign_Module": cluster.EuclidDistance_Assign_Module,
            "Space_EuclidDistance_Assign_Module": cluster.Space_EuclidDistance_Assign_Module,
            "NegSoftAssign": cluster.NegSoftAssign, "PosSoftAssign": cluster.PosSoftAssign},
        "model.Memory": {"Memory": memory.Memory, "MemoryLoss": memory.MemoryLoss},
        "loss_tool.Recon_Loss": {"Recon_Loss": losses.Recon_Loss},
        "misc.utils": {"psnr": scoring.psnr, "anomly_score": scoring.anomly_score},
        "model.backbone": {"cluster": cluster.EuclidDistance_Assign_Module,
                           "space_cluster": cluster.Space_EuclidDistance_Assign_Module},
    }
    for mod_name, attrs in table.items():
        mod = sys.modules.get(mod_name)
        if mod is None:
            continue
        for k, v in attrs.items():
            setattr(mod, k, v)
            patched.append(f"{mod_name}.{k}")
    return patched


def load_pretrain_model(ckp_path, model, map_location=None, verbose=False):
    """Checkpoint loader with the on-disk contract of misc/utils.py:51-76: the
    reference saves ``DistributedDataParallel(model).state_dict()`` (main_predict.py:204,339), so every key
    carries a 7-character ``module.`` prefix that the loader strips; keys that then match the model's
    ``state_dict`` overwrite it, others are reported and skipped; a missing file is a silent no-op (:52-53).
    The cluster heads' entries (``cluster1.cluster_center [K,C]``, ``cluster1.identity_matrix [K,K]``,
    ``cluster1.norm.weight/bias``, ``space_cluster.cluster_center [C,K,P]``, ...) keep the reference's names and
    shapes, so the authors' checkpoints load into the drop-in modules unchanged.  Returns (loaded, skipped) key
    lists.  ``map_location`` defaults to the reference's ``'cuda:0'`` when CUDA is present."""
    if not os.path.isfile(ckp_path):
        return [], []
    if map_location is None:
        map_location = "cuda:0" if torch.cuda.is_available() else "cpu"
    checkpoint = torch.load(ckp_path, map_location=map_location)
    model_dict = model.state_dict()
    loaded, skipped = [], []
    for key, value in checkpoint.items():
        key = key[7:]                                   # 'module.' (misc/utils.py:62)
        if key in model_dict and value is not None:
            model_dict[key] = value
            loaded.append(key)
        else:
            skipped.append(key)
        if verbose:
            print(("=> loaded '{}' from checkpoint '{}'" if key in loaded[-1:] else
                   "=> key '{}' not found in checkpoint: '{}'").format(key, ckp_path))
    model.load_state_dict(model_dict)
    return loaded, skipped


def save_checkpoint(model, path):
    """what main_predict.py:204,339-340 writes: the state_dict of the DDP-wrapped model, i.e. every key behind
    a ``module.`` prefix (so that ``load_pretrain_model`` of either code base reads it back)"""
    sd = model.state_dict()
    if not all(k.startswith("module.") for k in sd):
        sd = {"module." + k: v for k, v in sd.items()}
    torch.save(sd, path)


class ClusterFeatureBank:
    """The ``feature`` / ``feature_label`` consumer of the reference — 聚类可视化.py:117-160: tokens of
    every clip are grouped by their cluster label, and the tokens of four well-populated clusters go to t-SNE.

    The reference does this one token at a time on the host (``label.item()`` — a device sync per token — and an
    ``np.vstack`` per token, quadratic in the number of tokens).  Here ``add`` keeps the batch on the device (one
    ``bincount`` for the counts), and ``record`` / ``tsne_selection`` group ALL tokens with one stable sort by label and
    one device-to-host copy; the resulting arrays are identical to the reference's, row order included (tokens of a
    label stay in arrival order, as the vstack builds them)."""

    def __init__(self, num_clusters=1024):
        self.num_clusters = num_clusters
        self._feat, self._lab = [], []
        self._count = None

    def add(self, feature, feature_label):
        """feature [N,C], feature_label [N] (int64 or the float zeros ``Mymodel`` returns without clustering)"""
        lab = feature_label.reshape(-1).to(torch.int64)
        self._feat.append(feature.detach().reshape(lab.numel(), -1))
        self._lab.append(lab.to(feature.device))
        c = torch.bincount(self._lab[-1], minlength=self.num_clusters)
        self._count = c if self._count is None else self._count + c

    def label_num(self):
        """the reference's ``label_num`` (:118,:137-138): ones, plus one for every occurrence of a label after its first"""
        c = np.zeros(self.num_clusters) if self._count is None else self._count.cpu().numpy().astype(np.float64)
        return np.where(c > 0, c, 1.0)

    def record(self):
        """{label: [n_label, C] float32 array}, tokens in arrival order (:131-140)"""
        if not self._feat:
            return {}
        feat, lab = torch.cat(self._feat), torch.cat(self._lab)
        order = torch.argsort(lab, stable=True)
        feat, lab = feat[order].cpu().numpy(), lab[order].cpu().numpy()
        cuts = np.flatnonzero(np.diff(lab)) + 1
        return {int(l[0]): f for l, f in zip(np.split(lab, cuts), np.split(feat, cuts))}

    def tsne_selection(self):
        """(data [n,C], label [n]) exactly as :142-158 assemble them: the clusters with the 6th, 3rd, 4th and 5th largest
        token counts (``np.flip(np.argsort(label_num)[[-5, -4, -3, -6]])``), tagged 1..4"""
        rec = self.record()
        label_max = np.flip(np.argsort(self.label_num(), )[[-5, -4, -3, -6]])
        data = rec[int(label_max[0])]
        tag = 1
        label = np.ones(data.shape[0], dtype='int') * tag
        for num in label_max[1:]:
            tag = tag + 1
            temp = rec[int(num)]
            label = np.concatenate((label, np.ones(temp.shape[0], dtype='int') * tag))
            data = np.vstack((data, temp))
        return data, label

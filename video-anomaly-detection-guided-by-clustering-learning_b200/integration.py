"""Swap the B200 classes into an importable copy of the reference so that
``model.backbone.Mymodel`` (model/backbone.py:40-41) builds with them — the
drop-in test of SURVEY.md §4(ii).  Only used where the reference tree exists."""
import os
import sys

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

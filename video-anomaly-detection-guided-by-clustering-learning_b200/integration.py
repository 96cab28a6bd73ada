"""Swap the B200 classes into an importable copy of the reference so that
``model.backbone.Mymodel`` (model/backbone.py:40-41) builds with them — the
drop-in test of SURVEY.md §4(ii).  Only used where the reference tree exists."""
import sys


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

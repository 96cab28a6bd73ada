"""videoad_b200 — B200-native (sm_100a) implementation of the clustering /
memory / loss / scoring hot path of
Bun-TianYi/Video-anomaly-detection-guided-by-clustering-learning.

Python + PyTorch are plumbing (device memory, streams, autograd graph,
torch.distributed); the arithmetic runs in hand-written CUDA kernels behind the
C ABI in include/vadc.h (libvadc.so, loaded with ctypes).  There is no CPU or
PyTorch fallback: every op raises if the library or a CUDA device is missing.
"""
from . import _lib
from ._lib import IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05, launch_count
from .cluster import (EuclidDistance_Assign_Module, Space_EuclidDistance_Assign_Module,
                      NegSoftAssign, PosSoftAssign, cluster_alpha, cdist, soft_assign)
from .memory import Memory, MemoryLoss
from .losses import Recon_Loss, l1_mean, mse_mean, e4_norm, e4_sum
from .scoring import (frame_mse, clip_mse, psnr, anomly_score, roc_auc_score, regularity_auc,
                      minmax_score_device, evaluate_videos, eval_clip_starts, eval_clip_starts_stride1,
                      gather_video_scores)
from .distributed import (init_distributed_mode, fix_random_seeds, setup_for_distributed, get_sha,
                          allreduce_sum_packed, global_frobenius, shard_range,
                          numa_local, pinned_like_local, gpu_local_cpus, enable_peer_allreduce,
                          disable_peer_allreduce, peer_allreduce_status)
from .decoder_entry import norm_timedebd, fuse_decoder_entry
from .encoder_tail import downsample_gelu_tokens, fuse_encoder_tail
from .integration import patch_reference, load_pretrain_model, save_checkpoint, ClusterFeatureBank

__all__ = [
    "EuclidDistance_Assign_Module", "Space_EuclidDistance_Assign_Module", "NegSoftAssign",
    "PosSoftAssign", "cluster_alpha", "cdist", "soft_assign", "Memory", "MemoryLoss", "Recon_Loss",
    "l1_mean", "mse_mean", "e4_norm", "e4_sum", "frame_mse", "clip_mse", "psnr", "anomly_score",
    "roc_auc_score", "regularity_auc", "minmax_score_device", "evaluate_videos", "eval_clip_starts", "eval_clip_starts_stride1", "gather_video_scores", "init_distributed_mode",
    "fix_random_seeds", "setup_for_distributed", "get_sha", "allreduce_sum_packed",
    "global_frobenius", "shard_range", "numa_local", "pinned_like_local", "gpu_local_cpus", "enable_peer_allreduce", "disable_peer_allreduce", "peer_allreduce_status", "patch_reference", "load_pretrain_model", "save_checkpoint", "ClusterFeatureBank", "norm_timedebd", "fuse_decoder_entry", "downsample_gelu_tokens", "fuse_encoder_tail", "launch_count",
    "IMPL_AUTO", "IMPL_SIMT", "IMPL_TCGEN05",
]

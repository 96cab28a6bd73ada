"""CPU-side checks of the C-ABI boundary: the library loads, exports every
symbol include/vadc.h declares, and the host mirror fails loudly without CUDA
(no fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest
import torch

import videoad_b200 as V
from videoad_b200 import _lib


def test_header_declares_and_library_exports_every_symbol():
    protos = _lib.parse_header()
    with open(_lib.HEADER) as fh:
        declared = set(re.findall(r"\b(vadc_\w+)\s*\(", fh.read()))
    assert declared == set(protos), declared ^ set(protos)
    assert len(protos) >= 30
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(raw, name), f"libvadc.so does not export {name}"


def test_version_and_error_strings():
    l = _lib.lib()
    assert l.vadc_version().decode().startswith("vadc")
    for code in range(0, -8, -1):
        assert len(l.vadc_error_string(code).decode()) > 1
    assert "unknown" in l.vadc_error_string(-99).decode()


def test_workspace_queries_are_pure_host_functions():
    l = _lib.lib()
    assert l.vadc_cluster_fwd_workspace_bytes(1024, 192, 32, 0) > 0
    assert l.vadc_cluster_bwd_workspace_bytes(1024, 192, 32) > 1024 * 192 * 4
    assert l.vadc_space_cluster_fwd_workspace_bytes(8, 784, 192, 128) > 0
    assert l.vadc_memory_score_workspace_bytes(2048, 2000, 768) >= 2048 * 2000 * 4
    assert l.vadc_pixel_loss_workspace_bytes(1 << 20) > 0
    # round 2: opaque saved state of the space head (operand terms on the tensor-core path; without a device the query
    # answers for the fp32 fallback), decoder entry, encoder tail
    assert l.vadc_space_cluster_saved_bytes(8, 784, 192, 128) >= 192 * 8 * 784 * 4
    assert l.vadc_space_cluster_bwd_workspace_bytes(8, 784, 192, 128) > 0
    assert l.vadc_norm_timedebd_workspace_bytes(6272, 192) > 6272 * 192 * 4
    assert l.vadc_downsample_gelu_workspace_bytes(2, 96, 4, 28, 28, 192) > 2 * 4 * 28 * 28 * 96 * 4 * 6
    assert l.vadc_downsample_gelu_fwd(None, None, None, 2, 96, 4, 28, 28, 192, None, None, None, 0, None) == -2     # NULL pointers


def test_argument_validation_happens_before_any_cuda_call():
    l = _lib.lib()
    # bad shape (K % 4 != 0) and NULL pointers are rejected on the host
    assert l.vadc_cluster_fwd(None, None, None, None, 8, 192, 30, 16.0, 1e-5, None, None, None, None,
                              None, None, None, None, None, None, 0, 0, None) == -1
    assert l.vadc_cluster_fwd(None, None, None, None, 8, 192, 32, 16.0, 1e-5, None, None, None, None,
                              None, None, None, None, None, None, 0, 0, None) == -2
    assert l.vadc_pixel_loss(None, None, 0, None, 0, 0, None, None, 0, None) == -1
    assert l.vadc_pixel_loss(None, None, 16, None, 0, 7, None, None, 0, None) == -1


def test_no_cpu_fallback():
    mod = V.EuclidDistance_Assign_Module(32, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(torch.randn(1, 1, 2, 2, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        V.e4_norm(torch.rand(4, 4), torch.rand(4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        V.frame_mse(torch.rand(1, 3, 2, 4, 4), torch.rand(1, 3, 2, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        V.Memory(10, 8, 8, 0.1, 0.1)(torch.randn(1, 8, 2, 2), torch.rand(10, 8))


def test_missing_library_is_a_loud_error(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libvadc.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_state_dict_keys_match_reference_contract():
    """SURVEY.md §5 checkpoint contract: parameter names and shapes"""
    m = V.EuclidDistance_Assign_Module(192, 1024, soft_assign_alpha=16.0)
    sd = m.state_dict()
    assert set(sd) == {"cluster_center", "identity_matrix", "norm.weight", "norm.bias"}
    assert sd["cluster_center"].shape == (1024, 192) and sd["identity_matrix"].shape == (1024, 1024)
    assert m.cluster_center.requires_grad and not m.identity_matrix.requires_grad
    assert float(m.cluster_center.min()) >= 0 and float(m.cluster_center.max()) <= 1   # torch.rand init
    s = V.Space_EuclidDistance_Assign_Module(8, 6, space_size=4)
    sd = s.state_dict()
    assert sd["cluster_center"].shape == (8, 6, 16) and sd["identity_matrix"].shape == (8, 6, 6)
    assert torch.equal(sd["identity_matrix"][3], torch.eye(6))
    assert m.assign_func.alpha == 16.0 and m.assign_func.dims == -1


def test_host_scoring_functions_match_reference_fixture():
    from conftest import load_golden
    import numpy as np
    g = load_golden("losses_scoring")
    mses, labs = [], []
    for i in range(int(g["n_videos"])):
        ms = g[f"mse{i}"].tolist()
        assert V.psnr(ms) == g[f"psnr{i}"].tolist()
        assert V.anomly_score(V.psnr(ms)) == g[f"score{i}"].tolist()
        mses.append(ms); labs.append(g[f"label{i}"])
    auc, per = V.regularity_auc(mses, labs, [str(s) for s in g["scenes"]])
    assert abs(auc - float(g["auc"])) < 1e-12
    np.testing.assert_allclose(list(per.values()), g["scene_aucs"], atol=1e-12)
    with pytest.raises(ZeroDivisionError):
        V.anomly_score([1.0, 1.0])
    with pytest.raises(ValueError):
        V.roc_auc_score([1, 1, 1], [0.1, 0.2, 0.3])


def test_roc_auc_matches_sklearn():
    import numpy as np
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(1)
    for n in (2, 17, 400):
        y = rng.integers(0, 2, n); y[0], y[-1] = 0, 1
        s = np.round(rng.random(n), 2)
        assert abs(V.roc_auc_score(y, s) - roc_auc_score(y, s)) < 1e-12


def test_checkpoint_contract_roundtrip(tmp_path):
    """state_dict names / shapes of the drop-in cluster heads and the DDP 'module.' prefix convention of
    misc/utils.py:51-76: a checkpoint written the way main_predict.py:204 writes it loads back,
    unknown keys are skipped, a missing file is a no-op"""
    import torch
    import videoad_b200 as V

    class Net(torch.nn.Module):                      # the part of Mymodel the hot path owns (backbone.py:40-41)
        def __init__(self):
            super().__init__()
            self.cluster1 = V.EuclidDistance_Assign_Module(192, 1024, soft_assign_alpha=16.0)
            self.space_cluster = V.Space_EuclidDistance_Assign_Module(12, 8, space_size=4)

    a, b = Net(), Net()
    assert {k: tuple(v.shape) for k, v in a.state_dict().items()} == {
        "cluster1.cluster_center": (1024, 192), "cluster1.identity_matrix": (1024, 1024),
        "cluster1.norm.weight": (192,), "cluster1.norm.bias": (192,),
        "space_cluster.cluster_center": (12, 8, 16), "space_cluster.identity_matrix": (12, 8, 8),
        "space_cluster.norm.weight": (12,), "space_cluster.norm.bias": (12,)}
    # the reference's substring-based grad toggles (backbone.py:44-72) see the same names
    assert all(("cluster" in n) for n, _ in a.named_parameters())
    assert not a.cluster1.identity_matrix.requires_grad and a.cluster1.cluster_center.requires_grad
    path = str(tmp_path / "checkpoint0.pth")
    V.save_checkpoint(a, path)
    sd = torch.load(path)
    assert all(k.startswith("module.") for k in sd)
    sd["module.decoder.not_in_this_model"] = torch.zeros(3)
    torch.save(sd, path)
    loaded, skipped = V.load_pretrain_model(path, b, map_location="cpu")
    assert skipped == ["decoder.not_in_this_model"] and len(loaded) == 8
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(va, vb), k
    assert V.load_pretrain_model(str(tmp_path / "missing.pth"), b) == ([], [])


def test_clip_schedule_host_logic_matches_oracle_and_cfg4_layout():
    """the product's clip schedule (host integers, tool/contrast_evaluae.py:185-203) equals the oracle's for every
    (length, frame_num, batch_size) in a sweep; the synthetic cfg4 layout is one the reference's batched loop accepts"""
    import videoad_b200 as V
    from oracle import np_oracle as O
    import bench
    for T in list(range(0, 70)) + [381, 720]:
        for fn in (2, 4, 8, 10):
            for bs in (1, 3, 16):
                assert V.eval_clip_starts(T, fn, bs) == O.eval_clip_starts(T, fn, bs), (T, fn, bs)
    lengths, labels, scenes = bench.cfg4_layout()
    assert len(lengths) == 107 and sum(lengths) == 40791 and len(set(scenes)) == 12
    for T, lab in zip(lengths, labels):
        assert T % 8 in (0, 1) and len(lab) == T and 0 < lab.sum() < T          # both classes in every video
        assert all(len(b) > 0 for b in V.eval_clip_starts(T, 8, 16))


def test_timing_api_is_safe_without_a_device():
    """vadc_timing_enable / vadc_timing_read: no launches recorded -> count 0, no CUDA call"""
    import ctypes
    import videoad_b200 as V
    l = V._lib.lib()
    ms, cnt = ctypes.c_float(1.0), ctypes.c_int(-1)
    assert l.vadc_timing_enable(1) == 0 and l.vadc_timing_read(1, ctypes.byref(ms), ctypes.byref(cnt)) == 0
    assert cnt.value == 0 and ms.value == 0.0
    assert l.vadc_timing_read(7, ctypes.byref(ms), ctypes.byref(cnt)) != 0
    assert l.vadc_timing_enable(0) == 0


def test_cluster_feature_bank_matches_the_reference_visualisation_loop():
    """8f-4: the grouped feature dump == the reference's per-token record / label_num / t-SNE selection
    (聚类可视化.py:117-158, restated in oracle.np_oracle.cluster_feature_record); host logic, runs on CPU tensors"""
    import numpy as np
    import torch
    import videoad_b200 as V
    from oracle import np_oracle as O
    rng = np.random.default_rng(0)
    batches = []
    for n in (70, 55, 90):
        lab = rng.choice(12, size=n, p=np.arange(1, 13) / 78.0)          # 12 clusters of distinct popularity
        batches.append((rng.standard_normal((n, 5)).astype(np.float32), lab))
    bank = V.ClusterFeatureBank(num_clusters=16)
    for f, l in batches:
        bank.add(torch.tensor(f), torch.tensor(l))
    rec, label_num, data, label = O.cluster_feature_record(batches, 16)
    np.testing.assert_array_equal(bank.label_num(), label_num)
    got = bank.record()
    assert set(got) == set(rec)
    for k in rec:
        np.testing.assert_array_equal(got[k], np.atleast_2d(rec[k]))
    d2, l2 = bank.tsne_selection()
    np.testing.assert_array_equal(d2, data)
    np.testing.assert_array_equal(l2, label)


def test_fuse_functions_leave_parameters_alone_and_pass_cpu_inputs_through():
    """``fuse_encoder_tail`` / ``fuse_decoder_entry`` bind new forwards on the reference's module instances: the state_dict
    is untouched, and inputs the CUDA kernels do not take (here: CPU tensors) run through the original modules — or, for
    the decoder entry whose LayerNorm was folded in, fail loudly instead of silently skipping the normalisation"""
    import torch
    import torch.nn as nn
    import videoad_b200 as V

    class Enc(nn.Module):
        def __init__(self):
            super().__init__()
            self.downsample = nn.ModuleList([nn.Sequential(nn.Conv3d(4, 8, (1, 2, 2), stride=(1, 2, 2)), nn.GELU()), nn.Identity()])

    class Dec(nn.Module):
        def __init__(self):
            super().__init__()
            self.timedebd = nn.Conv3d(8, 8, (2, 1, 1), stride=(2, 1, 1))

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder, self.decoder, self.norm = Enc(), Dec(), nn.LayerNorm(8)

    torch.manual_seed(0)
    m = M()
    keys = list(m.state_dict().keys())
    x = torch.randn(1, 4, 2, 6, 6)
    want = m.encoder.downsample[0](x)
    assert V.fuse_encoder_tail(m) == 1                     # the Identity stage is left alone
    assert list(m.state_dict().keys()) == keys
    assert torch.equal(m.encoder.downsample[0](x), want)   # CPU input: the original conv + GELU
    V.fuse_decoder_entry(m)
    assert list(m.state_dict().keys()) == keys
    with pytest.raises(RuntimeError):                      # no CPU path behind the fused entry
        m.decoder.timedebd(torch.randn(1, 8, 2, 3, 3))

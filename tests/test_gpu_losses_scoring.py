"""L2 / L3 pixel losses and E1-E4 scoring parity on the GPU through the C ABI."""
import numpy as np
import pytest
import torch

import videoad_b200 as V
from oracle import np_oracle as O
from conftest import load_golden
from gpu_util import T, N, rel, dev

pytestmark = pytest.mark.gpu


def test_losses_golden():
    g = load_golden("losses_scoring")
    rl = V.Recon_Loss((2, 4, 4))
    assert abs(float(rl(T(g["l1_x"]), T(g["l1_t_pad"]))) - float(g["l1_pad"])) < 1e-6
    assert abs(float(rl(T(g["l1_x"]), T(g["l1_t"]))) - float(g["l1"])) < 1e-6
    assert abs(float(V.mse_mean(T(g["l1_x"]), T(g["l1_t"]))) - float(g["mse"])) < 1e-6
    assert abs(float(V.e4_norm(T(g["l1_x"]), T(g["l1_t"]))) - float(g["e4"])) < 1e-5 * float(g["e4"])
    with pytest.raises(AssertionError):
        rl(T(g["l1_x"])[:, :, :3], T(g["l1_t"]))            # Recon_Loss.py:27 shape assert


@pytest.mark.parametrize("shape", [(2, 3, 8, 64, 64), (1, 3, 5, 37, 41), (1, 1, 1, 1, 3), (2, 3, 16, 256, 256)])
def test_losses_vs_oracle_and_autograd(shape):
    rng = np.random.default_rng(sum(shape))
    x = rng.random(shape).astype(np.float32)
    t = rng.random(shape).astype(np.float32)
    for fn, ofn, tfn in (
        (V.l1_mean, lambda a, b: O.recon_l1(a, b), lambda a, b: torch.nn.functional.l1_loss(a, b)),
        (V.mse_mean, O.mse_mean, lambda a, b: torch.mean(torch.nn.MSELoss(reduction="none")(a, b))),
        (V.e4_norm, O.e4_norm, lambda a, b: torch.norm(torch.nn.MSELoss(reduction="none")(a, b))),
    ):
        xt = T(x, grad=True)
        out = fn(xt, T(t))
        ref = float(ofn(x, t))
        assert abs(float(out) - ref) < 1e-5 * max(abs(ref), 1e-12)
        (out * 3.0).backward()
        xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        (tfn(xr, torch.tensor(t, dtype=torch.float64)) * 3.0).backward()
        assert rel(N(xt.grad), xr.grad.numpy()) < 1e-4
    xt = T(x, grad=True)
    s = V.e4_sum(xt, T(t))
    assert abs(float(s) - float(O.e4_norm(x, t)) ** 2) < 1e-5 * float(s)
    torch.sqrt(s).backward()
    xr = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    torch.norm((xr - torch.tensor(t, dtype=torch.float64)) ** 2).backward()
    assert rel(N(xt.grad), xr.grad.numpy()) < 1e-4


def test_scoring_golden_end_to_end_auc():
    """per-frame MSE on the GPU -> psnr -> anomly_score -> per-scene AUC == the
    reference evaluation loop's value (AUC within 1e-4: BASELINE.json)"""
    g = load_golden("losses_scoring")
    mses, labs = [], []
    for i in range(int(g["n_videos"])):
        m = V.frame_mse(T(g[f"recon{i}"]), T(g[f"clip{i}"]))
        assert m.shape == (1, g[f"mse{i}"].shape[0])
        assert rel(N(m)[0], g[f"mse{i}"]) < 2e-6
        mses.append(N(m)[0]); labs.append(g[f"label{i}"])
    auc, per = V.regularity_auc(mses, labs, [str(s) for s in g["scenes"]])
    assert abs(auc - float(g["auc"])) < 1e-4
    np.testing.assert_allclose(list(per.values()), g["scene_aucs"], atol=1e-4)


@pytest.mark.parametrize("B,C,Tt,H,W", [(2, 3, 8, 224, 224), (1, 3, 5, 30, 22), (3, 1, 4, 64, 64), (1, 3, 1, 7, 9)])
def test_frame_mse_vs_oracle(B, C, Tt, H, W):
    rng = np.random.default_rng(B + H)
    clip = rng.random((B, C, Tt, H, W)).astype(np.float32)
    recon = (clip + 0.05 * rng.standard_normal(clip.shape)).astype(np.float32)
    m, ps = V.frame_mse(T(recon), T(clip), want_psnr=True)
    ref = O.frame_mse(recon, clip)
    assert rel(N(m), ref) < 2e-6
    np.testing.assert_allclose(N(ps).reshape(-1), O.psnr(N(m).reshape(-1).tolist()), rtol=1e-12)
    cm = V.clip_mse(T(recon), T(clip))
    assert rel(N(cm), O.clip_mse(recon, clip)) < 2e-6
    assert V.frame_mse(T(recon)[:0], T(clip)[:0]).shape == (0, Tt)


def test_minmax_score_device_and_synthetic_auc():
    """synthetic ShanghaiTech-shaped labelled set at reduced resolution: the
    device pipeline (frame_mse+psnr+minmax) equals the host float64 functions and
    the AUC equals the oracle's"""
    rng = np.random.default_rng(0)
    lens = [24, 40, 16, 32, 48, 24]
    scenes = ["01", "01", "02", "03", "02", "03"]
    mses, labs, ps_all = [], [], []
    for Tn in lens:
        clip = rng.random((1, 3, Tn, 32, 32)).astype(np.float32)
        lab = (rng.random(Tn) < 0.4).astype(np.int64); lab[0], lab[1] = 0, 1
        recon = clip + (0.05 * rng.standard_normal(clip.shape) * (1 + 2 * lab)[None, None, :, None, None]).astype(np.float32)
        m, ps = V.frame_mse(T(recon), T(clip), want_psnr=True)
        mses.append(N(m)[0]); labs.append(lab); ps_all.append(ps[0])
    off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int64, device=dev())
    score = V.minmax_score_device(torch.cat(ps_all), off)
    host = np.concatenate([V.anomly_score(V.psnr(m.tolist())) for m in mses])
    np.testing.assert_allclose(N(score), host, rtol=0, atol=1e-12)
    auc, _ = V.regularity_auc(mses, labs, scenes)
    oauc, _ = O.scene_auc([m.tolist() for m in mses], labs, scenes)
    assert abs(auc - oauc) < 1e-12 and 0.5 < auc <= 1.0


@pytest.mark.parametrize("name", ["eval_loop_b1", "eval_loop_b3"])
def test_device_resident_evaluation_loop(name):
    """V.evaluate_videos (videos resident on the GPU, fused per-frame MSE+PSNR, device min-max, one
    host transfer) against the AUCs of the reference's own loop (golden) and the oracle's per-video scores"""
    g = load_golden(name)
    n = len(g["lengths"])
    videos = [g[f"video{i}"].astype(np.float32) for i in range(n)]
    labels = [g[f"label{i}"] for i in range(n)]
    scenes = [str(s) for s in g["scenes"]]
    fn, bs = int(g["frame_num"]), int(g["batch_size"])
    auc, per, scores, labs = V.evaluate_videos(lambda c: c + 0.05 * torch.sin(37.0 * c) * (1.0 + c),
                                               [torch.tensor(v) for v in videos], labels, scenes, fn, bs)
    assert abs(auc - float(g["auc"])) < 1e-4                      # north_star: frame-level AUC within 1e-4
    np.testing.assert_allclose(list(per.values()), g["scene_aucs"], atol=1e-4)
    _, _, oscores, olabs = O.evaluate_videos(lambda c: c + 0.05 * np.sin(37.0 * c) * (1.0 + c), videos, labels, scenes,
                                             fn, bs, dtype=np.float32)
    for a, b, la, lb in zip(scores, oscores, labs, olabs):
        np.testing.assert_allclose(a, b, atol=2e-5)
        assert list(la) == list(lb)


def _tiny_t(variant, ispredict):
    if variant == "predict" and ispredict:
        return lambda c: c[:, :, -1:] + 0.05 * torch.sin(37.0 * c[:, :, -1:]) * (1.0 + c.mean(dim=2, keepdim=True))
    return lambda c: c + 0.05 * torch.sin(37.0 * c) * (1.0 + c)


def _tiny_n(variant, ispredict):
    if variant == "predict" and ispredict:
        return lambda c: c[:, :, -1:] + 0.05 * np.sin(37.0 * c[:, :, -1:]) * (1.0 + c.mean(2, keepdims=True))
    return lambda c: c + 0.05 * np.sin(37.0 * c) * (1.0 + c)


@pytest.mark.parametrize("name,variant", [("eval_predict_b1", "predict"), ("eval_predict_b3_pred", "predict"),
                                          ("eval_first_frame", "first_frame")])
def test_other_evaluation_modes(name, variant):
    """the reference's two other evaluators (tool/predict_evaluae.py:170-284: clips one frame apart, one score per clip,
    optional frame prediction; main_predict.py:389-457: first-frame error) through the device-resident loop: the AUCs those
    loops print (golden, exec'd unmodified) within 1e-4, per-video scores against the oracle"""
    g = load_golden(name)
    n = len(g["lengths"])
    videos = [g[f"video{i}"].astype(np.float32) for i in range(n)]
    labels = [g[f"label{i}"] for i in range(n)]
    scenes = [str(s) for s in g["scenes"]]
    fn, bs, isp = int(g["frame_num"]), int(g["batch_size"]), bool(int(g["ispredict"]))
    auc, per, scores, labs = V.evaluate_videos(_tiny_t(variant, isp), [torch.tensor(v) for v in videos], labels, scenes,
                                               fn, bs, mode=variant, ispredict=isp)
    assert abs(auc - float(g["auc"])) < 1e-4
    if variant == "predict":
        np.testing.assert_allclose(list(per.values()), g["scene_aucs"], atol=1e-4)
        _, _, oscores, olabs = O.evaluate_videos_predict(_tiny_n(variant, isp), videos, labels, scenes, fn, bs, isp, dtype=np.float32)
    else:
        _, _, oscores, olabs = O.evaluate_videos_first_frame(_tiny_n(variant, isp), videos, labels, scenes, fn, dtype=np.float32)
    for a, b, la, lb in zip(scores, oscores, labs, olabs):
        np.testing.assert_allclose(a, b, atol=5e-5)
        assert list(la) == list(lb)


def test_evaluation_loop_degenerate_videos_raise_like_the_reference():
    """misc/utils.py:128,135: a perfectly reconstructed frame (mse 0) and a video of constant PSNR divide by zero; the
    device loop raises the same ZeroDivisionError instead of feeding NaN scores to the AUC"""
    v = torch.rand(3, 13, 8, 8)
    lab = np.array([0, 1] * 6 + [0])
    with pytest.raises(ZeroDivisionError):
        V.evaluate_videos(lambda c: c.clone(), [v], [lab], ["01"], 4, 1)                  # recon == clip: mse = 0
    const = torch.full((3, 13, 8, 8), 0.25)
    with pytest.raises(ZeroDivisionError):
        V.evaluate_videos(lambda c: c + 0.1, [const], [lab], ["01"], 4, 1)                # every frame: mse = 0.01


def test_frame_mse_reads_strided_clip_views_in_place():
    """clip batches are views of the resident video (consecutive or overlapping clips, single frames): same numbers as
    the contiguous copy, no copy made"""
    v = torch.rand(3, 40, 16, 20, device=dev())
    from videoad_b200.scoring import _clip_batch_view, _plane_view
    for step in (8, 1):
        view = _clip_batch_view(v, 3, 4, step, 8)
        assert view.data_ptr() == v[:, 3].data_ptr() and _plane_view(view).data_ptr() == view.data_ptr()
        ref = torch.stack([v[:, 3 + i * step: 3 + i * step + 8] for i in range(4)])
        assert torch.equal(view, ref)
        recon = ref + 0.05 * torch.randn_like(ref)
        assert torch.equal(V.frame_mse(recon, view), V.frame_mse(recon, ref))
        assert torch.equal(V.frame_mse(recon[:, :, :1], view[:, :, :1]), V.frame_mse(recon[:, :, :1].contiguous(), ref[:, :, :1].contiguous()))


def test_evaluation_loop_ragged_clip_raises_like_the_reference():
    """with batch_size > 1 the reference's loop raises RuntimeError (torch.cat) on a video whose tail clip is
    short; the drop-in raises the same exception type from torch.stack"""
    v = torch.rand(3, 23, 4, 4)
    with pytest.raises(RuntimeError):
        V.evaluate_videos(lambda c: c, [v], [np.zeros(23, np.int64)], ["01"], 4, 3)


from bench import cfg4_layout          # the synthetic cfg4 layout is shared with bench.py's scoring line


def test_cfg4_full_size_scoring_auc():
    """BASELINE configs[3] at full size (107 videos / 40 791 frames of 3x256x256, 32 GB of clips streamed through
    the device-resident loop).  The synthetic model adds a known per-frame offset a_f to every pixel, so the
    per-frame MSE is a_f^2 in closed form: scores and the mean per-scene AUC of the whole pipeline are checked
    against float64 host arithmetic (north_star: AUC within 1e-4) and against the oracle's AUC on the same MSEs."""
    fn, bs, S = 8, 16, 256
    lengths, labels, scenes = cfg4_layout()
    rng = np.random.default_rng(11)
    total_scored, all_mse, kept_labels = 0, [], []
    offsets = [(0.02 + 0.03 * rng.random(T) + 0.02 * lab * rng.random(T)).astype(np.float32)
               for T, lab in zip(lengths, labels)]
    state = {"vid": -1, "batches": None}

    def video_iter():
        for i, T in enumerate(lengths):
            state["vid"], state["batches"] = i, iter(V.eval_clip_starts(T, fn, bs))
            yield torch.rand(3, T, S, S, device=dev())           # one video resident at a time (<= 0.5 GB)

    def model_fn(clip):
        starts = next(state["batches"])
        off = np.stack([offsets[state["vid"]][s0:s0 + fn] for s0 in starts])          # [B, D]
        return clip + torch.tensor(off, device=dev())[:, None, :, None, None]

    auc, per, scores, labs = V.evaluate_videos(model_fn, video_iter(), labels, scenes, fn, bs)
    assert len(scores) == len(lengths) and len(per) == 12
    exp_mse = []
    for i, T in enumerate(lengths):
        keep = np.concatenate([np.arange(s0, s0 + fn) for st in V.eval_clip_starts(T, fn, bs) for s0 in st])
        assert len(scores[i]) == len(keep) and list(labs[i]) == list(labels[i][keep])
        assert scores[i].min() == 0.0 and scores[i].max() == 1.0                       # per-video min-max
        exp_mse.append(offsets[i][keep].astype(np.float64) ** 2)
        total_scored += len(keep)
    assert total_scored > 40000
    # closed form: psnr = 10 log10(1 / a^2) -> 1 - minmax per video -> per-scene AUC -> mean
    oauc, oper = O.scene_auc([m.tolist() for m in exp_mse], labs, scenes)
    assert abs(auc - oauc) < 1e-4 and 0.55 < auc < 0.999
    for k in oper:
        assert abs(per[k] - oper[k]) < 1e-4
    for i in (0, 53, 106):                                                             # scores, not only their ranks
        p = 10 * np.log10(1.0 / exp_mse[i])
        np.testing.assert_allclose(scores[i], 1 - (p - p.min()) / (p.max() - p.min()), atol=2e-4)

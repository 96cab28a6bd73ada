"""The numpy oracle vs fixtures produced by the reference's own modules
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import np_oracle as O
from conftest import load_golden, rel_err

CLUSTER = ["cluster_c64_k32", "cluster_c32_k16", "cluster_c192_k48_peaked"]
SPACE = ["space_c8_k6_p16", "space_c16_k40_p36"]
MEMORY = ["memory_d32_m10", "memory_d64_m50"]


def assert_selfdist_close(S, Sref):
    """cdist(centers, centers): sqrt amplifies fp32 cancellation noise on the
    (near-)zero diagonal (the reference's own diagonal is ~3e-3, not 0), so the
    comparison is on the squared distance, relative to its scale."""
    S2, R2 = np.asarray(S, np.float64) ** 2, np.asarray(Sref, np.float64) ** 2
    assert np.abs(S2 - R2).max() < 2e-6 * R2.max()


@pytest.mark.parametrize("name", CLUSTER)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cluster_forward(name, dtype):
    g = load_golden(name)
    f = O.cluster_forward(g["x"], g["centers"], g["ln_w"], g["ln_b"], float(g["alpha"]), dtype=dtype)
    assert rel_err(f["feature"], g["feature"]) < 2e-6
    assert rel_err(f["D"], g["D"]) < 2e-6
    np.testing.assert_allclose(f["A"], g["A"], rtol=2e-4, atol=1e-7)
    assert rel_err(f["x_rec"], g["x_rec"]) < 5e-6
    assert_selfdist_close(f["S"], g["S"])
    np.testing.assert_array_equal(f["label"], g["label"])
    assert abs(O.frobenius_loss(f["D"], f["A"], dtype) - g["cluster_loss"]) < 1e-5 * g["cluster_loss"]


@pytest.mark.parametrize("name", CLUSTER)
def test_cluster_backward(name):
    g = load_golden(name)
    for dtype, tol in ((np.float64, 2e-5), (np.float32, 2e-4)):
        f = O.cluster_forward(g["x"], g["centers"], g["ln_w"], g["ln_b"], float(g["alpha"]), dtype=dtype)
        gD, gA = O.frobenius_loss_grads(f["D"], f["A"], 1.0, dtype)
        gx, gc, gw, gb = O.cluster_backward(g["x"], g["centers"], g["ln_w"], g["ln_b"], float(g["alpha"]),
                                            gD=gD, gA=gA, gR=g["gR"], gF=g["gF"], dtype=dtype)
        assert rel_err(gx.reshape(g["gx"].shape), g["gx"]) < tol
        assert rel_err(gc, g["gcenters"]) < tol
        assert rel_err(gw, g["g_ln_w"]) < tol
        assert rel_err(gb, g["g_ln_b"]) < tol


@pytest.mark.parametrize("name", SPACE)
def test_space_cluster(name):
    g = load_golden(name)
    f = O.space_cluster_forward(g["x"], g["centers"], g["ln_w"], g["ln_b"], float(g["alpha"]))
    assert f["D"].shape == g["D"].shape
    assert rel_err(f["D"], g["D"]) < 2e-6
    np.testing.assert_allclose(f["A"], g["A"], rtol=2e-4, atol=1e-7)
    assert_selfdist_close(f["S"], g["S"])
    assert abs(O.frobenius_loss(f["D"], f["A"]) - g["space_loss"]) < 1e-5 * g["space_loss"]
    f64 = O.space_cluster_forward(g["x"], g["centers"], g["ln_w"], g["ln_b"], float(g["alpha"]), dtype=np.float64)
    gD, gA = O.frobenius_loss_grads(f64["D"], f64["A"], 1.0, np.float64)
    gx, gc, gw, gb = O.space_cluster_backward(g["x"], g["centers"], g["ln_w"], g["ln_b"], float(g["alpha"]),
                                              gD=gD, gA=gA, dtype=np.float64)
    assert rel_err(gx, g["gx"]) < 5e-5
    assert rel_err(gc, g["gcenters"]) < 5e-5
    assert rel_err(gw, g["g_ln_w"]) < 5e-5
    assert rel_err(gb, g["g_ln_b"]) < 5e-5


@pytest.mark.parametrize("name", MEMORY)
def test_memory(name):
    g = load_golden(name)
    o = O.memory_forward(g["query"], g["keys"], train=True)
    assert rel_err(o["score_query"], g["score_query"]) < 1e-5
    assert rel_err(o["score_memory"], g["score_memory"]) < 1e-5
    assert rel_err(o["updated_query"], g["updated_query"]) < 1e-5
    assert rel_err(o["updated_memory"], g["updated_memory"]) < 1e-5
    assert abs(o["gathering_loss"] - g["gathering_loss"]) < 1e-5 * abs(g["gathering_loss"])
    assert abs(o["spreading_loss"] - g["spreading_loss"]) < 1e-5 * abs(g["spreading_loss"])
    t = O.memory_forward(g["query"], g["keys"], train=False)
    assert rel_err(t["updated_query"], g["test_updated_query"]) < 1e-5
    np.testing.assert_array_equal(t["updated_memory"], g["test_updated_memory"])
    assert abs(t["gathering_loss"] - g["test_gathering_loss"]) < 1e-5 * abs(g["test_gathering_loss"])
    assert abs(O.memory_separateness(g["keys"]) - g["separateness"]) < 1e-5 * abs(g["separateness"])
    # gradient with respect to the query against the reference module's own autograd
    gq = O.memory_query_backward(g["query"], g["keys"], o["top1"], o["top2"], g["g_updated_query"], 0.7, 0.3)
    assert rel_err(gq, g["g_query_train"]) < 2e-5
    gq = O.memory_query_backward(g["query"], g["keys"], t["top1"], None, g["g_updated_query"], 0.7, None)
    assert rel_err(gq, g["g_query_test"]) < 2e-5


def test_losses():
    g = load_golden("losses_scoring")
    assert abs(O.recon_l1(g["l1_x"], g["l1_t_pad"], patch_d=2) - g["l1_pad"]) < 1e-6
    assert abs(O.recon_l1(g["l1_x"], g["l1_t"], patch_d=2) - g["l1"]) < 1e-6
    assert abs(O.mse_mean(g["l1_x"], g["l1_t"]) - g["mse"]) < 1e-6
    assert abs(O.e4_norm(g["l1_x"], g["l1_t"]) - g["e4"]) < 1e-5 * g["e4"]
    with pytest.raises(AssertionError):
        O.recon_l1(g["l1_x"][:, :, :3], g["l1_t"], patch_d=2)


def test_scoring_and_auc():
    g = load_golden("losses_scoring")
    n = int(g["n_videos"])
    mses, labs = [], []
    for i in range(n):
        m = O.frame_mse(g[f"recon{i}"], g[f"clip{i}"])[0]
        assert rel_err(m, g[f"mse{i}"]) < 2e-6
        np.testing.assert_allclose(O.psnr(g[f"mse{i}"].tolist()), g[f"psnr{i}"], rtol=0, atol=0)
        np.testing.assert_allclose(O.anomly_score(O.psnr(g[f"mse{i}"].tolist())), g[f"score{i}"], rtol=0, atol=0)
        mses.append(g[f"mse{i}"].tolist()); labs.append(g[f"label{i}"])
    auc, per = O.scene_auc(mses, labs, [str(s) for s in g["scenes"]])
    assert abs(auc - float(g["auc"])) < 1e-12
    np.testing.assert_allclose(list(per.values()), g["scene_aucs"], atol=1e-12)
    with pytest.raises(ZeroDivisionError):
        O.anomly_score([3.0, 3.0])


def test_roc_auc_matches_sklearn_with_ties():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(0)
    for _ in range(20):
        y = rng.integers(0, 2, 200)
        y[0], y[1] = 0, 1
        s = np.round(rng.random(200), 1)          # many ties
        assert abs(O.roc_auc(y, s) - roc_auc_score(y, s)) < 1e-12
    with pytest.raises(ValueError):
        O.roc_auc(np.ones(5), np.arange(5.0))


@pytest.mark.parametrize("name", ["eval_loop_b1", "eval_loop_b3"])
def test_evaluation_loop_matches_the_references_own_loop(name):
    """oracle.evaluate_videos vs the AUCs printed by the reference's ``predict`` (tool/contrast_evaluae.py:170-300,
    exec'd unmodified by tests/golden/make_golden.py::gen_eval_loop) on the same videos and model"""
    g = load_golden(name)
    n = len(g["lengths"])
    videos = [g[f"video{i}"] for i in range(n)]
    labels = [g[f"label{i}"] for i in range(n)]
    model = lambda c: c + 0.05 * np.sin(37.0 * c) * (1.0 + c)  # noqa: E731  (TinyModel of the generator)
    auc, per, _, _ = O.evaluate_videos(model, [v.astype(np.float32) for v in videos], labels, [str(s) for s in g["scenes"]],
                                       int(g["frame_num"]), int(g["batch_size"]), dtype=np.float32)
    assert abs(auc - float(g["auc"])) < 1e-12
    np.testing.assert_allclose(list(per.values()), g["scene_aucs"], atol=1e-12)


def _tiny(variant, ispredict, xp):
    if variant == "predict" and ispredict:
        return lambda c: c[:, :, -1:] + 0.05 * xp.sin(37.0 * c[:, :, -1:]) * (1.0 + c.mean(2, keepdims=True))
    return lambda c: c + 0.05 * xp.sin(37.0 * c) * (1.0 + c)


@pytest.mark.parametrize("name,variant", [("eval_predict_b1", "predict"), ("eval_predict_b3_pred", "predict"),
                                          ("eval_first_frame", "first_frame")])
def test_other_evaluation_loops_match_the_references_own(name, variant):
    """oracle restatements of tool/predict_evaluae.py:170-284 and main_predict.py:389-457 vs the AUCs those loops print
    when exec'd unmodified (make_golden.py::gen_eval_loop_variant)"""
    g = load_golden(name)
    n = len(g["lengths"])
    videos = [g[f"video{i}"].astype(np.float32) for i in range(n)]
    labels = [g[f"label{i}"] for i in range(n)]
    scenes = [str(s) for s in g["scenes"]]
    fn, bs, isp = int(g["frame_num"]), int(g["batch_size"]), bool(int(g["ispredict"]))
    if variant == "predict":
        auc, per, _, _ = O.evaluate_videos_predict(_tiny(variant, isp, np), videos, labels, scenes, fn, bs, isp, dtype=np.float32)
        np.testing.assert_allclose(list(per.values()), g["scene_aucs"], atol=1e-12)
    else:
        auc, per, _, _ = O.evaluate_videos_first_frame(_tiny(variant, isp, np), videos, labels, scenes, fn, dtype=np.float32)
    assert abs(auc - float(g["auc"])) < 1e-12


def test_evaluation_clip_schedule():
    """tool/contrast_evaluae.py:185-203: strict '<' bounds drop a clip that would end exactly at T"""
    assert O.eval_clip_starts(8, 4, 1) == [[0]]
    assert O.eval_clip_starts(9, 4, 1) == [[0], [4]]
    assert O.eval_clip_starts(24, 4, 3) == [[0, 4, 8], [12, 16, 20]]      # the inner loop's bound is one frame looser
    assert O.eval_clip_starts(3, 4, 1) == []

"""oracle/ref_port.py (torch-CPU restatement of the reference's op chain, the
timed CPU baseline) vs the reference-generated fixtures and vs np_oracle."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import ref_port as P
from conftest import load_golden, rel_err


@pytest.mark.parametrize("name", ["cluster_c64_k32", "cluster_c32_k16", "cluster_c192_k48_peaked"])
def test_cluster_port_matches_reference_fixture(name):
    g = load_golden(name)
    t = lambda a: torch.tensor(a)
    D, A, S, R, F, lab = P.cluster_head(t(g["x"]), t(g["centers"]), t(g["ln_w"]), t(g["ln_b"]), float(g["alpha"]))
    # same ATen ops as the reference -> bit-identical on the same machine
    assert torch.equal(D, t(g["D"])) and torch.equal(A, t(g["A"])) and torch.equal(R, t(g["x_rec"]))
    assert torch.equal(lab, t(g["label"])) and torch.equal(F, t(g["feature"]))
    loss, gx, gc, gw, gb = P.cluster_train_step(t(g["x"]), t(g["centers"]), t(g["ln_w"]), t(g["ln_b"]),
                                                float(g["alpha"]), t(g["gR"]).reshape(g["x"].shape))
    assert abs(float(loss) - float(g["cluster_loss"])) < 1e-6 * float(g["cluster_loss"])
    o = O.cluster_forward(g["x"], g["centers"], g["ln_w"], g["ln_b"], float(g["alpha"]), dtype=np.float64)
    gD, gA = O.frobenius_loss_grads(o["D"], o["A"], 1.0, np.float64)
    ogx, ogc, ogw, ogb = O.cluster_backward(g["x"], g["centers"], g["ln_w"], g["ln_b"], float(g["alpha"]),
                                            gD=gD, gA=gA, gR=g["gR"], dtype=np.float64)
    assert rel_err(gx.numpy().reshape(ogx.shape), ogx) < 2e-4 and rel_err(gc.numpy(), ogc) < 2e-4
    assert rel_err(gw.numpy(), ogw) < 2e-4 and rel_err(gb.numpy(), ogb) < 2e-4


@pytest.mark.parametrize("name", ["space_c8_k6_p16", "space_c16_k40_p36"])
def test_space_port(name):
    g = load_golden(name)
    t = lambda a: torch.tensor(a)
    Ds, As, S, rec = P.space_head(t(g["x"]), t(g["centers"]), t(g["ln_w"]), t(g["ln_b"]), float(g["alpha"]))
    assert rec == [] and torch.equal(Ds, t(g["D"])) and torch.equal(As, t(g["A"]))


@pytest.mark.parametrize("name", ["memory_d32_m10", "memory_d64_m50"])
def test_memory_port(name):
    g = load_golden(name)
    t = lambda a: torch.tensor(a)
    uq, um, sq, sm, gl, sl = P.memory_forward(t(g["query"]), t(g["keys"]), train=True)
    assert rel_err(uq.contiguous().numpy(), g["updated_query"]) < 1e-6
    assert rel_err(um.numpy(), g["updated_memory"]) < 1e-6
    assert abs(float(gl) - float(g["gathering_loss"])) < 1e-6 and abs(float(sl) - float(g["spreading_loss"])) < 1e-6


def test_losses_scoring_port():
    g = load_golden("losses_scoring")
    t = lambda a: torch.tensor(a)
    m, e4 = P.pixel_losses(t(g["l1_x"]), t(g["l1_t"]))
    assert abs(float(m) - float(g["mse"])) < 1e-7 and abs(float(e4) - float(g["e4"])) < 1e-6 * float(g["e4"])
    lf, ps = P.frame_scores(t(g["recon0"]), t(g["clip0"]))
    np.testing.assert_allclose(lf, g["mse0"], rtol=1e-6)
    np.testing.assert_allclose(ps, g["psnr0"], rtol=1e-6)

"""Generator loss: oracle/loss_ref.py against the goldens of the UNMODIFIED lib/core/loss.py TePoseLoss (CPU),
and tepose_b200.loss.TePoseLoss (tp_tepose_loss: values + gradients in one native call) against the same goldens (GPU)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import loss_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(f for f in os.listdir(GOLD) if f.startswith("loss_"))


def _case(fname):
    z = np.load(os.path.join(GOLD, fname))
    cfg = ast.literal_eval(str(z["cfg"]))
    return z, cfg, loss_ref.make_loss_case(cfg["seed"], cfg["batch"], cfg["n2d"])


@pytest.mark.parametrize("fname", CASES)
def test_oracle_loss_matches_reference_golden(fname):
    z, cfg, (out, d2, d3, pre, mosh) = _case(fname)
    out = {k: v.clone().requires_grad_(True) for k, v in out.items()}
    gen, dis, ld = loss_ref.tepose_loss([out], d2 if cfg["n2d"] else None, d3, pre, mosh, loss_ref.StubDiscriminator(cfg["seed"]))
    gen.backward()
    assert abs(float(gen) - float(z["gen_loss"])) < 1e-5 * abs(float(z["gen_loss"]))
    assert abs(float(dis) - float(z["dis_loss"])) < 1e-6
    for k, v in ld.items():
        assert abs(float(v) - float(z["term:" + k])) <= 1e-5 * max(1.0, abs(float(z["term:" + k]))), k
    for k, v in out.items():
        assert float((v.grad - torch.from_numpy(z["grad:" + k])).abs().max()) < 1e-6, k


@pytest.mark.gpu
@pytest.mark.parametrize("fname", CASES)
def test_native_loss_matches_reference_golden(fname):
    from tepose_b200.loss import TePoseLoss
    dev = "cuda:0"
    z, cfg, (out, d2, d3, pre, mosh) = _case(fname)
    to = lambda d: {k: v.to(dev) for k, v in d.items()}
    out = {k: v.to(dev).requires_grad_(True) for k, v in out.items()}
    disc = loss_ref.StubDiscriminator(cfg["seed"]).to(dev)
    crit = TePoseLoss(device=dev)
    gen, dis, ld = crit([out], to(d2) if cfg["n2d"] else None, to(d3), pre_mosh=pre.to(dev), data_motion_mosh=to(mosh), motion_discriminator=disc)
    gen.backward()
    assert set(ld) == {k[5:] for k in z.files if k.startswith("term:")}
    assert abs(float(gen) - float(z["gen_loss"])) < 2e-5 * abs(float(z["gen_loss"]))
    assert abs(float(dis) - float(z["dis_loss"])) < 1e-5
    for k, v in ld.items():
        assert abs(float(v) - float(z["term:" + k])) <= 2e-5 * max(1.0, abs(float(z["term:" + k]))), k
    for k, v in out.items():
        err = float((v.grad.cpu() - torch.from_numpy(z["grad:" + k])).abs().max())
        assert err < 2e-6, (k, err)
    # without a discriminator the data terms alone form the generator loss; twice the same call is bit-identical
    out2 = {k: v.detach().clone().requires_grad_(True) for k, v in out.items()}
    g1, _, _ = crit([out2], to(d2) if cfg["n2d"] else None, to(d3))
    g2, _, _ = crit([out2], to(d2) if cfg["n2d"] else None, to(d3))
    assert torch.equal(g1, g2)
    want = sum(float(z["term:" + k]) for k in ("loss_kp_2d", "loss_kp_3d", "loss_shape", "loss_pose") if "term:" + k in z.files)
    assert abs(float(g1) - want) < 2e-5 * want


@pytest.mark.gpu
def test_native_loss_drives_the_training_backward():
    """TePoseLoss on the train-mode outputs of TePose: gen_loss.backward() reaches every path parameter through the native backward."""
    from tepose_b200 import synthetic as psynth
    from tepose_b200.loss import TePoseLoss
    from tepose_b200.train import path_parameters
    dev = "cuda:0"
    B, T = 4, 6
    model, _ = psynth.build_synthetic_model(61, T, 1, 128, "fp32", dev)
    model.train()
    x = torch.from_numpy(psynth.make_input(61, B, T)).to(dev)
    _, d2, d3, pre, mosh = loss_ref.make_loss_case(61, B, 0)
    to = lambda d: {k: v.to(dev) for k, v in d.items()}
    out = model(x, is_train=True)
    gen, _, ld = TePoseLoss(device=dev)(out, None, to(d3))
    gen.backward()
    grads = [p.grad for _, p in path_parameters(model)]
    assert all(g is not None and bool(torch.isfinite(g).all()) for g in grads)
    assert sum(float(g.abs().sum()) for g in grads) > 0

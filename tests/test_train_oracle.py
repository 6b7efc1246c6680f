"""oracle/train_ref.py (train-mode forward + autograd gradients) against the goldens the UNMODIFIED reference modules
produced in train mode (oracle/make_golden.py:make_train_golden).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import synth, train_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "train_*.npz")))


def load_case(path):
    z = np.load(path)
    cfg = eval(str(z["cfg"]))      # noqa: S307 - our own fixture
    return z, cfg


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_train_oracle_matches_reference_golden(path):
    z, cfg = load_case(path)
    seed, B, T, L, H = cfg["seed"], cfg["batch"], cfg["seqlen"], cfg["n_layers"], cfg["hidden"]
    sd = synth.make_state_dict(seed, L, H)
    orc = train_ref.TrainOracle(sd, seed, L, H)
    out, loss, grads = orc.loss_and_grads(synth.make_input(seed, B, T), train_ref.make_masks(seed, 2 * B),
                                          train_ref.make_targets(seed, 2 * B))
    assert abs(loss - float(z["loss"])) <= 1e-6 * max(1.0, abs(loss))
    for k in ("theta", "kp_2d", "kp_3d", "rotmat"):
        assert float(np.abs(out[k].numpy() - z[k]).max()) <= 2e-6, k
    names = [k[5:] for k in z.files if k.startswith("grad:")]
    assert sorted(names) == sorted(grads)                         # every reference parameter has a gradient here and vice versa
    for k in names:
        ref = z["grad:" + k]
        got = train_ref.grad_probe(grads[k])
        scale = max(1e-12, float(np.abs(ref[2:]).max()), abs(ref[1]) / max(1, grads[k].numel()))
        assert float(np.abs(got[2:] - ref[2:]).max()) <= 1e-5 * scale + 1e-9, k
        assert abs(got[1] - ref[1]) <= 1e-5 * abs(ref[1]) + 1e-9, k


def test_masks_are_inputs_and_scale_like_nn_dropout():
    """An all-ones mask with scale 1/(1-p) is NOT the eval path: the train forward doubles the activations, exactly like
    nn.Dropout(p=0.5) would for a mask that keeps everything."""
    torch.manual_seed(0)
    W = {f"{n}.{s}": torch.randn(*shape) * 0.05 for n, s, shape in
         [("fc1", "weight", (1024, 2205)), ("fc1", "bias", (1024,)), ("fc2", "weight", (1024, 1024)), ("fc2", "bias", (1024,)),
          ("decpose", "weight", (144, 1024)), ("decpose", "bias", (144,)), ("decshape", "weight", (10, 1024)),
          ("decshape", "bias", (10,)), ("deccam", "weight", (3, 1024)), ("deccam", "bias", (3,))]}
    feat = torch.randn(4, 2048)
    init = (torch.zeros(4, 144), torch.zeros(4, 10), torch.zeros(4, 3))
    drop = torch.nn.Dropout()
    drop.train()
    ones = torch.ones(1, 2, 4, 1024)
    p, s, c = train_ref.ief_forward_train(W, feat, ones, init)
    xc = torch.cat([feat, *init], 1)
    a = torch.nn.functional.linear(xc, W["fc1.weight"], W["fc1.bias"]) * 2.0
    a = torch.nn.functional.linear(a, W["fc2.weight"], W["fc2.bias"]) * 2.0
    assert torch.allclose(p, torch.nn.functional.linear(a, W["decpose.weight"], W["decpose.bias"]), atol=1e-5)

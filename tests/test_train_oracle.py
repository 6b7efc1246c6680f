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


def test_smpl_backward_decomposition_matches_autograd():
    """The stage-by-stage backward the CUDA kernels implement (oracle/smpl_backward_proto.py) against torch.autograd through
    the oracle's SMPL forward + projection, in float64 (tight) -- verifies the kernel's algebra on the CPU."""
    from oracle import smpl_backward_proto as proto, torch_ref
    m = torch_ref.SmplModel.synthetic(5, dtype=torch.float64)
    g = torch.Generator().manual_seed(3)
    N = 3
    betas = torch.randn(N, 10, generator=g, dtype=torch.float64).requires_grad_(True)
    x6 = torch.randn(N, 144, generator=g, dtype=torch.float64)
    R = torch_ref.rot6d_to_rotmat(x6).reshape(N, 24, 3, 3).detach().requires_grad_(True)
    cam = (torch.tensor([0.9, 0.0, 0.0], dtype=torch.float64) + 0.1 * torch.randn(N, 3, generator=g, dtype=torch.float64)).requires_grad_(True)
    verts, j49, _ = torch_ref.smpl_forward(m, betas, R=R)
    kp2d = torch_ref.projection(j49, cam)
    gv = torch.randn(verts.shape, generator=g, dtype=torch.float64) * 0.01
    gj = torch.randn(j49.shape, generator=g, dtype=torch.float64)
    gk = torch.randn(kp2d.shape, generator=g, dtype=torch.float64)
    gRx = torch.randn(R.shape, generator=g, dtype=torch.float64)
    loss = (verts * gv).sum() + (j49 * gj).sum() + (kp2d * gk).sum() + (R * gRx).sum()
    loss.backward()
    g_R, g_b, g_c = proto.smpl_backward(m, torch_ref.JOINT_SOURCE_49, betas.detach(), R.detach(), cam.detach(), gv, gj, gk, gRx)
    rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))
    assert rel(g_R, R.grad) < 1e-10
    assert rel(g_b, betas.grad) < 1e-10
    assert rel(g_c, cam.grad) < 1e-10

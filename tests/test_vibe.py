"""VIBE bootstrap model (lib/models/vibe.py; evaluate.py:89-99,234): oracle vs the reference's golden
vectors, the host layer on the emulated C ABI (CPU), and the CUDA path on the GPU."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import synth, torch_ref
from tests import fake_native
from tests.helpers import compare_outputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")
VIBE = sorted(f for f in os.listdir(GOLD) if f.startswith("vibe_"))
KEYS = ("theta", "verts", "kp_2d", "kp_3d", "rotmat")
BF16_TOL = dict(vert_tol=1e-3, rot_tol=2e-3, kp2d_tol=5e-2)


def _case(fname):
    z = np.load(os.path.join(GOLD, fname))
    cfg = ast.literal_eval(str(z["cfg"]))
    arch = dict(n_layers=cfg["n_layers"], hidden=cfg["hidden"], add_linear=cfg.get("add_linear", False),
                bidirectional=cfg.get("bidirectional", False))
    return cfg, arch, {k: z[k] for k in KEYS}


def _product(cfg, arch, precision, device):
    from tepose_b200.synthetic import build_synthetic_vibe
    return build_synthetic_vibe(cfg["seed"], cfg["seqlen"], arch["n_layers"], arch["hidden"], arch["add_linear"],
                                arch["bidirectional"], True, precision, device)


def test_golden_fixtures_present():
    assert len(VIBE) == 3


@pytest.mark.parametrize("fname", VIBE)
def test_oracle_matches_reference_golden(fname):
    cfg, arch, gold = _case(fname)
    sd = synth.make_vibe_state_dict(cfg["seed"], **arch)
    m = torch_ref.SmplModel.synthetic(cfg["seed"])
    x = torch.from_numpy(synth.make_vibe_input(cfg["seed"], cfg["batch"], cfg["seqlen"]))
    out = torch_ref.vibe_forward(sd, m, x, J_regressor=m.J_regressor_h36m if cfg.get("use_h36m") else None, **arch)
    for k, g in gold.items():
        assert out[k].shape == g.shape
        np.testing.assert_allclose(out[k].numpy(), g, atol=2e-5, rtol=1e-5, err_msg=k)


@pytest.mark.parametrize("fname", VIBE)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_host_layer_against_reference_golden(fname, precision):
    cfg, arch, gold = _case(fname)
    model, sd = _product(cfg, arch, precision, "cpu")
    m = torch_ref.SmplModel.synthetic(cfg["seed"])
    x = torch.from_numpy(synth.make_vibe_input(cfg["seed"], cfg["batch"], cfg["seqlen"]))
    with fake_native.install() as fake:
        out = model(x, J_regressor=m.J_regressor_h36m if cfg.get("use_h36m") else None)[-1]
    compare_outputs(out, gold, label=fname, **({} if precision == "fp32" else BF16_TOL))
    kinds = [c[0] for c in fake.calls]
    assert kinds.count("gru") == arch["n_layers"]
    assert kinds.count("unpack_rows_residual") == 1


def test_state_dict_keys_match_reference_layout():
    from tepose_b200.synthetic import build_synthetic_vibe
    model, sd = build_synthetic_vibe(3, 4, 2, 32, add_linear=True)
    keys = {k for k in model.state_dict() if not k.startswith("regressor.smpl.")}
    assert keys == set(sd.keys())
    model, sd = build_synthetic_vibe(3, 4, 1, 32, add_linear=False, bidirectional=True)
    keys = {k for k in model.state_dict() if not k.startswith("regressor.smpl.")}
    assert keys == set(sd.keys()) and "encoder.gru.weight_hh_l0_reverse" in keys and "encoder.linear.weight" in keys


def test_no_residual_when_width_differs():
    """lib/models/vibe.py:60: without a linear layer and H != 2048 the encoder returns the GRU states."""
    from tepose_b200.vibe import TemporalEncoder
    enc = TemporalEncoder(n_layers=1, hidden_size=64)
    x = torch.from_numpy(synth.make_vibe_input(5, 2, 3))
    ref_gru = torch.nn.GRU(2048, 64)
    ref_gru.load_state_dict(enc.gru.state_dict())
    with torch.no_grad():
        want = ref_gru(x.permute(1, 0, 2))[0].permute(1, 0, 2)
    with fake_native.install():
        got = enc(x)
    assert got.shape == (2, 3, 64)
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=1e-5)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("fname", VIBE)
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gpu_forward_against_reference_golden(fname, precision):
    cfg, arch, gold = _case(fname)
    model, sd = _product(cfg, arch, precision, "cuda:0")
    m = torch_ref.SmplModel.synthetic(cfg["seed"])
    x = torch.from_numpy(synth.make_vibe_input(cfg["seed"], cfg["batch"], cfg["seqlen"])).cuda()
    out = model(x, J_regressor=m.J_regressor_h36m.cuda() if cfg.get("use_h36m") else None)[-1]
    errs = compare_outputs(out, gold, label=fname, **({} if precision == "fp32" else BF16_TOL))
    print(fname, precision, errs)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,T", [(1, 16), (5, 16)])
def test_gpu_released_bootstrap_config_against_oracle(B, T, precision):
    """evaluate.py:89-99: n_layers=2, hidden=1024, add_linear, residual; B=1 is the evaluate.py call,
    B=5 makes N = 80 regressor rows (more than one 64-row chunk of the skinny kernels)."""
    from tepose_b200.synthetic import build_synthetic_vibe
    seed = 60 + B
    model, sd = build_synthetic_vibe(seed, T, 2, 1024, True, False, True, precision, "cuda:0")
    m = torch_ref.SmplModel.synthetic(seed)
    x = synth.make_vibe_input(seed, B, T)
    ref = torch_ref.vibe_forward(sd, m, torch.from_numpy(x), 2, 1024, True, False, True, J_regressor=m.J_regressor_h36m)
    out = model(torch.from_numpy(x).cuda(), J_regressor=m.J_regressor_h36m.cuda())[-1]
    errs = compare_outputs(out, ref, label=f"vibe B{B}", **({} if precision == "fp32" else BF16_TOL))
    print(B, T, precision, errs)
    feat = model.encoder(torch.from_numpy(x).cuda())
    want = torch_ref.vibe_encoder_forward(sd, torch.from_numpy(x), 2, 1024, True, False, True)
    tol = 2e-4 if precision == "fp32" else 3e-2
    assert float((feat.cpu() - want).abs().max()) < tol

"""The windowed live loop (evaluate.py:229-269 / demo.py:229-252) kept on the device: oracle vs a golden run of
the unmodified reference modules, host logic on the emulated C ABI (CPU), CUDA path (GPU)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import synth, torch_ref
from tests import fake_native
from tests.helpers import build_product_model, compare_outputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")
STREAMS = sorted(f for f in os.listdir(GOLD) if f.startswith("stream_"))
KEYS = ("theta", "verts", "kp_2d", "kp_3d", "rotmat")
BF16_TOL = dict(vert_tol=1e-3, rot_tol=2e-3, kp2d_tol=5e-2)


def _case(fname):
    z = np.load(os.path.join(GOLD, fname))
    cfg = ast.literal_eval(str(z["cfg"]))
    return cfg, {k: z[k] for k in KEYS}


def _thin(out):
    """The fixture keeps every 4th frame's mesh."""
    out = dict(out)
    out["verts"] = out["verts"][:, ::4]
    return out


def _models(cfg, precision, device):
    from tepose_b200.synthetic import build_synthetic_vibe
    model, sd = build_product_model(cfg["seed"], cfg["seqlen"], cfg["n_layers"], cfg["hidden"], precision, device)
    vibe, sd_v = build_synthetic_vibe(cfg["seed"], cfg["seqlen"], cfg["vibe_layers"], cfg["vibe_hidden"], True, False, True,
                                      precision, device)
    return model, sd, vibe, sd_v


ARCH = dict(add_linear=True, bidirectional=False, use_residual=True)


@pytest.mark.parametrize("fname", STREAMS)
def test_oracle_loop_matches_reference_golden(fname):
    cfg, gold = _case(fname)
    sd_v = synth.make_vibe_state_dict(cfg["seed"], cfg["vibe_layers"], cfg["vibe_hidden"], True, False)
    sd = synth.make_state_dict(cfg["seed"], cfg["n_layers"], cfg["hidden"])
    m = torch_ref.SmplModel.synthetic(cfg["seed"])
    feats = torch.from_numpy(synth.make_vibe_input(cfg["seed"], cfg["batch"], cfg["frames"]))
    out = torch_ref.windowed_stream(sd_v, dict(n_layers=cfg["vibe_layers"], hidden=cfg["vibe_hidden"], **ARCH), sd, m, feats,
                                    cfg["seqlen"], cfg["n_layers"], cfg["hidden"], J_regressor=m.J_regressor_h36m)
    out = _thin(out)
    for k, g in gold.items():
        assert out[k].shape == g.shape, k
        np.testing.assert_allclose(out[k].numpy(), g, atol=3e-5, rtol=1e-5, err_msg=k)


@pytest.mark.parametrize("fname", STREAMS)
def test_host_loop_against_reference_golden(fname):
    from tepose_b200.stream import WindowedTePose
    cfg, gold = _case(fname)
    model, sd, vibe, sd_v = _models(cfg, "fp32", "cpu")
    m = torch_ref.SmplModel.synthetic(cfg["seed"])
    feats = torch.from_numpy(synth.make_vibe_input(cfg["seed"], cfg["batch"], cfg["frames"]))
    with fake_native.install():
        ws = WindowedTePose(model, vibe, J_regressor=m.J_regressor_h36m, batch=cfg["batch"], use_graph=False)
        out = ws.run(feats)
    compare_outputs(_thin(out), gold, label=fname)


def test_ring_seeded_from_given_thetas_and_batched_streams():
    """evaluate.py:219 seeds the ring from the dataset's pseudo thetas; B streams advance side by side."""
    from tepose_b200.stream import WindowedTePose
    seed, T, N, B = 33, 3, 6, 2
    cfg = dict(seed=seed, seqlen=T, n_layers=1, hidden=32, vibe_layers=1, vibe_hidden=32)
    model, sd, vibe, sd_v = _models(cfg, "fp32", "cpu")
    m = torch_ref.SmplModel.synthetic(seed)
    feats = torch.from_numpy(synth.make_vibe_input(seed, B, N))
    theta0 = 0.1 * torch.randn(T - 1, 85, generator=torch.Generator().manual_seed(3))
    ref = torch_ref.windowed_stream(sd_v, dict(n_layers=1, hidden=32, **ARCH), sd, m, feats, T, 1, 32, theta_input=theta0)
    with fake_native.install():
        ws = WindowedTePose(model, vibe, batch=B, use_graph=False)
        with pytest.raises(RuntimeError):
            ws.step(feats[:, :T])                       # ring not seeded
        out = ws.run(feats, theta_input=theta0)
        with pytest.raises(ValueError):
            ws.step(feats[:, :T + 1])
    compare_outputs(out, ref, label="seeded ring")


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("fname", STREAMS)
@pytest.mark.parametrize("precision,graph", [("fp32", False), ("fp32", True), ("bf16", True)])
def test_gpu_loop_against_reference_golden(fname, precision, graph):
    from tepose_b200.stream import WindowedTePose
    cfg, gold = _case(fname)
    model, sd, vibe, sd_v = _models(cfg, precision, "cuda:0")
    m = torch_ref.SmplModel.synthetic(cfg["seed"])
    feats = torch.from_numpy(synth.make_vibe_input(cfg["seed"], cfg["batch"], cfg["frames"])).cuda()
    ws = WindowedTePose(model, vibe, J_regressor=m.J_regressor_h36m, batch=cfg["batch"], use_graph=graph)
    out = ws.run(feats)
    errs = compare_outputs(_thin(out), gold, label=fname, **({} if precision == "fp32" else BF16_TOL))
    print(fname, precision, graph, errs)


@pytest.mark.gpu
def test_gpu_loop_full_width_against_oracle():
    """Default widths (TePose L1 H2048 T16; VIBE L2 H1024), 2 streams, 20 frames, graphs on, fp32."""
    from tepose_b200.stream import WindowedTePose
    seed, T, N, B = 35, 16, 20, 2
    cfg = dict(seed=seed, seqlen=T, n_layers=1, hidden=2048, vibe_layers=2, vibe_hidden=1024)
    model, sd, vibe, sd_v = _models(cfg, "fp32", "cuda:0")
    m = torch_ref.SmplModel.synthetic(seed)
    feats = torch.from_numpy(synth.make_vibe_input(seed, B, N))
    ref = torch_ref.windowed_stream(sd_v, dict(n_layers=2, hidden=1024, **ARCH), sd, m, feats, T, 1, 2048,
                                    J_regressor=m.J_regressor_h36m)
    ws = WindowedTePose(model, vibe, J_regressor=m.J_regressor_h36m, batch=B)
    out = ws.run(feats.cuda())
    errs = compare_outputs(out, ref, label="stream full width")
    print(errs)
    again = ws.run(feats.cuda())                         # the ring is re-seeded per run: identical results
    for k in out:
        assert torch.equal(out[k], again[k]), k

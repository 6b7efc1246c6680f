"""Evaluation metrics (lib/utils/eval_utils.py) on the device: oracle vs golden outputs of the unmodified reference
functions, host logic on the emulated C ABI (CPU), CUDA kernels (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_ref
from oracle.make_golden import metric_inputs
from tests import fake_native

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics.npz"))
T = lambda a: torch.from_numpy(np.asarray(a).copy())


def test_oracle_metrics_match_reference_golden():
    d = metric_inputs()
    m = torch_ref.pose_metrics(T(d["pred"]), T(d["target"]), (2, 3))
    np.testing.assert_allclose(m["mpjpe"].numpy(), GOLD["mpjpe"], atol=1e-6)
    np.testing.assert_allclose(m["aligned"].numpy(), GOLD["aligned"], atol=2e-5)
    np.testing.assert_allclose(m["mpjpe_pa"].numpy(), GOLD["mpjpe_pa"], atol=2e-5)
    np.testing.assert_allclose(m["accel_err"][1:-1].numpy(), GOLD["accel_err"], atol=1e-6)
    np.testing.assert_allclose(m["mpjpe"].numpy(), GOLD["errors"], atol=1e-6)            # compute_errors == the batched path
    np.testing.assert_allclose(m["mpjpe_pa"].numpy(), GOLD["errors_pa"], atol=2e-5)
    np.testing.assert_allclose(torch_ref.procrustes_align(T(d["pred17"]), T(d["target17"])).numpy(), GOLD["aligned17"], atol=2e-5)
    vl = T(d["vidlen"])
    np.testing.assert_allclose(float(torch_ref.accel_summary(torch_ref.accel_error(T(d["seq_p"])), vl, 4, 2, 1)), GOLD["accel_seq"], rtol=1e-5)
    np.testing.assert_allclose(float(torch_ref.accel_summary(torch_ref.accel_error(T(d["seq_p"]), T(d["seq_g"])), vl, 4, 4, 3)),
                               GOLD["accel_err_seq"], rtol=1e-5)
    np.testing.assert_allclose(torch_ref.vertex_error(T(d["va"]), T(d["vb"])).numpy(), GOLD["mpvpe"], rtol=1e-5)
    # float64 evaluation of the same formulas agrees with the fp32 reference run
    m64 = torch_ref.pose_metrics(T(d["pred"]).double(), T(d["target"]).double(), (2, 3))
    np.testing.assert_allclose(m64["aligned"].numpy(), GOLD["aligned"], atol=2e-5)


def _check_against_golden(dev):
    from tepose_b200 import eval_utils as eu
    d = metric_inputs()
    to = lambda a: T(a).to(dev)
    m = eu.pose_metrics(to(d["pred"]), to(d["target"]), pelvis=(2, 3), want_aligned=True)
    np.testing.assert_allclose(m["mpjpe"].cpu().numpy(), GOLD["mpjpe"], atol=1e-6)
    np.testing.assert_allclose(m["aligned"].cpu().numpy(), GOLD["aligned"], atol=2e-5)
    np.testing.assert_allclose(m["mpjpe_pa"].cpu().numpy(), GOLD["mpjpe_pa"], atol=2e-5)
    np.testing.assert_allclose(m["accel_err"][1:-1].cpu().numpy(), GOLD["accel_err"], atol=1e-6)
    assert float(m["accel_err"][0]) == 0.0 and float(m["accel_err"][-1]) == 0.0
    e, epa = eu.compute_errors(to(d["target"]), to(d["pred"]))
    np.testing.assert_allclose(e.cpu().numpy(), GOLD["errors"], atol=1e-6)
    np.testing.assert_allclose(epa.cpu().numpy(), GOLD["errors_pa"], atol=2e-5)
    hat = eu.batch_compute_similarity_transform_torch(to(d["pred17"]), to(d["target17"]))
    np.testing.assert_allclose(hat.cpu().numpy(), GOLD["aligned17"], atol=2e-5)
    # evaluate.py:424-442 call pattern: align on the host side, then the un-aligned entry points
    P, G = to(d["pred"]), to(d["target"])
    P = P - (P[:, [2]] + P[:, [3]]) / 2.0
    G = G - (G[:, [2]] + G[:, [3]]) / 2.0
    np.testing.assert_allclose(eu.compute_error_accel_eval(G, P).cpu().numpy(), GOLD["accel_err"], atol=1e-6)
    vis = np.ones(40, dtype=bool); vis[[3, 20, 21]] = False
    np.testing.assert_allclose(eu.compute_error_accel_eval(G, P, vis=vis).cpu().numpy(), GOLD["accel_err_vis"], atol=1e-6)
    vl = T(d["vidlen"])
    np.testing.assert_allclose(float(eu.compute_accel(to(d["seq_p"]), vl, 4)), GOLD["accel_seq"], rtol=1e-5)
    np.testing.assert_allclose(float(eu.compute_error_accel(to(d["seq_g"]), to(d["seq_p"]), vl, 4)), GOLD["accel_err_seq"], rtol=1e-5)
    np.testing.assert_allclose(eu.compute_error_verts(pred_verts=to(d["va"]), target_verts=to(d["vb"])).cpu().numpy(), GOLD["mpvpe"], rtol=1e-5)


def test_host_layer_against_reference_golden():
    with fake_native.install():
        _check_against_golden("cpu")


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_metrics_against_reference_golden():
    _check_against_golden("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("n,J,pelvis", [(1, 14, (2, 3)), (257, 14, (2, 3)), (100, 17, -3), (33, 49, None), (5000, 14, (2, 3))])
def test_gpu_pose_metrics_against_float64_oracle(n, J, pelvis):
    from tepose_b200 import eval_utils as eu
    g = torch.Generator().manual_seed(n + J)
    target = torch.randn(n, J, 3, generator=g) * 0.4
    pred = target + 0.08 * torch.randn(n, J, 3, generator=g)
    ref = torch_ref.pose_metrics(pred.double(), target.double(), (J + pelvis) if isinstance(pelvis, int) else pelvis)
    got = eu.pose_metrics(pred.cuda(), target.cuda(), pelvis=pelvis, want_aligned=True)
    for k in ("mpjpe", "mpjpe_pa", "accel_err", "aligned"):
        err = float((got[k].cpu().double() - ref[k]).abs().max())
        assert err < 2e-6, (k, err)


@pytest.mark.gpu
def test_gpu_procrustes_is_invariant_to_similarity_transforms():
    """Size-independent property: aligning s R x + t onto x returns x (PA error 0), for random proper rotations."""
    from tepose_b200 import eval_utils as eu
    g = torch.Generator().manual_seed(5)
    n = 4096
    x = torch.randn(n, 14, 3, generator=g)
    R = torch_ref.batch_rodrigues_smplx(torch.randn(n, 3, generator=g))
    s = 0.5 + torch.rand(n, 1, 1, generator=g)
    y = s * torch.einsum("nij,nkj->nki", R, x) + torch.randn(n, 1, 3, generator=g)
    m = eu.pose_metrics(y.cuda(), x.cuda(), pelvis=None, want_aligned=True)
    assert float(m["mpjpe_pa"].max()) < 5e-6
    assert float((m["aligned"].cpu() - x).abs().max()) < 2e-5


@pytest.mark.gpu
def test_gpu_mpvpe_from_target_theta():
    """compute_error_verts(target_theta=...): target mesh from the SMPL forward (pose2rot=True), then the vertex error."""
    from tepose_b200 import eval_utils as eu
    from tepose_b200 import synthetic as synth
    from tests.helpers import build_product_model
    model, _ = build_product_model(71, 4, 1, 64, "fp32", "cuda:0")
    m = torch_ref.SmplModel.synthetic(71)
    b = synth.make_bodies(71, 37)
    theta = torch.cat([torch.from_numpy(b["cam"]), torch.from_numpy(b["pose_aa"]), torch.from_numpy(b["betas"])], dim=1)
    tv, _, _ = torch_ref.smpl_forward(m, torch.from_numpy(b["betas"]), pose_aa=torch.from_numpy(b["pose_aa"]))
    pred = tv + 0.01 * torch.randn(tv.shape, generator=torch.Generator().manual_seed(1))
    want = torch_ref.vertex_error(tv, pred)
    got = eu.compute_error_verts(pred_verts=pred.cuda(), target_theta=theta.cuda(), smpl=model.regressor.smpl)
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=2e-4)

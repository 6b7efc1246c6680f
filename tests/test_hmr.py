"""HMR ResNet-50 feature extractor + regressor (lib/models/spin.py:16-204, caller demo.py:183-198).
CPU: state_dict keys and the BatchNorm fold against the unmodified reference's key list / torch's own conv + BN.
GPU: data-movement kernels bit-exact against torch, the GEMM epilogue options, and the whole extractor / forward against
tests/golden/hmr_N2.npz (outputs of the unmodified reference HMR with tepose_b200.synthetic.make_hmr_state weights)."""
import os

import numpy as np
import pytest
import torch

from tepose_b200 import synthetic as psynth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hmr_N2.npz")
DEV = "cuda:0"


def test_hmr_state_dict_keys_match_the_reference():
    z = np.load(GOLD)
    model = psynth.build_synthetic_hmr(11, "cpu")
    ours = sorted(k for k in model.state_dict() if not k.startswith("smpl."))
    assert ours == [str(k) for k in z["keys"]]


def test_batchnorm_fold_matches_conv_then_bn():
    from tepose_b200.hmr import _fold
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(8, 16, 3, stride=2, padding=1, bias=False)
    bn = torch.nn.BatchNorm2d(16).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    x = torch.randn(2, 8, 10, 10)
    with torch.no_grad():
        want = bn(conv(x))
    w, b = _fold(conv, bn)
    cols = torch.nn.functional.unfold(x, 3, padding=1, stride=2)               # [N, C*9, L], (c, ky, kx) order
    cols = cols.view(2, 8, 9, -1).permute(0, 3, 2, 1).reshape(2, -1, 72)         # -> (ky, kx, c) columns
    got = cols @ w.float()[:, :72].T + b
    got = got.permute(0, 2, 1).reshape(want.shape)
    assert float((got - want).abs().max()) < 0.03 * float(want.abs().max())      # weights rounded to bf16
    assert w.shape == (16, 128) and float(w[:, 72:].abs().max()) == 0.0


def test_hmr_oracle_restatement_matches_the_reference_golden():
    """oracle/hmr_ref.py (functional fp32 restatement of lib/models/spin.py:16-141) against the unmodified reference's features."""
    from oracle import hmr_ref
    z = np.load(GOLD)
    model = psynth.build_synthetic_hmr(11, "cpu")
    sd = {k: v.float() for k, v in model.state_dict().items() if v.is_floating_point()}
    with torch.no_grad():
        xf = hmr_ref.feature_extractor(sd, torch.from_numpy(psynth.make_image_batch(11, 2)))
    assert float((xf - torch.from_numpy(z["xf"])).abs().max()) < 1e-4 * float(np.abs(z["xf"]).max())


def test_hmr_host_schedule_against_reference_golden_on_cpu():
    """The Python layer of HMR.feature_extractor (BatchNorm fold, weight column order, block wiring, shortcut / stride handling)
    with the native calls emulated on CPU (tests/fake_native.py: bf16 activations, fp32 accumulation) against the unmodified
    reference's features for the first golden crop."""
    from tests import fake_native
    z = np.load(GOLD)
    model = psynth.build_synthetic_hmr(11, "cpu")
    x = torch.from_numpy(psynth.make_image_batch(11, 2))[:1]
    with fake_native.install() as fake, torch.no_grad():
        xf = model.feature_extractor(x)
    ref = torch.from_numpy(z["xf"])[:1]
    rel = float((xf - ref).norm() / ref.norm())
    assert xf.shape == (1, 2048) and rel < 3e-2, rel
    assert sum(1 for c in fake.calls if c[0] == "gemm_bf16_tc") == 53 and sum(1 for c in fake.calls if c[0] == "im2col") == 20


@pytest.mark.gpu
def test_conv_data_movement_kernels_are_exact():
    from tepose_b200 import _native as nv
    L = nv.lib()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 3, 20, 18, generator=g).to(DEV)
    y = torch.empty(3, 20, 18, 4, device=DEV, dtype=torch.bfloat16)
    nv.check(L.tp_nchw_to_nhwc_bf16(nv.ptr(x), nv.ptr(y), 3, 3, 20, 18, 4, nv.stream()))
    want = torch.zeros(3, 20, 18, 4, device=DEV, dtype=torch.bfloat16)
    want[..., :3] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(y, want)
    for (C, kh, s, pad) in [(4, 7, 2, 3), (16, 3, 1, 1), (16, 3, 2, 1), (24, 1, 2, 0)]:
        a = torch.randn(2, 13, 11, C, generator=g).to(DEV).to(torch.bfloat16)
        Ho, Wo = (13 + 2 * pad - kh) // s + 1, (11 + 2 * pad - kh) // s + 1
        K = kh * kh * C
        KP = (K + 63) // 64 * 64
        out = torch.full((2 * Ho * Wo, KP), 7.0, device=DEV, dtype=torch.bfloat16)
        nv.check(L.tp_im2col_nhwc_bf16(nv.ptr(a), nv.ptr(out), 2, 13, 11, C, kh, kh, s, pad, KP, nv.stream()))
        cols = torch.nn.functional.unfold(a.float().permute(0, 3, 1, 2), kh, padding=pad, stride=s)      # [N, C*kh*kh, L]
        cols = cols.view(2, C, kh * kh, -1).permute(0, 3, 2, 1).reshape(2 * Ho * Wo, K)
        assert torch.equal(out[:, :K].float(), cols), (C, kh, s, pad)
        assert float(out[:, K:].float().abs().max()) == 0.0 if KP > K else True
    a = torch.randn(2, 14, 14, 16, generator=g).to(DEV).to(torch.bfloat16)
    out = torch.empty(2, 7, 7, 16, device=DEV, dtype=torch.bfloat16)
    nv.check(L.tp_maxpool3x3s2_nhwc_bf16(nv.ptr(a), nv.ptr(out), 2, 14, 14, 16, nv.stream()))
    want = torch.nn.functional.max_pool2d(a.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(out.float(), want)
    avg = torch.empty(2, 16, device=DEV)
    nv.check(L.tp_avgpool_nhwc_bf16(nv.ptr(out), nv.ptr(avg), 2, 49, 16, nv.stream()))
    assert float((avg - out.float().view(2, 49, 16).mean(1)).abs().max()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cout,kp", [(300, 64, 64), (1000, 256, 576), (98, 2048, 512)])
def test_gemm_epilogue_relu_residual_bf16(rows, cout, kp):
    """tp_gemm_seg flags: bf16 output, ReLU, residual add (the bottleneck's conv3 + shortcut + ReLU, lib/models/spin.py:46-54)."""
    from tepose_b200 import _native as nv
    g = torch.Generator().manual_seed(rows)
    A = torch.randn(rows, kp, generator=g).to(torch.bfloat16)
    W = (torch.randn(cout, kp, generator=g) / kp ** 0.5).to(torch.bfloat16)
    bias = torch.randn(cout, generator=g)
    R = torch.randn(rows, cout, generator=g).to(torch.bfloat16)
    a, w, b, r = A.to(DEV), W.to(DEV), bias.to(DEV), R.to(DEV)
    full = A.double() @ W.double().T + bias.double()
    for relu, res in [(True, True), (False, False), (True, False)]:
        out = torch.full((rows, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
        seg = (nv.GemmSeg * 1)()
        seg[0] = nv.GemmSeg(0, rows, 0, cout, nv.ptr(out), cout, nv.ptr(b), (nv.GEMM_RELU if relu else 0) | nv.GEMM_OUT_BF16,
                            cout if res else 0, nv.ptr(r) if res else None)
        nv.check(nv.lib().tp_gemm_bf16_tc(nv.ptr(a), rows, nv.ptr(w), cout, kp, seg, 1, nv.stream()))
        torch.cuda.synchronize()
        want = full + (R.double() if res else 0)
        if relu:
            want = want.clamp_min(0)
        err = float((out.cpu().double() - want).abs().max())
        assert err < 0.02 * float(want.abs().max()) + 1e-3, (relu, res, err)      # one bf16 rounding of the output


@pytest.mark.gpu
def test_hmr_against_reference_golden():
    z = np.load(GOLD)
    model = psynth.build_synthetic_hmr(11, DEV)
    x = torch.from_numpy(psynth.make_image_batch(11, 2)).to(DEV)
    with torch.no_grad():
        xf = model.feature_extractor(x)
        xf2, out = model(x, return_features=True)
    torch.cuda.synchronize()
    assert xf.shape == (2, 2048) and torch.equal(xf, xf2)
    ref = torch.from_numpy(z["xf"])
    rel = float((xf.cpu() - ref).norm() / ref.norm())
    cos = float(torch.nn.functional.cosine_similarity(xf.cpu().flatten(), ref.flatten(), dim=0))
    # bf16 activations and weights through 53 convolutions, fp32 accumulation: a few 1e-3 relative per layer
    assert rel < 3e-2 and cos > 0.9995, (rel, cos)
    assert set(out[0]) == {"theta", "verts", "kp_2d", "kp_3d"}
    for k, tol in (("verts", 2e-3), ("kp_3d", 2e-3), ("kp_2d", 2e-2)):
        err = float((out[0][k].cpu() - torch.from_numpy(z[k])).abs().max())
        assert err < tol, (k, err)
    # another batch size / other crops: against the oracle restatement (pinned to the same golden on CPU)
    from oracle import hmr_ref
    x5 = torch.from_numpy(psynth.make_image_batch(5, 5))
    sd = {k: v.float().cpu() for k, v in model.state_dict().items() if v.is_floating_point()}
    with torch.no_grad():
        want5 = hmr_ref.feature_extractor(sd, x5)
        got5 = model.feature_extractor(x5.to(DEV)).cpu()
    assert float((got5 - want5).norm() / want5.norm()) < 3e-2
    from tepose_b200.graph import GraphedHMRFeatures
    g = GraphedHMRFeatures(model, 2)
    assert torch.equal(g(x), xf) and g.launches_per_replay > 50
    with pytest.raises(NotImplementedError):
        model.train().feature_extractor(x)

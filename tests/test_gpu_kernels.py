"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against the CPU oracle.
Tolerances are written next to each check."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import tepose_b200
import tepose_b200._native as nv
from oracle import synth, torch_ref
from tests.helpers import base_data_cwd

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def cu(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a)).to(DEV, dtype).contiguous()


# ------------------------------------------------------------------ geometry
def test_geometry_against_reference_golden():
    z = np.load(os.path.join(GOLD, "geometry.npz"))
    R = tepose_b200.rot6d_to_rotmat(cu(z["x6"])).cpu().numpy()
    # row 3 has a2 parallel to a1: b2 is normalised rounding noise in the reference itself, so only
    # its first column is comparable; every other row (incl. the two below-eps rows) must agree.
    keep = np.array([i for i in range(len(R)) if i != 3])
    np.testing.assert_allclose(R[keep], z["rot6d"][keep], atol=2e-6)
    np.testing.assert_allclose(R[3][:, 0], z["rot6d"][3][:, 0], atol=2e-6)
    aa = tepose_b200.rotation_matrix_to_angle_axis(cu(z["R"])).cpu()
    ref = torch.from_numpy(z["r2aa"])
    # compare as rotations (axis-angle is ill-conditioned near pi), and directly away from pi
    Ra = torch_ref.batch_rodrigues_smplx(aa)
    Rb = torch_ref.batch_rodrigues_smplx(ref)
    assert float((Ra - Rb).abs().max()) < 5e-5
    small = ref.norm(dim=1) < 3.0
    assert float((aa[small] - ref[small]).abs().max()) < 2e-5
    q = tepose_b200.batch_rodrigues(cu(z["aa"])).cpu().numpy().reshape(-1, 3, 3)
    np.testing.assert_allclose(q, z["rod_q"], atol=2e-6)
    s = tepose_b200.batch_rodrigues(cu(z["aa"]), form="smplx").cpu()
    assert float((s - torch_ref.batch_rodrigues_smplx(torch.from_numpy(z["aa"]))).abs().max()) < 2e-6
    kp = tepose_b200.projection(cu(z["joints"]), cu(z["cam"])).cpu().numpy()
    np.testing.assert_allclose(kp, z["proj"], rtol=2e-6, atol=2e-4)


def test_geometry_empty_input():
    out = tepose_b200.rot6d_to_rotmat(torch.empty(0, 6, device=DEV))
    assert out.shape == (0, 3, 3)


# ------------------------------------------------------------------ GEMMs
@pytest.mark.parametrize("M,N,K", [(1, 64, 64), (32, 2048, 2048), (32, 160, 1024), (64, 1024, 160),
                                   (70, 100, 36), (512, 384, 2176), (130, 72, 8)])
def test_gemm_f32(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    Cin = torch.randn(M, N, generator=g)
    ref = 0.5 * (A.clamp_min(0).double() @ W.double().t() + b.double()) + 2.0 * Cin.double()
    a, w, bb, cin = cu(A), cu(W), cu(b), cu(Cin)
    out = torch.empty(M, N, device=DEV)
    nv.check(nv.lib().tp_gemm_f32(nv.ptr(a), K, nv.ptr(w), K, nv.ptr(bb), nv.ptr(cin), N, nv.ptr(out), N,
                                  M, N, K, 0.5, 2.0, 1, nv.stream()))
    err = float((out.cpu().double() - ref).abs().max())
    assert err < 2e-5, err                       # fp32 accumulate over K <= 2176 terms of O(1/sqrt(K))
    # in-place accumulate (Cin aliases C), no bias, no relu
    out2 = cin.clone()
    nv.check(nv.lib().tp_gemm_f32(nv.ptr(a), K, nv.ptr(w), K, nv.vp(0), nv.ptr(out2), N, nv.ptr(out2), N,
                                  M, N, K, 1.0, 1.0, 0, nv.stream()))
    ref2 = A.double() @ W.double().t() + Cin.double()
    assert float((out2.cpu().double() - ref2).abs().max()) < 2e-5


@pytest.mark.parametrize("M,N,K,splits", [(32, 2048, 4096, 8), (64, 1024, 1024, 4), (32, 160, 1024, 16), (5, 96, 160, 3)])
def test_gemm_f32_splitk(M, N, K, splits):
    g = torch.Generator().manual_seed(M + N + K)
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    b, Cin = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    a, w, bb, c = cu(A), cu(W), cu(b), cu(Cin)
    L = nv.lib()
    ws = torch.zeros(L.tp_gemm_f32_splitk_workspace_bytes(M, N, splits), dtype=torch.uint8, device=DEV)
    ref = 0.5 * (A.clamp_min(0).double() @ W.double().t() + b.double()) + Cin.double()
    outs = []
    for _ in range(3):   # tickets must reset themselves; result must be bitwise reproducible
        out = c.clone()
        nv.check(L.tp_gemm_f32_splitk(nv.ptr(a), K, nv.ptr(w), K, nv.ptr(bb), nv.ptr(out), N, nv.ptr(out), N, M, N, K,
                                      0.5, 1.0, 1, splits, nv.ptr(ws), ws.numel(), nv.stream()))
        outs.append(out.cpu())
    assert float((outs[0].double() - ref).abs().max()) < 2e-5
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0


@pytest.mark.parametrize("M,N,K,splits", [(32, 2048, 4096, 9), (32, 2048, 2048, 1), (64, 1024, 1024, 8), (32, 160, 1024, 8),
                                          (1, 18432, 2176, 1), (7, 1024, 160, 1), (32, 1000, 100, 2)])
def test_skinny_bf16(M, N, K, splits):
    g = torch.Generator().manual_seed(M + N + K)
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    b, Cin = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    a, bb, c = cu(A), cu(b), cu(Cin)
    L = nv.lib()
    wp = nv.pack_linear(cu(W), "bf16")
    assert wp.numel() == L.tp_pack_mma_a_bytes(N, K)
    ws = torch.zeros(L.tp_skinny_bf16_workspace_bytes(M, N, splits), dtype=torch.uint8, device=DEV)
    Ab = A.clamp_min(0).to(torch.bfloat16).double()
    ref = 0.5 * (Ab @ W.to(torch.bfloat16).double().t() + b.double()) + Cin.double()
    outs = []
    for _ in range(2):
        out = c.clone()
        nv.check(L.tp_skinny_bf16(nv.ptr(a), K, M, K, nv.ptr(wp), N, nv.ptr(bb), nv.ptr(out), N, nv.ptr(out), N, 0.5, 1.0, 1,
                                  splits, nv.ptr(ws), ws.numel(), nv.stream()))
        outs.append(out.cpu())
    err = float((outs[0].double() - ref).abs().max())
    assert err < 1e-4, err                        # exact bf16 products; fp32 accumulation order only
    assert torch.equal(outs[0], outs[1])
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0


@pytest.mark.parametrize("M,N,K", [(32, 1024, 1024), (32, 160, 1024), (32, 1024, 160), (64, 1024, 1024), (32, 1024, 2048), (3, 96, 64)])
def test_skinny_bf16_single_phase_and_bf16_io(M, N, K):
    g = torch.Generator().manual_seed(3 * M + N + K)
    A, W = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5
    b, Cin = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    a, bb, c = cu(A), cu(b), cu(Cin)
    a_lp = A.to(torch.bfloat16).to(DEV)
    L = nv.lib()
    wp = nv.pack_linear(cu(W), "bf16")
    ref = (A.to(torch.bfloat16).double() @ W.to(torch.bfloat16).double().t() + b.double()) + Cin.double()
    for use_lp in (False, True):
        out = c.clone()
        out_lp = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
        nv.check(L.tp_skinny_bf16_ex(nv.vp(0) if use_lp else nv.ptr(a), K, nv.ptr(a_lp) if use_lp else nv.vp(0), K, M, K,
                                     nv.ptr(wp), N, nv.ptr(bb), nv.ptr(out), N, nv.ptr(out), N, nv.ptr(out_lp), N,
                                     1.0, 1.0, 0, 1, 2, nv.vp(0), 0, nv.stream()))
        err = float((out.cpu().double() - ref).abs().max())
        assert err < 1e-4, (use_lp, err)
        assert torch.equal(out_lp.cpu(), out.cpu().to(torch.bfloat16))


def test_pack_rows():
    x = torch.randn(3, 5, 2133)
    xs = cu(x)
    for prec, dt in ((nv.PRECISION_FP32, torch.float32), (nv.PRECISION_BF16, torch.bfloat16)):
        dst = torch.full((15, 2176), 7.0, device=DEV, dtype=dt)
        nv.check(nv.lib().tp_pack_rows(nv.ptr(xs), xs.stride(0), xs.stride(1), 3, 5, 2133, nv.ptr(dst), 2176, prec, 0,
                                       nv.stream()))
        ref = torch.zeros(15, 2176)
        ref[:, :2133] = x.permute(1, 0, 2).reshape(15, 2133)
        assert torch.equal(dst.cpu(), ref.to(dt))          # bit-exact: a copy + RN conversion


def test_unpack_rows_residual():
    """y [T*B, ld] time-major + x [B,T,F] (a strided view) -> [B*T, F]; bit-exact (one fp32 add, RN conversion)."""
    B, T, F, ld = 3, 5, 2048, 2112
    y = torch.randn(T * B, ld)
    xbig = torch.randn(B, 2 * T, F)
    xs = cu(xbig)[:, ::2]
    ys = cu(y)
    for with_x in (True, False):
        out = torch.full((B * T, F), 7.0, device=DEV)
        out_lp = torch.full((B * T, F), 7.0, device=DEV, dtype=torch.bfloat16)
        nv.check(nv.lib().tp_unpack_rows_residual(nv.ptr(ys), ld, nv.vp(xs.data_ptr() if with_x else 0), xs.stride(0),
                                                  xs.stride(1), B, T, F, nv.ptr(out), nv.ptr(out_lp), nv.stream()))
        ref = y[:, :F].reshape(T, B, F).permute(1, 0, 2)
        if with_x:
            ref = ref + xbig[:, ::2]
        ref = ref.reshape(B * T, F)
        assert torch.equal(out.cpu(), ref)
        assert torch.equal(out_lp.cpu(), ref.to(torch.bfloat16))


@pytest.mark.parametrize("rows,wrows,kp", [(8, 192, 128), (512, 768, 2176), (130, 384, 64), (32, 96, 192),
                                            (512, 9216, 256), (768, 9216, 128), (300, 9216, 64),    # > 1 wave: 192-wide tiles, CTA pairs sharing W / A / nothing
                                            (544, 18432, 2176), (512, 1152, 192)])
def test_gemm_bf16_tcgen05(rows, wrows, kp):
    g = torch.Generator().manual_seed(rows + wrows)
    A = torch.randn(rows, kp, generator=g).to(torch.bfloat16)
    W = (torch.randn(wrows, kp, generator=g) / kp ** 0.5).to(torch.bfloat16)
    bias = torch.randn(wrows, generator=g)
    a, w, b = A.to(DEV), W.to(DEV), cu(bias)
    third = wrows // 3
    m_tail = max(1, rows // 4)
    segs = [(0, rows, 0, 2 * third), (rows - m_tail, m_tail, 2 * third, third)]
    outs = [torch.full((mr, nc), float("nan"), device=DEV) for (_, mr, _, nc) in segs]
    arr = (nv.GemmSeg * 2)()
    for i, ((m0, mr, n0, nc), o) in enumerate(zip(segs, outs)):
        arr[i] = nv.GemmSeg(m0, mr, n0, nc, nv.ptr(o), nc, nv.vp(b.data_ptr() + 4 * n0))
    nv.check(nv.lib().tp_gemm_bf16_tc(nv.ptr(a), rows, nv.ptr(w), wrows, kp, arr, 2, nv.stream()))
    torch.cuda.synchronize()
    full = A.double() @ W.double().t() + bias.double()
    for (m0, mr, n0, nc), o in zip(segs, outs):
        ref = full[m0:m0 + mr, n0:n0 + nc]
        err = float((o.cpu().double() - ref).abs().max())
        assert err < 1e-4, (err, (m0, mr, n0, nc))       # exact bf16 products, fp32 accumulation order only


def test_gemm_bf16_tcgen05_cta_pairs():
    """TP_TC_2CTA=1: the same GEMM cases on the tcgen05.mma.cta_group::2 kernel (k_gemm_bf16_tc2, opt-in: slower at these sizes).
    The switch is read once per process, so the cases run in a child process."""
    import subprocess, sys
    env = dict(os.environ, TP_TC_2CTA="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "--no-header", "-p", "no:cacheprovider",
                        "-k", "test_gemm_bf16_tcgen05 and not cta_pairs"], env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


# ------------------------------------------------------------------ GRU recurrence
def _gru_case(B, T, H, precision, seed, with_h0=False, reverse=False):
    g = torch.Generator().manual_seed(seed)
    k = 1.0 / H ** 0.5
    u = lambda *s: (torch.rand(*s, generator=g) * 2 - 1) * k
    w_hh, b_hh = u(3 * H, H), u(3 * H)
    gi = torch.randn(T, B, 3 * H, generator=g) * 0.5
    h0 = torch.randn(B, H, generator=g) * 0.3 if with_h0 else None
    wdt = torch.bfloat16 if precision == "bf16" else torch.float32
    # oracle: the GRU cell in float64 (torch.nn.GRU semantics)
    Wd = w_hh.to(wdt).double()
    h = torch.zeros(B, H, dtype=torch.float64) if h0 is None else h0.double()
    ys = torch.zeros(T, B, H, dtype=torch.float64)
    for s in range(T):
        t = T - 1 - s if reverse else s
        hm = h.float().to(wdt).double() if precision == "bf16" else h
        gh = hm @ Wd.t() + b_hh.double()
        gx = gi[t].double()
        r = torch.sigmoid(gx[:, :H] + gh[:, :H])
        z = torch.sigmoid(gx[:, H:2 * H] + gh[:, H:2 * H])
        n = torch.tanh(gx[:, 2 * H:] + r * gh[:, 2 * H:])
        h = (1 - z) * n + z * h
        ys[t] = h
    return gi, w_hh.to(wdt), b_hh, h0, ys, h


def _dev_whh(w, precision):
    """W_hh on the device in the layout tp_gru_recurrence expects for `precision`."""
    wf = w.float().to(DEV).contiguous()
    if precision == "fp32":
        return wf
    out = torch.empty(wf.shape, device=DEV, dtype=torch.bfloat16)
    nv.check(nv.lib().tp_pack_whh_bf16(nv.ptr(wf), nv.ptr(out), wf.shape[1], nv.stream()))
    return out


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("B,T,H", [(2, 4, 64), (32, 16, 256), (5, 3, 96), (1, 6, 1024), (40, 2, 128)])
def test_gru_recurrence(B, T, H, precision):
    L = nv.lib()
    cases = [_gru_case(B, T, H, precision, 1, False, False), _gru_case(B, T, H, precision, 2, True, True),
             _gru_case(B, 1, H, precision, 3, False, False)]
    keep, jobs, outs = [], [], []
    for i, (gi, w, b, h0, ys, hT) in enumerate(cases):
        Tj = gi.shape[0]
        rev = i == 1
        d = dict(gi=gi.to(DEV), w=_dev_whh(w, precision), b=cu(b), h0=None if h0 is None else cu(h0),
                 y=torch.zeros(Tj, B, H + 8, device=DEV), ylp=torch.zeros(Tj, B, H, device=DEV, dtype=torch.bfloat16),
                 hf=torch.zeros(B, 2 * H, device=DEV))
        keep.append(d)
        j = nv.GruJob()
        j.gi, j.ldg, j.w_hh, j.b_hh = d["gi"].data_ptr(), 3 * H, d["w"].data_ptr(), d["b"].data_ptr()
        j.h0 = 0 if d["h0"] is None else d["h0"].data_ptr()
        j.y, j.ldy, j.y_lp, j.ldy_lp = d["y"].data_ptr(), H + 8, d["ylp"].data_ptr(), H
        j.h_final, j.ld_hf = d["hf"].data_ptr() + 4 * H, 2 * H
        j.steps = Tj
        j.t_in0, j.t_in_step = (Tj - 1, -1) if rev else (0, 1)
        j.t_out0, j.t_out_step = (Tj - 1, -1) if rev else (0, 1)
        jobs.append(j)
        outs.append((ys, hT))
    arr = (nv.GruJob * 3)(*jobs)
    ws = nv.workspace(L.tp_gru_workspace_bytes(3, B, H), DEV)
    nv.check(L.tp_gru_recurrence(arr, 3, B, H, nv.PRECISIONS[precision], nv.ptr(ws), ws.numel(), nv.stream()))
    torch.cuda.synchronize()
    tol = 2e-5 if precision == "fp32" else 2e-4   # bf16: h is re-quantised each step in both; fp32 accumulation order differs
    for d, (ys, hT) in zip(keep, outs):
        y = d["y"][:, :, :H].cpu().double()
        assert float((y - ys).abs().max()) < tol
        assert float((d["hf"][:, H:].cpu().double() - hT).abs().max()) < tol
        assert float(d["hf"][:, :H].abs().max()) == 0.0          # neighbouring columns untouched
        assert float((d["ylp"].float().cpu().double() - ys).abs().max()) < 1e-2


def test_gru_recurrence_full_size_fp32_vs_torch_gru():
    """B=32, T=16, H=2048 (BASELINE config 2 shape) against torch.nn.GRU on CPU."""
    B, T, H, F = 32, 16, 2048, 64
    torch.manual_seed(0)
    gru = torch.nn.GRU(F, H)
    x = torch.randn(T, B, F)
    with torch.no_grad():
        y_ref, _ = gru(x)
        gi = x @ gru.weight_ih_l0.t() + gru.bias_ih_l0
    L = nv.lib()
    for precision, tol in (("fp32", 3e-5), ("bf16", 3e-3)):
        wdt = torch.bfloat16 if precision == "bf16" else torch.float32
        d = dict(gi=gi.to(DEV).contiguous(), w=_dev_whh(gru.weight_hh_l0.detach(), precision),
                 b=gru.bias_hh_l0.detach().to(DEV), y=torch.zeros(T, B, H, device=DEV))
        j = nv.GruJob()
        j.gi, j.ldg, j.w_hh, j.b_hh = d["gi"].data_ptr(), 3 * H, d["w"].data_ptr(), d["b"].data_ptr()
        j.y, j.ldy, j.steps, j.t_in0, j.t_in_step, j.t_out0, j.t_out_step = d["y"].data_ptr(), H, T, 0, 1, 0, 1
        arr = (nv.GruJob * 1)(j)
        ws = nv.workspace(L.tp_gru_workspace_bytes(1, B, H), DEV)
        nv.check(L.tp_gru_recurrence(arr, 1, B, H, nv.PRECISIONS[precision], nv.ptr(ws), ws.numel(), nv.stream()))
        torch.cuda.synchronize()
        err = float((d["y"].cpu() - y_ref).abs().max())
        assert err < tol, (precision, err)


# ------------------------------------------------------------------ SMPL
@pytest.mark.parametrize("n", [1, 5, 32, 100])
def test_smpl_forward_all_pose_kinds(n):
    with base_data_cwd(7):
        smpl = tepose_b200.SMPL(tepose_b200.SMPL_MODEL_DIR, batch_size=1, create_transl=False).to(DEV)
    m = torch_ref.SmplModel.synthetic(7)
    bodies = synth.make_bodies(7, n)
    aa, betas = torch.from_numpy(bodies["pose_aa"]), torch.from_numpy(bodies["betas"])
    v_ref, j_ref, R = torch_ref.smpl_forward(m, betas, pose_aa=aa)
    out = smpl(betas=cu(betas), body_pose=cu(aa[:, 3:]), global_orient=cu(aa[:, :3]), pose2rot=True)
    # north_star: max vertex / joint abs error <= 1e-4 m in fp32 (we hold 2e-5)
    assert float((out.vertices.cpu() - v_ref).abs().max()) < 2e-5
    assert float((out.joints.cpu() - j_ref).abs().max()) < 2e-5
    out = smpl(betas=cu(betas), body_pose=cu(R[:, 1:]), global_orient=cu(R[:, :1]), pose2rot=False)
    assert float((out.vertices.cpu() - v_ref).abs().max()) < 2e-5
    assert float((out.joints.cpu() - j_ref).abs().max()) < 2e-5


@pytest.mark.parametrize("n", [1, 33, 100])
def test_smpl_forward_tensor_core_blend(n):
    """bf16 tensor-core blend (K4): <= 1 mm by the north star; we hold 1e-4 m on SMPL-shaped data."""
    with base_data_cwd(9):
        smpl = tepose_b200.SMPL(tepose_b200.SMPL_MODEL_DIR, batch_size=1, create_transl=False).to(DEV)
    smpl.blend_precision = "bf16"
    m = torch_ref.SmplModel.synthetic(9)
    bodies = synth.make_bodies(9, n)
    aa, betas = torch.from_numpy(bodies["pose_aa"]), torch.from_numpy(bodies["betas"])
    v_ref, j_ref, R = torch_ref.smpl_forward(m, betas, pose_aa=aa)
    out = smpl(betas=cu(betas), body_pose=cu(aa[:, 3:]), global_orient=cu(aa[:, :3]), pose2rot=True)
    ev = float((out.vertices.cpu() - v_ref).abs().max())
    ej = float((out.joints.cpu() - j_ref).abs().max())
    assert ev < 1e-4 and ej < 1e-4, (ev, ej)


def test_smpl_size_independent_properties_large_batch():
    """4096 bodies (oracle too slow): identity pose => verts == v_shaped; a global rotation is
    rigid about the root joint."""
    n = 4096
    with base_data_cwd(8):
        smpl = tepose_b200.SMPL(tepose_b200.SMPL_MODEL_DIR, batch_size=1, create_transl=False).to(DEV)
    m = torch_ref.SmplModel.synthetic(8)
    betas = cu(synth.make_bodies(8, n)["betas"])
    zero = torch.zeros(n, 72, device=DEV)
    out0 = smpl(betas=betas, body_pose=zero[:, 3:], global_orient=zero[:, :3], pose2rot=True)
    v_shaped = m.v_template.to(DEV)[None] + torch.einsum("bl,mkl->bmk", betas, m.shapedirs.to(DEV))
    assert float((out0.vertices - v_shaped).abs().max()) < 1e-5
    rot = zero.clone()
    rot[:, :3] = torch.tensor([0.3, -0.5, 0.2], device=DEV)
    out1 = smpl(betas=betas, body_pose=rot[:, 3:], global_orient=rot[:, :3], pose2rot=True)
    Rg = torch_ref.batch_rodrigues_smplx(torch.tensor([[0.3, -0.5, 0.2]]))[0].to(DEV)
    root = torch.einsum("bik,i->bk", v_shaped, m.J_regressor[0].to(DEV))[:, None]
    assert float((out1.vertices - ((out0.vertices - root) @ Rg.t() + root)).abs().max()) < 1e-5


@pytest.mark.parametrize("n", [1024, 1500])
def test_smpl_large_batch_split_path(n):
    """>= 1024 bodies in bf16 blend mode: tcgen05 blend GEMM over L2-resident chunks + the skinning kernel.  Checked
    against the oracle on a sample of bodies and against the fused kernel (same bf16 operands) on all of them."""
    with base_data_cwd(10):
        smpl = tepose_b200.SMPL(tepose_b200.SMPL_MODEL_DIR, batch_size=1, create_transl=False).to(DEV)
    smpl.blend_precision = "bf16"
    m = torch_ref.SmplModel.synthetic(10)
    bodies = synth.make_bodies(10, n)
    aa, betas = torch.from_numpy(bodies["pose_aa"]), torch.from_numpy(bodies["betas"])
    out = smpl(betas=cu(betas), body_pose=cu(aa[:, 3:]), global_orient=cu(aa[:, :3]), pose2rot=True)
    pick = torch.tensor([0, 1, 15, 16, 511, 512, 513, 1023, n - 1])
    v_ref, j_ref, _ = torch_ref.smpl_forward(m, betas[pick], pose_aa=aa[pick])
    ev = float((out.vertices.cpu()[pick] - v_ref).abs().max())
    ej = float((out.joints.cpu()[pick] - j_ref).abs().max())
    assert ev < 1e-4 and ej < 1e-4, (ev, ej)
    # the fused kernel on sub-batches below the threshold: same operands, fp32 summation order differs only
    fv, fj = [], []
    for lo in range(0, n, 512):
        o = smpl(betas=cu(betas[lo:lo + 512]), body_pose=cu(aa[lo:lo + 512, 3:]), global_orient=cu(aa[lo:lo + 512, :3]), pose2rot=True)
        fv.append(o.vertices); fj.append(o.joints)
    assert float((out.vertices - torch.cat(fv)).abs().max()) < 2e-5
    assert float((out.joints - torch.cat(fj)).abs().max()) < 2e-5


def test_smpl_large_batch_h36m_regressor_folded():
    """Large-batch tcgen05 path with the 17-row H36M regressor (lib/models/spin.py:275-278 on a big batch): more rows than the
    in-kernel regressor handles, so the joints come from the folded-regressor GEMM (tp_smpl_regfold).  Against the oracle's
    J_regressor_h36m . verts and against the small-batch kernel."""
    n = 2048
    with base_data_cwd(10):
        smpl = tepose_b200.SMPL(tepose_b200.SMPL_MODEL_DIR, batch_size=1, create_transl=False).to(DEV)
    from tepose_b200.smpl import smpl_forward_native
    m = torch_ref.SmplModel.synthetic(10)
    p = smpl.packed()
    jreg, src = smpl.h36m_tables(m.J_regressor_h36m.to(DEV))
    bodies = synth.make_bodies(12, n)
    aa, betas = cu(torch.from_numpy(bodies["pose_aa"])), cu(torch.from_numpy(bodies["betas"]))
    big = smpl_forward_native(p, aa, 72, nv.POSE_AXIS_ANGLE, betas, 10, None, 0, n, jreg, src, want_theta=False, want_rotmat=False, blend_mode=1)
    pick = torch.tensor([0, 5, 1023, 1024, n - 1])
    v_ref, _, _ = torch_ref.smpl_forward(m, betas.cpu()[pick], pose_aa=aa.cpu()[pick])
    j_ref = torch.einsum("jv,bvc->bjc", m.J_regressor_h36m, v_ref)[:, tepose_b200.H36M_TO_J14]
    assert float((big[0].cpu()[pick] - v_ref).abs().max()) < 1e-4
    assert float((big[1].cpu()[pick] - j_ref).abs().max()) < 1e-4
    small = smpl_forward_native(p, aa[:512], 72, nv.POSE_AXIS_ANGLE, betas[:512], 10, None, 0, 512, jreg, src, want_theta=False,
                                want_rotmat=False, blend_mode=1)
    assert float((big[1][:512] - small[1]).abs().max()) < 3e-5


@pytest.mark.parametrize("um", ["0", "1"])
def test_smpl_large_batch_older_paths_still_agree(um):
    """TP_SMPL_UM=1 (tcgen05 blend + shared-memory skinning gathers) and =0 (GEMM + k_smpl_skin) stay selectable; the switch is
    read once per process, so the split-path test runs in a child process."""
    import subprocess, sys
    env = dict(os.environ, TP_SMPL_UM=um)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "--no-header", "-p", "no:cacheprovider",
                        "-k", "test_smpl_large_batch_split_path"], env=env, capture_output=True, text=True, timeout=600,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("skinning", ["random", "coherent"])
def test_smpl_config3_65536_bodies(skinning):
    """BASELINE.json configs[3] at its full size: 65,536 bodies through the fused tcgen05 blend + skinning kernel (k_smpl_lbs_um).
    256 sampled bodies against the oracle (<= 1e-4 m), and ALL bodies against the small-batch fused kernel (k_smpl_verts_tc, same
    bf16 blend operands, different summation order) in chunks of 512.  Both skinning layouts of the synthetic body model: 4 random
    joints per vertex (no locality) and SMPL-like runs of vertices that share their joints."""
    from tepose_b200 import synthetic as psynth
    n = 65536
    model, _ = psynth.build_synthetic_model(10, 16, 1, 64, "bf16", "cuda:0", skinning=skinning)
    smpl = model.regressor.smpl
    smpl.blend_precision = "bf16"
    bodies = synth.make_bodies(11, n)
    aa, betas = torch.from_numpy(bodies["pose_aa"]), torch.from_numpy(bodies["betas"])
    aa_d, betas_d = cu(aa), cu(betas)
    out = smpl(betas=betas_d, body_pose=aa_d[:, 3:], global_orient=aa_d[:, :3], pose2rot=True)
    assert tuple(out.vertices.shape) == (n, 6890, 3) and tuple(out.joints.shape) == (n, 49, 3)
    assert bool(torch.isfinite(out.vertices).all())
    g = torch.Generator().manual_seed(5)
    pick = torch.cat([torch.tensor([0, 1, 31, 32, 33, n - 33, n - 32, n - 1]), torch.randint(0, n, (248,), generator=g)])
    sm = psynth.make_smpl_model(10, skinning)
    m = torch_ref.SmplModel.synthetic(10)
    m.lbs_weights = torch.from_numpy(sm["weights"])
    v_ref, j_ref, _ = torch_ref.smpl_forward(m, betas[pick], pose_aa=aa[pick])
    ev = float((out.vertices[pick.to(DEV)].cpu() - v_ref).abs().max())
    ej = float((out.joints[pick.to(DEV)].cpu() - j_ref).abs().max())
    assert ev < 1e-4 and ej < 1e-4, (ev, ej)
    worst_v = worst_j = 0.0
    for lo in range(0, n, 512):
        o = smpl(betas=betas_d[lo:lo + 512], body_pose=aa_d[lo:lo + 512, 3:], global_orient=aa_d[lo:lo + 512, :3], pose2rot=True)
        worst_v = max(worst_v, float((out.vertices[lo:lo + 512] - o.vertices).abs().max()))
        worst_j = max(worst_j, float((out.joints[lo:lo + 512] - o.joints).abs().max()))
    assert worst_v < 3e-5 and worst_j < 3e-5, (worst_v, worst_j)


@pytest.mark.parametrize("B,T0,T1,H", [(32, 16, 16, 2048), (9, 5, 3, 128), (17, 2, 6, 1024), (32, 1, 4, 256), (1, 16, 16, 2048), (5, 3, 4, 256)])
def test_gru_recurrence_two_interleaved_directions(B, T0, T1, H):
    """bf16, exactly two matmul jobs without h0 at batch <= 32 -> k_gru_bf16_dual (independent 5-warp teams per direction,
    different step counts allowed) + one single-step job handled as plain gate math; also with a caller-provided barrier."""
    L = nv.lib()
    cases = [_gru_case(B, T0, H, "bf16", 11, False, False), _gru_case(B, T1, H, "bf16", 12, False, True),
             _gru_case(B, 1, H, "bf16", 13, False, False)]
    for use_slot in (False, True):
        keep, jobs, outs = [], [], []
        for i, (gi, w, b, h0, ys, hT) in enumerate(cases):
            Tj = gi.shape[0]
            rev = i == 1
            d = dict(gi=gi.to(DEV), w=_dev_whh(w, "bf16"), b=cu(b), y=torch.zeros(Tj, B, H, device=DEV),
                     ylp=torch.zeros(Tj, B, H, device=DEV, dtype=torch.bfloat16), hf=torch.zeros(B, H, device=DEV))
            keep.append(d)
            j = nv.GruJob()
            j.gi, j.ldg, j.w_hh, j.b_hh = d["gi"].data_ptr(), 3 * H, d["w"].data_ptr(), d["b"].data_ptr()
            j.h0 = 0
            j.y, j.ldy, j.y_lp, j.ldy_lp = d["y"].data_ptr(), H, d["ylp"].data_ptr(), H
            j.h_final, j.ld_hf = d["hf"].data_ptr(), H
            j.steps = Tj
            j.t_in0, j.t_in_step = (Tj - 1, -1) if rev else (0, 1)
            j.t_out0, j.t_out_step = (Tj - 1, -1) if rev else (0, 1)
            jobs.append(j)
            outs.append((ys, hT))
        arr = (nv.GruJob * 3)(*jobs)
        ws = nv.workspace(L.tp_gru_workspace_bytes(3, B, H), DEV)
        slot = torch.zeros(256, device=DEV, dtype=torch.int32)
        nv.check(L.tp_gru_recurrence_ex(arr, 3, B, H, nv.PRECISION_BF16, nv.ptr(ws), ws.numel(),
                                        nv.vp(slot.data_ptr() if use_slot else 0), nv.stream()))
        torch.cuda.synchronize()
        for d, (ys, hT) in zip(keep, outs):
            assert float((d["y"].cpu().double() - ys).abs().max()) < 2e-4
            assert float((d["hf"].cpu().double() - hT).abs().max()) < 2e-4
            assert float((d["ylp"].float().cpu().double() - ys).abs().max()) < 1e-2


def _dev_whh_umma(w):
    wf = w.float().to(DEV).contiguous()
    out = torch.empty(nv.lib().tp_whh_umma_bytes(wf.shape[1]), device=DEV, dtype=torch.uint8)
    nv.check(nv.lib().tp_pack_whh_umma(nv.ptr(wf), nv.ptr(out), wf.shape[1], nv.stream()))
    return out


@pytest.mark.parametrize("B,steps,H", [(32, (16, 16), 2048), (1, (16, 16), 2048), (9, (5, 3), 128), (17, (2, 6, 4), 1024),
                                       (32, (1, 4), 256), (5, (3, 4, 2, 5), 384), (32, (7,), 2048), (3, (6, 6), 640)])
def test_gru_recurrence_umma(B, steps, H):
    """bf16, every matmul job carries a tp_pack_whh_umma image -> k_gru_umma (tcgen05, W_hh resident in TMEM + shared
    memory, cluster pairs splitting K); 1..4 matmul jobs with different step counts + one single-step job as plain gate
    math (only when a job slot is left); ragged batch rows, strided outputs, caller-provided barrier slot.  Same tolerance
    as the mma.sync kernels against the float64 cell with the state re-quantised to bf16 per step."""
    L = nv.lib()
    cases = [_gru_case(B, t, H, "bf16", 21 + i, False, i % 2 == 1) for i, t in enumerate(steps)]
    if len(cases) < 4:
        cases.append(_gru_case(B, 1, H, "bf16", 29, False, False))
    n = len(cases)
    for use_slot in (False, True):
        keep, jobs, outs = [], [], []
        for i, (gi, w, b, h0, ys, hT) in enumerate(cases):
            Tj = gi.shape[0]
            rev = i % 2 == 1 and i < len(steps)
            d = dict(gi=gi.to(DEV), w=_dev_whh(w, "bf16"), wu=_dev_whh_umma(w), b=cu(b), y=torch.zeros(Tj, B, H + 8, device=DEV),
                     ylp=torch.zeros(Tj, B, H, device=DEV, dtype=torch.bfloat16), hf=torch.zeros(B, 2 * H, device=DEV))
            keep.append(d)
            j = nv.GruJob()
            j.gi, j.ldg, j.w_hh, j.b_hh = d["gi"].data_ptr(), 3 * H, d["w"].data_ptr(), d["b"].data_ptr()
            j.w_hh_umma = d["wu"].data_ptr()
            j.h0 = 0
            j.y, j.ldy, j.y_lp, j.ldy_lp = d["y"].data_ptr(), H + 8, d["ylp"].data_ptr(), H
            j.h_final, j.ld_hf = d["hf"].data_ptr() + 4 * H, 2 * H
            j.steps = Tj
            j.t_in0, j.t_in_step = (Tj - 1, -1) if rev else (0, 1)
            j.t_out0, j.t_out_step = (Tj - 1, -1) if rev else (0, 1)
            jobs.append(j)
            outs.append((ys, hT))
        arr = (nv.GruJob * n)(*jobs)
        ws = nv.workspace(L.tp_gru_workspace_bytes(n, B, H), DEV)
        slot = torch.zeros(256, device=DEV, dtype=torch.int32)
        nv.check(L.tp_gru_recurrence_ex(arr, n, B, H, nv.PRECISION_BF16, nv.ptr(ws), ws.numel(),
                                        nv.vp(slot.data_ptr() if use_slot else 0), nv.stream()))
        torch.cuda.synchronize()
        for i, (d, (ys, hT)) in enumerate(zip(keep, outs)):
            y = d["y"][:, :, :H].cpu().double()
            assert float((y - ys).abs().max()) < 2e-4, (i, float((y - ys).abs().max()))
            assert float(d["y"][:, :, H:].abs().max()) == 0.0
            assert float((d["hf"][:, H:].cpu().double() - hT).abs().max()) < 2e-4
            assert float(d["hf"][:, :H].abs().max()) == 0.0
            assert float((d["ylp"].float().cpu().double() - ys).abs().max()) < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("N,n_iter,per_row_init", [(32, 3, False), (1, 3, False), (5, 1, True), (32, 4, True), (17, 0, False), (8, 3, True)])
def test_ief_cluster_kernel(N, n_iter, per_row_init):
    """k_ief_cluster (the IEF iterations of lib/models/spin.py:250-261 on one 16-CTA cluster, weights on chip) against
    (a) the same iterations inside the grid-barrier kernel (same bf16 operands, another summation order) and
    (b) a float64 evaluation of the loop whose inter-layer activations are rounded to bf16 like the kernels' are.  The decoder
    weights are made large so that the state really moves between iterations."""
    from tepose_b200 import _native as nv
    from tepose_b200.spin import PSC
    from tepose_b200.synthetic import build_synthetic_model
    model, _ = build_synthetic_model(77, 4, 1, 128, "bf16", DEV)
    reg = model.regressor
    g = torch.Generator().manual_seed(N * 10 + n_iter)
    with torch.no_grad():
        for lin, s in ((reg.decpose, 0.02), (reg.decshape, 0.02), (reg.deccam, 0.02)):
            lin.weight.copy_((torch.randn(lin.weight.shape, generator=g) * s).to(DEV))
            lin.bias.copy_((torch.randn(lin.bias.shape, generator=g) * 0.1).to(DEV))
        reg.fc2.bias.copy_((torch.randn(1024, generator=g) * 0.1).to(DEV))
    pk = reg.packed()
    feat = (torch.randn(N, 2048, generator=g)).to(DEV)
    init = pk["init"]
    if per_row_init:
        init = (pk["init"].cpu() + 0.2 * torch.randn(N, PSC, generator=g)).to(DEV).contiguous()
        init[:, 157:] = 0
    L = nv.lib()

    def run(cluster):
        was = L.tp_set_ief_cluster(1 if cluster else 0)
        try:
            psc = torch.full((N, PSC), float("nan"), device=DEV)
            ws = nv.workspace(L.tp_ief_workspace_bytes(N), torch.device(DEV))
            nv.check(L.tp_ief_forward(nv.PRECISIONS["bf16"], pk["c"], nv.ptr(feat), None, N, nv.ptr(init), init.shape[0], n_iter, nv.ptr(psc),
                                      nv.ptr(ws), ws.numel(), nv.stream()), "tp_ief_forward")
            torch.cuda.synchronize()
            return psc.cpu()
        finally:
            L.tp_set_ief_cluster(was)

    got, grid = run(True), run(False)
    bf = lambda t: t.to(torch.bfloat16).double()
    W1 = bf(reg.fc1.weight.detach().cpu()); W2 = bf(reg.fc2.weight.detach().cpu())
    Wd = bf(torch.cat([reg.decpose.weight, reg.decshape.weight, reg.deccam.weight]).detach().cpu())
    bd = torch.cat([reg.decpose.bias, reg.decshape.bias, reg.deccam.bias]).detach().cpu().double()
    st = init.cpu().double().expand(N, PSC)[:, :157].clone()
    base = bf(feat.cpu()) @ W1[:, :2048].T + reg.fc1.bias.detach().cpu().double()
    for _ in range(n_iter):
        u1 = bf((base + bf(st.float()) @ W1[:, 2048:].T).float())
        u2 = bf((u1 @ W2.T + reg.fc2.bias.detach().cpu().double()).float())
        st = st + u2 @ Wd.T + bd
    assert torch.isfinite(got).all()
    err_ref = float((got[:, :157].double() - st).abs().max())
    err_grid = float((got[:, :157] - grid[:, :157]).abs().max())
    moved = float((st - init.cpu().double().expand(N, PSC)[:, :157]).abs().max())
    if n_iter:
        assert moved > 0.05, moved                      # the test would be vacuous otherwise
    # a bf16 rounding of u1 / u2 / the state may flip on a last-bit difference of the fp32 sums: 2^-9 relative on values ~1, times |Wd| row sums
    assert err_ref < 5e-3, (err_ref, err_grid)
    assert err_grid < 5e-3, (err_ref, err_grid)
    if n_iter == 0:
        assert torch.equal(got[:, :157], init.cpu().expand(N, PSC)[:, :157])

"""CPU emulation of the C-ABI entry points -- TEST INFRASTRUCTURE for the host-side logic.

`install()` swaps tepose_b200._native.lib() for an object that interprets the same calls
(same argument order, pointers, strides, structs) on CPU tensors with plain torch math.  It
lets the `-m "not gpu"` suite check the Python layer's packing, segment tables and
time-index bookkeeping against the oracle without a GPU.  It is never importable from the
product package and never used for a parity claim about the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C

import torch

import tepose_b200._native as nv
from oracle import torch_ref


def _view(ptr, count, dtype):
    addr = ptr.value if isinstance(ptr, C.c_void_p) else int(ptr or 0)
    if not addr:
        return None
    nbytes = count * torch.empty((), dtype=dtype).element_size()
    buf = (C.c_char * nbytes).from_address(addr)
    return torch.frombuffer(buf, dtype=dtype)


def _mat(ptr, rows, cols, ld, dtype=torch.float32):
    """rows x cols view with row stride ld (elements)."""
    if rows == 0 or cols == 0:
        return torch.empty(rows, cols, dtype=dtype)
    flat = _view(ptr, (rows - 1) * ld + cols, dtype)
    return torch.as_strided(flat, (rows, cols), (ld, 1))


class FakeLib:
    def __init__(self):
        self.calls = []

    # ---- misc
    def tp_version(self):
        return 100

    def tp_last_error(self):
        return b"fake"

    # ---- geometry
    def tp_rot6d_to_rotmat(self, x, R, n, stream):
        _view(R, n * 9, torch.float32).copy_(torch_ref.rot6d_to_rotmat(_view(x, n * 6, torch.float32).clone().reshape(n, 6)).reshape(-1))
        return 0

    def tp_rotmat_to_angle_axis(self, R, aa, n, stream):
        _view(aa, n * 3, torch.float32).copy_(torch_ref.rotmat_to_angle_axis(_view(R, n * 9, torch.float32).clone().reshape(n, 3, 3)).reshape(-1))
        return 0

    def tp_batch_rodrigues(self, aa, R, n, form, stream):
        a = _view(aa, n * 3, torch.float32).clone().reshape(n, 3)
        out = torch_ref.batch_rodrigues_quat(a) if form == nv.RODRIGUES_QUAT else torch_ref.batch_rodrigues_smplx(a)
        _view(R, n * 9, torch.float32).copy_(out.reshape(-1))
        return 0

    def tp_projection(self, joints, cam, kp2d, n, nj, stream):
        j = _view(joints, n * nj * 3, torch.float32).clone().reshape(n, nj, 3)
        c = _view(cam, n * 3, torch.float32).clone().reshape(n, 3)
        _view(kp2d, n * nj * 2, torch.float32).copy_(torch_ref.projection(j, c).reshape(-1))
        return 0

    # ---- packing / GEMMs
    def tp_pack_rows(self, src, stride_b, stride_t, rows_b, rows_t, k, dst, kp, prec, relu, stream):
        flat = _view(src, (rows_b - 1) * stride_b + (rows_t - 1) * stride_t + k, torch.float32)
        s = torch.as_strided(flat, (rows_t, rows_b, k), (stride_t, stride_b, 1))
        dt = torch.bfloat16 if prec == nv.PRECISION_BF16 else torch.float32
        d = _mat(dst, rows_t * rows_b, kp, kp, dt)
        d.zero_()
        v = s.reshape(rows_t * rows_b, k)
        d[:, :k] = (v.clamp_min(0) if relu else v).to(dt)
        self.calls.append(("pack_rows", rows_b, rows_t, k, kp, prec))
        return 0

    def tp_pack_rows_ex(self, src, stride_b, stride_t, rows_b, rows_t, k, dst, kp, prec, relu, zero, zero_bytes, stream):
        z = _view(zero, zero_bytes, torch.uint8)
        if z is not None:
            z.zero_()
        return self.tp_pack_rows(src, stride_b, stride_t, rows_b, rows_t, k, dst, kp, prec, relu, stream)

    def tp_pack_rows_f16(self, src, stride_b, stride_t, rows_b, rows_t, k, dst, kp, prec, relu, zero, zero_bytes, stream):
        z = _view(zero, zero_bytes, torch.uint8)
        if z is not None:
            z.zero_()
        flat = _view(src, (rows_b - 1) * stride_b + (rows_t - 1) * stride_t + k, torch.float16)
        s = torch.as_strided(flat, (rows_t, rows_b, k), (stride_t, stride_b, 1)).float()
        dt = torch.bfloat16 if prec == nv.PRECISION_BF16 else torch.float32
        d = _mat(dst, rows_t * rows_b, kp, kp, dt)
        d.zero_()
        v = s.reshape(rows_t * rows_b, k)
        d[:, :k] = (v.clamp_min(0) if relu else v).to(dt)
        return 0

    def tp_split3_bf16(self, src, ld_src, rows, k, dst, stream):
        v = _mat(src, rows, k, ld_src).clone()
        hi = v.to(torch.bfloat16)
        lo = (v - hi.float()).to(torch.bfloat16)
        _mat(dst, rows, 3 * k, 3 * k, torch.bfloat16).copy_(torch.cat([hi, lo, hi], dim=1))
        return 0

    def tp_unpack_rows_residual(self, y, ld_y, x, stride_b, stride_t, rows_b, rows_t, k, out, out_bf16, stream):
        v = _mat(y, rows_t * rows_b, k, ld_y).clone().reshape(rows_t, rows_b, k).permute(1, 0, 2)
        if (x.value if isinstance(x, C.c_void_p) else x):
            flat = _view(x, (rows_b - 1) * stride_b + (rows_t - 1) * stride_t + k, torch.float32)
            v = v + torch.as_strided(flat, (rows_b, rows_t, k), (stride_b, stride_t, 1))
        v = v.reshape(rows_b * rows_t, k)
        _mat(out, rows_b * rows_t, k, k).copy_(v)
        if (out_bf16.value if isinstance(out_bf16, C.c_void_p) else out_bf16):
            _mat(out_bf16, rows_b * rows_t, k, k, torch.bfloat16).copy_(v.to(torch.bfloat16))
        self.calls.append(("unpack_rows_residual", rows_b, rows_t, k))
        return 0

    def tp_gemm_f32(self, A, lda, W, ldw, bias, Cin, ldcin, Cout, ldc, M, N, K, alpha, beta, relu_a, stream):
        a = _mat(A, M, K, lda).clone()
        w = _mat(W, N, K, ldw)
        if relu_a:
            a = a.clamp_min(0)
        v = a @ w.t()
        b = _view(bias, N, torch.float32)
        if b is not None:
            v = v + b
        v = v * alpha
        if (Cin.value if isinstance(Cin, C.c_void_p) else Cin):
            v = v + beta * _mat(Cin, M, N, ldcin).clone()
        _mat(Cout, M, N, ldc).copy_(v)
        self.calls.append(("gemm_f32", M, N, K))
        return 0

    def tp_gemm_f32_splitk_workspace_bytes(self, M, N, splits):
        return 4096 + 4 * M * N * splits

    def tp_gemm_f32_splitk(self, A, lda, W, ldw, bias, Cin, ldcin, Cout, ldc, M, N, K, alpha, beta, relu_a, splits, ws, ws_bytes, stream):
        return self.tp_gemm_f32(A, lda, W, ldw, bias, Cin, ldcin, Cout, ldc, M, N, K, alpha, beta, relu_a, stream)

    def tp_gemm_bf16_tc(self, A, a_rows, W, w_rows, kp, segs, nseg, stream):
        a = _mat(A, a_rows, kp, kp, torch.bfloat16).float()
        w = _mat(W, w_rows, kp, kp, torch.bfloat16).float()
        for i in range(nseg):
            sg = segs[i]
            v = a[sg.m_start:sg.m_start + sg.m_rows] @ w[sg.n_start:sg.n_start + sg.n_cols].t()
            if sg.bias:
                v = v + _view(sg.bias, sg.n_cols, torch.float32)
            if sg.residual:
                v = v + _mat(sg.residual, sg.m_rows, sg.n_cols, sg.ldr, torch.bfloat16).float()
            if sg.flags & nv.GEMM_RELU:
                v = v.clamp_min(0)
            if sg.flags & nv.GEMM_OUT_BF16:
                _mat(sg.out, sg.m_rows, sg.n_cols, sg.ldc, torch.bfloat16).copy_(v.to(torch.bfloat16))
            else:
                _mat(sg.out, sg.m_rows, sg.n_cols, sg.ldc).copy_(v)
        self.calls.append(("gemm_bf16_tc", a_rows, w_rows, kp, nseg))
        return 0

    # ---- HMR data movement (NHWC bf16)
    def tp_nchw_to_nhwc_bf16(self, x, y, N, Cc, H, W, CP, stream):
        xin = _view(x, N * Cc * H * W, torch.float32).reshape(N, Cc, H, W)
        out = torch.zeros(N, H, W, CP, dtype=torch.bfloat16)
        out[..., :Cc] = xin.permute(0, 2, 3, 1).to(torch.bfloat16)
        _view(y, N * H * W * CP, torch.bfloat16).copy_(out.reshape(-1))
        return 0

    def tp_im2col_nhwc_bf16(self, inp, out, N, H, W, Cc, kh, kw, stride, pad, KP, stream):
        a = _view(inp, N * H * W * Cc, torch.bfloat16).reshape(N, H, W, Cc).float().permute(0, 3, 1, 2)
        Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        cols = torch.nn.functional.unfold(a, (kh, kw), padding=pad, stride=stride)              # [N, C*kh*kw, L], (c, ky, kx) order
        cols = cols.view(N, Cc, kh * kw, Ho * Wo).permute(0, 3, 2, 1).reshape(N * Ho * Wo, kh * kw * Cc)
        o = torch.zeros(N * Ho * Wo, KP, dtype=torch.bfloat16)
        o[:, :kh * kw * Cc] = cols.to(torch.bfloat16)
        _view(out, N * Ho * Wo * KP, torch.bfloat16).copy_(o.reshape(-1))
        self.calls.append(("im2col", N, H, W, Cc, kh, stride))
        return 0

    def tp_maxpool3x3s2_nhwc_bf16(self, inp, out, N, H, W, Cc, stream):
        a = _view(inp, N * H * W * Cc, torch.bfloat16).reshape(N, H, W, Cc).float().permute(0, 3, 1, 2)
        o = torch.nn.functional.max_pool2d(a, 3, 2, 1).permute(0, 2, 3, 1).contiguous()
        _view(out, o.numel(), torch.bfloat16).copy_(o.to(torch.bfloat16).reshape(-1))
        return 0

    def tp_avgpool_nhwc_bf16(self, inp, out, N, HW, Cc, stream):
        a = _view(inp, N * HW * Cc, torch.bfloat16).reshape(N, HW, Cc).float()
        _view(out, N * Cc, torch.float32).copy_(a.mean(1).reshape(-1))
        return 0

    # ---- recurrence
    def tp_skinny_bf16_workspace_bytes(self, M, N, splits):
        return 4096

    def tp_whh_umma_bytes(self, H):
        return 0        # the fake has no tcgen05 kernel: the host layer then leaves w_hh_umma NULL

    def tp_pack_whh_bf16(self, w_hh, dst, H, stream):
        # the emulation keeps W_hh row-major bf16 (the fragment order only matters to the CUDA kernel)
        _mat(dst, 3 * H, H, H, torch.bfloat16).copy_(_mat(w_hh, 3 * H, H, H).to(torch.bfloat16))
        return 0

    def tp_gru_set_trace(self, buf):
        return None

    def tp_gru_workspace_bytes(self, njobs, B, H):
        return 256

    def tp_gru_recurrence(self, jobs, njobs, B, H, precision, ws, ws_bytes, stream):
        wdt = torch.bfloat16 if precision == nv.PRECISION_BF16 else torch.float32
        for i in range(njobs):
            jb = jobs[i]
            W = _mat(jb.w_hh, 3 * H, H, H, wdt).float()
            bh = _view(jb.b_hh, 3 * H, torch.float32)
            h = _mat(jb.h0, B, H, H).clone() if jb.h0 else torch.zeros(B, H)
            for s in range(jb.steps):
                t_in = jb.t_in0 + s * jb.t_in_step
                gi = _mat(jb.gi + 4 * t_in * B * jb.ldg, B, 3 * H, jb.ldg)
                hm = h.to(wdt).float() if precision == nv.PRECISION_BF16 else h
                gh = hm @ W.t() + bh
                r = torch.sigmoid(gi[:, :H] + gh[:, :H])
                z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
                n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
                h = (1 - z) * n + z * h
                t_out = jb.t_out0 + s * jb.t_out_step
                if jb.y:
                    _mat(jb.y + 4 * t_out * B * jb.ldy, B, H, jb.ldy).copy_(h)
                if jb.y_lp:
                    _mat(jb.y_lp + 2 * t_out * B * jb.ldy_lp, B, H, jb.ldy_lp, torch.bfloat16).copy_(h.to(torch.bfloat16))
            if jb.h_final:
                _mat(jb.h_final, B, H, jb.ld_hf).copy_(h)
        self.calls.append(("gru", njobs, B, H, precision, [jobs[i].steps for i in range(njobs)]))
        return 0

    # ---- regressor
    def tp_encoder_heads_workspace_bytes(self, B):
        return 256

    def tp_pack_mma_a_bytes(self, rows, cols):
        return ((rows + 15) // 16) * ((cols + 31) // 32) * 1024

    def tp_pack_mma_a_bf16(self, w, ld, rows, cols, dst, stream):
        # the emulation keeps the matrix row-major bf16 at the head of the (larger) packed buffer
        _mat(dst, rows, cols, cols, torch.bfloat16).copy_(_mat(w, rows, cols, ld).to(torch.bfloat16))
        return 0

    @staticmethod
    def _lin(precision, w, n, k):
        return _mat(w, n, k, k, torch.bfloat16).float() if precision == nv.PRECISION_BF16 else _mat(w, n, k, k)

    @staticmethod
    def _act(precision, a):
        return a.to(torch.bfloat16).float() if precision == nv.PRECISION_BF16 else a

    def tp_encoder_heads(self, precision, w_fwd, b_fwd, w_rec, b_rec, h_fwd, ld_hf, h_rec, ld_hr, B, H, is_train, feat, feat_lp, ws, ws_bytes, stream):
        q = lambda t: self._act(precision, t)
        a = q(_mat(h_fwd, B, H, ld_hf).clamp_min(0)) @ self._lin(precision, w_fwd, 2048, H).t() + _view(b_fwd, 2048, torch.float32)
        b = q(_mat(h_rec, B, 2 * H, ld_hr).clamp_min(0)) @ self._lin(precision, w_rec, 2048, 2 * H).t() + _view(b_rec, 2048, torch.float32)
        if is_train:
            _view(feat, B * 4096, torch.float32).copy_(torch.stack([a, b], 1).reshape(-1))
        else:
            _view(feat, B * 2048, torch.float32).copy_(((a + b) / 2).reshape(-1))
        return 0

    def tp_encoder_heads_cat(self, precision, w_cat, b_cat, h_cat, ld_h, B, H, feat, feat_lp, ws, ws_bytes, stream):
        a = self._act(precision, _mat(h_cat, B, 3 * H, ld_h).clamp_min(0))
        v = a @ self._lin(precision, w_cat, 2048, 3 * H).t() + _view(b_cat, 2048, torch.float32)
        _view(feat, B * 2048, torch.float32).copy_(v.reshape(-1))
        self.calls.append(("heads_cat", B, H))
        return 0

    def tp_gru_recurrence_ex(self, jobs, njobs, B, H, precision, ws, ws_bytes, barrier, stream):
        return self.tp_gru_recurrence(jobs, njobs, B, H, precision, ws, ws_bytes, stream)

    def tp_heads_ief_forward(self, w_cat, b_cat, h_cat, ld_h, H, w, n_rows, init, init_rows, n_iter, psc, ws, ws_bytes, barrier, stream):
        P = nv.PRECISION_BF16
        feat = torch.empty(n_rows, 2048)
        self.tp_encoder_heads_cat(P, w_cat, b_cat, h_cat, ld_h, n_rows, H, C.c_void_p(feat.data_ptr()), None, ws, ws_bytes, stream)
        return self.tp_ief_forward(P, w, C.c_void_p(feat.data_ptr()), None, n_rows, init, init_rows, n_iter, psc, ws, ws_bytes, stream)

    def tp_ief_workspace_bytes(self, n):
        return 256

    def tp_ief_forward(self, precision, w, feat, feat_lp, n_rows, init, init_rows, n_iter, psc, ws, ws_bytes, stream):
        w = w.contents if hasattr(w, "contents") else w
        q = lambda t: self._act(precision, t)
        lin = lambda ptr_, n, k: self._lin(precision, ptr_, n, k)
        x = _mat(feat, n_rows, 2048, 2048)
        p = _mat(init, init_rows, 160, 160).clone().expand(n_rows, -1).clone()
        base = q(x) @ lin(w.w1x, 1024, 2048).t() + _view(w.b1, 1024, torch.float32)
        for _ in range(n_iter):
            u1 = q(p) @ lin(w.w1p, 1024, 160).t() + base
            u2 = q(u1) @ lin(w.w2, 1024, 1024).t() + _view(w.b2, 1024, torch.float32)
            p = p + q(u2) @ lin(w.wdec, 160, 1024).t() + _view(w.bdec, 160, torch.float32)
        _mat(psc, n_rows, 160, 160).copy_(p)
        return 0

    # ---- SMPL
    def tp_smpl_workspace_bytes(self, m, n, nreg, blend_mode=0):
        return 256

    def tp_smpl_forward_ex(self, m, n, pose, ld_pose, pose_kind, betas, ld_betas, cam, ld_cam, jreg, nreg, fold, joint_src, nj, *rest):
        return self.tp_smpl_forward(m, n, pose, ld_pose, pose_kind, betas, ld_betas, cam, ld_cam, jreg, nreg, joint_src, nj, *rest)

    def tp_smpl_forward(self, m, n, pose, ld_pose, pose_kind, betas, ld_betas, cam, ld_cam, jreg, nreg, joint_src, nj,
                        verts, joints, kp2d, rotmat, theta, blend_mode, ws, ws_bytes, stream):
        m = m.contents if hasattr(m, "contents") else m
        V, vp = m.n_verts, m.vp
        blend = _view(m.blend, 218 * 3 * vp, torch.float32).reshape(218, 3, vp)[:, :, :V]
        jt = _view(m.j_template, 72, torch.float32).reshape(24, 3)
        jsd = _view(m.j_shapedirs, 720, torch.float32).reshape(24, 3, 10)
        parents = _view(m.parents, 24, torch.int32).long()
        sidx = _view(m.skin_idx, vp * m.ks, torch.int32).reshape(vp, m.ks)[:V].long()
        sw = _view(m.skin_w, vp * m.ks, torch.float32).reshape(vp, m.ks)[:V]
        width = {nv.POSE_ROTMAT: 216, nv.POSE_AXIS_ANGLE: 72, nv.POSE_ROT6D: 144}[pose_kind]
        P = _mat(pose, n, width, ld_pose).clone()
        beta = _mat(betas, n, 10, ld_betas).clone()
        if pose_kind == nv.POSE_ROTMAT:
            R = P.reshape(n, 24, 3, 3)
        elif pose_kind == nv.POSE_AXIS_ANGLE:
            R = torch_ref.batch_rodrigues_smplx(P.reshape(-1, 3)).reshape(n, 24, 3, 3)
        else:
            R = torch_ref.rot6d_to_rotmat(P).reshape(n, 24, 3, 3)
        coef = torch.cat([(R[:, 1:] - torch.eye(3)).reshape(n, 207), beta, torch.ones(n, 1)], 1)
        v_posed = torch.einsum("nk,kcv->nvc", coef, blend)
        J = jt[None] + torch.einsum("jcl,nl->njc", jsd, beta)
        WR, Wt = [None] * 24, [None] * 24
        for i in range(24):
            if parents[i] < 0:
                WR[i], Wt[i] = R[:, i], J[:, i]
            else:
                pa = int(parents[i])
                WR[i] = WR[pa] @ R[:, i]
                Wt[i] = (WR[pa] @ (J[:, i] - J[:, pa])[..., None])[..., 0] + Wt[pa]
        A = torch.stack([torch.cat([WR[i], (Wt[i] - (WR[i] @ J[:, i][..., None])[..., 0])[..., None]], -1) for i in range(24)], 1)
        Tm = (sw[None, :, :, None, None] * A[:, sidx]).sum(2)                     # [n,V,3,4]
        vout = (Tm[..., :3] @ v_posed[..., None])[..., 0] + Tm[..., 3]
        posedJ = torch.stack(Wt, 1)
        if verts:
            _view(verts, n * V * 3, torch.float32).copy_(vout.reshape(-1))
        if nreg:
            Jr = torch.einsum("rv,nvc->nrc", _view(jreg, nreg * vp, torch.float32).reshape(nreg, vp)[:, :V], vout)
        codes = _view(joint_src, nj, torch.int32).tolist() if nj else []
        outj = torch.zeros(n, nj, 3)
        for o, c in enumerate(codes):
            outj[:, o] = vout[:, c - 1000] if c >= 1000 else (Jr[:, c - 100] if c >= 100 else posedJ[:, c])
        if joints:
            _view(joints, n * nj * 3, torch.float32).copy_(outj.reshape(-1))
        cm = _mat(cam, n, 3, ld_cam).clone() if cam else None
        if kp2d and cm is not None:
            _view(kp2d, n * nj * 2, torch.float32).copy_(torch_ref.projection(outj, cm).reshape(-1))
        if rotmat:
            _view(rotmat, n * 216, torch.float32).copy_(R.reshape(-1))
        if theta:
            aa = torch_ref.rotmat_to_angle_axis(R.reshape(-1, 3, 3)).reshape(n, 72)
            c3 = cm if cm is not None else torch.zeros(n, 3)
            _view(theta, n * 85, torch.float32).copy_(torch.cat([c3, aa, beta], 1).reshape(-1))
        self.calls.append(("smpl", n, pose_kind, nreg, nj))
        return 0


    # ---- evaluation metrics
    def tp_pose_metrics(self, pred, target, n, J, p0, p1, aligned, mpjpe, pa, stream):
        P = _view(pred, n * J * 3, torch.float32).clone().reshape(n, J, 3)
        G = _view(target, n * J * 3, torch.float32).clone().reshape(n, J, 3)
        pelvis = None if p0 < 0 else (p0 if p1 < 0 else (p0, p1))
        m = torch_ref.pose_metrics(P.double(), G.double(), pelvis)
        for ptr_, key, cnt in ((aligned, "aligned", n * J * 3), (mpjpe, "mpjpe", n), (pa, "mpjpe_pa", n)):
            v = _view(ptr_, cnt, torch.float32)
            if v is not None:
                v.copy_(m[key].float().reshape(-1))
        self.calls.append(("pose_metrics", n, J, p0, p1))
        return 0

    def tp_accel_error(self, pred, target, n_seq, ln, J, p0, p1, out, stream):
        if n_seq == 0 or ln < 3:
            return 0
        pelvis = None if p0 < 0 else (p0 if p1 < 0 else (p0, p1))
        P = torch_ref.align_pelvis(_view(pred, n_seq * ln * J * 3, torch.float32).clone().reshape(n_seq, ln, J, 3).double(), pelvis)
        G = _view(target, n_seq * ln * J * 3, torch.float32)
        if G is not None:
            G = torch_ref.align_pelvis(G.clone().reshape(n_seq, ln, J, 3).double(), pelvis)
        _view(out, n_seq * (ln - 2), torch.float32).copy_(torch_ref.accel_error(P, G).float().reshape(-1))
        self.calls.append(("accel_error", n_seq, ln, J))
        return 0

    def tp_vertex_error(self, a, b, n, V, out, stream):
        A = _view(a, n * V * 3, torch.float32).clone().reshape(n, V, 3)
        B = _view(b, n * V * 3, torch.float32).clone().reshape(n, V, 3)
        _view(out, n, torch.float32).copy_(torch_ref.vertex_error(A, B))
        self.calls.append(("vertex_error", n, V))
        return 0


class _Patch:
    def __init__(self):
        self.fake = FakeLib()

    def __enter__(self):
        self._saved = (nv.lib, nv.require_cuda, nv.stream)
        nv.lib = lambda: self.fake
        nv.require_cuda = lambda t, name="tensor": None
        nv.stream = lambda: C.c_void_p(0)
        return self.fake

    def __exit__(self, *exc):
        nv.lib, nv.require_cuda, nv.stream = self._saved


def install():
    """Context manager: route tepose_b200's native calls to the CPU emulation."""
    return _Patch()

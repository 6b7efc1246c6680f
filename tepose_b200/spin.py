"""IEF Regressor with the reference's API (lib/models/spin.py:209-291) on the sm_100a kernels.

forward(x, init_pose=None, init_shape=None, init_cam=None, n_iter=3, is_train=False,
        J_regressor=None) -> [ {theta, verts, kp_2d, kp_3d, rotmat} ]
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _native as nv
from .geometry import projection  # noqa: F401  (re-exported: lib/models/spin.py:307)
from .smpl import SMPL, SMPL_MEAN_PARAMS, SMPL_MODEL_DIR, smpl_forward_native

NPOSE = 24 * 6
PSC = 160  # pose6d(144) | shape(10) | cam(3) | pad(3)


def folded_gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, relu_a: bool) -> torch.Tensor:
    """out [M,160] = act(a)[M,K] . w[160,K]^T + bias on the fp32 split-K kernel (M small, K long: every SM streams
    a K slice of the folded matrix)."""
    L = nv.lib()
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    if M > 64:
        nv.check(L.tp_gemm_f32(nv.vp(a.data_ptr()), a.stride(0), nv.ptr(w), K, nv.ptr(bias), nv.vp(0), 0, nv.ptr(out), N,
                               M, N, K, 1.0, 0.0, 1 if relu_a else 0, nv.stream()), "tp_gemm_f32")
        return out
    splits = max(1, min(K // 128, 24))
    ws = nv.workspace(max(int(L.tp_gemm_f32_splitk_workspace_bytes(M, N, splits)), 4096), a.device)
    ws[:4096].zero_()                                   # ticket area (the kernel leaves it zeroed)
    nv.check(L.tp_gemm_f32_splitk(nv.vp(a.data_ptr()), a.stride(0), nv.ptr(w), K, nv.ptr(bias), nv.vp(0), 0, nv.ptr(out), N,
                                  M, N, K, 1.0, 0.0, 1 if relu_a else 0, splits, nv.ptr(ws), ws.numel(), nv.stream()),
             "tp_gemm_f32_splitk")
    return out


class Regressor(nn.Module):
    def __init__(self, smpl_mean_params=SMPL_MEAN_PARAMS, precision="fp32"):
        super().__init__()
        self.precision = precision
        # parameter containers: same names / shapes / default init as lib/models/spin.py:215-224
        self.fc1 = nn.Linear(512 * 4 + NPOSE + 13, 1024)
        self.drop1 = nn.Dropout()
        self.fc2 = nn.Linear(1024, 1024)
        self.drop2 = nn.Dropout()
        self.decpose = nn.Linear(1024, NPOSE)
        self.decshape = nn.Linear(1024, 10)
        self.deccam = nn.Linear(1024, 3)
        nn.init.xavier_uniform_(self.decpose.weight, gain=0.01)
        nn.init.xavier_uniform_(self.decshape.weight, gain=0.01)
        nn.init.xavier_uniform_(self.deccam.weight, gain=0.01)
        self.smpl = SMPL(SMPL_MODEL_DIR, batch_size=64, create_transl=False)
        mean_params = np.load(smpl_mean_params)
        self.register_buffer('init_pose', torch.from_numpy(mean_params['pose'][:]).unsqueeze(0))
        self.register_buffer('init_shape', torch.from_numpy(mean_params['shape'][:].astype('float32')).unsqueeze(0))
        self.register_buffer('init_cam', torch.from_numpy(mean_params['cam']).unsqueeze(0))
        self._pack = None
        self._pack_key = None

    # ------------------------------------------------------------------ packing
    def _key(self):
        ts = [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.decpose.weight, self.decpose.bias,
              self.decshape.weight, self.decshape.bias, self.deccam.weight, self.deccam.bias,
              self.init_pose, self.init_shape, self.init_cam]
        return (self.precision,) + tuple((t.device, t.data_ptr(), t._version) for t in ts)

    def packed(self):
        key = self._key()
        if self._pack is None or self._pack_key != key:
            w = self.fc1.weight
            nv.require_cuda(w, "Regressor parameters (call .cuda() first)")
            dev = w.device
            f = lambda t: t.detach().to(dev, torch.float32)
            pk = {}
            pk["w1x"] = f(w)[:, :2048].contiguous()
            w1p = torch.zeros(1024, PSC, device=dev)
            w1p[:, :157] = f(w)[:, 2048:]
            pk["w1p"] = w1p
            pk["b1"] = f(self.fc1.bias).contiguous()
            pk["w2"] = f(self.fc2.weight).contiguous()
            pk["b2"] = f(self.fc2.bias).contiguous()
            wdec = torch.zeros(PSC, 1024, device=dev)
            wdec[:144] = f(self.decpose.weight)
            wdec[144:154] = f(self.decshape.weight)
            wdec[154:157] = f(self.deccam.weight)
            pk["wdec"] = wdec
            bdec = torch.zeros(PSC, device=dev)
            bdec[:144] = f(self.decpose.bias)
            bdec[144:154] = f(self.decshape.bias)
            bdec[154:157] = f(self.deccam.bias)
            pk["bdec"] = bdec
            init = torch.zeros(1, PSC, device=dev)
            init[0, :144] = f(self.init_pose)[0]
            init[0, 144:154] = f(self.init_shape)[0]
            init[0, 154:157] = f(self.init_cam)[0]
            pk["init"] = init
            for name in ("w1x", "w1p", "w2", "wdec"):
                pk[name] = nv.pack_linear(pk[name], self.precision)
            pk["c"] = nv.IefWeights(*[nv.ptr(pk[n]) for n in ("w1x", "b1", "w1p", "w2", "b2", "wdec", "bdec")])
            self._pack, self._pack_key = pk, key
        return self._pack

    # ------------------------------------------------------------------ closed form of the IEF loop
    def folded(self, n_iter=3):
        """The IEF loop has NO nonlinearity (lib/models/spin.py:250-261 in eval mode: fc1 -> drop -> fc2 -> drop ->
        dec*, dropout = identity), so n iterations are one affine map of (feature, initial state):

            p' = p + Wd (W2 (W1x f + W1p p + b1) + b2) + bd  =  A p + Q f + c,     A = I + Wd W2 W1p
            p_n = A^n p_0 + S_n (Q f + c),                                         S_n = I + A + ... + A^(n-1)

        Returns fp32 tensors {G [160,2048] = S_n Q, g [160] = S_n c, An [160,160] = A^n} composed in float64
        (rows / columns 157..159 are the zero padding of the [N,160] state).  Opt-in (TePose.fold_linear): the
        result differs from the layer-by-layer evaluation by fp32 rounding only, but it is a different
        sequence of floating-point operations than the reference's."""
        key = (self._key(), int(n_iter))
        if getattr(self, "_fold_key", None) != key:
            with torch.no_grad():
                d = lambda t: t.detach().double()
                W1 = d(self.fc1.weight)
                W1x, W1p = W1[:, :2048], W1[:, 2048:]
                Wd = torch.cat([d(self.decpose.weight), d(self.decshape.weight), d(self.deccam.weight)], dim=0)     # [157,1024]
                bd = torch.cat([d(self.decpose.bias), d(self.decshape.bias), d(self.deccam.bias)])
                WdW2 = Wd @ d(self.fc2.weight)
                A = torch.eye(157, dtype=torch.float64, device=W1.device) + WdW2 @ W1p
                Q = WdW2 @ W1x
                c = Wd @ (d(self.fc2.weight) @ d(self.fc1.bias) + d(self.fc2.bias)) + bd
                S = torch.zeros_like(A)
                An = torch.eye(157, dtype=torch.float64, device=W1.device)
                for _ in range(int(n_iter)):
                    S = S + An
                    An = An @ A
                G = torch.zeros(PSC, 2048, dtype=torch.float64, device=W1.device)
                G[:157] = S @ Q
                g = torch.zeros(PSC, dtype=torch.float64, device=W1.device)
                g[:157] = S @ c
                Anp = torch.zeros(PSC, PSC, dtype=torch.float64, device=W1.device)
                Anp[:157, :157] = An
            self._fold = {"G64": G, "g64": g, "An64": Anp, "G": G.float().contiguous(), "g": g.float().contiguous(),
                          "An": Anp.float().contiguous()}
            self._fold_key = key
        return self._fold

    @nv.device_guard
    def forward_folded(self, x, n_iter=3, is_train=False, J_regressor=None):
        """forward() through the closed form: psc = x . G^T + (g + A^n p_0), one fp32 split-K GEMM."""
        if self.training:
            raise NotImplementedError("tepose_b200.Regressor implements the inference path; call .eval()")
        nv.require_cuda(x, "x")
        pk, fd = self.packed(), self.folded(n_iter)
        feat = x.detach().contiguous().float()
        bias = (fd["g64"] + fd["An64"] @ pk["init"][0].double()).float().contiguous()
        psc = folded_gemm(feat, fd["G"], bias, relu_a=False)
        nv.mark("k3_ief")
        return self.decode(psc, is_train=is_train, J_regressor=J_regressor)

    # ------------------------------------------------------------------ forward
    @nv.device_guard
    def forward(self, x, init_pose=None, init_shape=None, init_cam=None, n_iter=3, is_train=False, J_regressor=None):
        if self.training:
            raise NotImplementedError(
                "tepose_b200.Regressor implements the inference path (dropout = identity); call .eval(). "
                "The reference's train-mode dropout (lib/models/spin.py:256-258) is not implemented yet.")
        nv.require_cuda(x, "x")
        pk = self.packed()
        dev = x.device
        N = x.shape[0]
        feat = x.detach().contiguous().float()
        feat_lp = getattr(x, "_tp_bf16", None) if self.precision == "bf16" else None
        if feat_lp is not None and (feat_lp.shape != feat.shape or not feat_lp.is_contiguous()):
            feat_lp = None
        if init_pose is None and init_shape is None and init_cam is None:
            init, init_rows = pk["init"], 1
        else:
            init = pk["init"].expand(N, -1).clone()
            if init_pose is not None:
                init[:, :144] = init_pose
            if init_shape is not None:
                init[:, 144:154] = init_shape
            if init_cam is not None:
                init[:, 154:157] = init_cam
            init_rows = N
        L = nv.lib()
        psc = torch.empty(N, PSC, device=dev, dtype=torch.float32)
        ws = nv.workspace(L.tp_ief_workspace_bytes(N), dev)
        nv.check(L.tp_ief_forward(nv.PRECISIONS[self.precision], pk["c"], nv.ptr(feat), nv.ptr(feat_lp), N, nv.ptr(init), init_rows, n_iter, nv.ptr(psc),
                                  nv.ptr(ws), ws.numel(), nv.stream()), "tp_ief_forward")
        nv.mark("k3_ief")
        return self.decode(psc, is_train=is_train, J_regressor=J_regressor)

    @nv.device_guard
    def decode(self, psc: torch.Tensor, is_train=False, J_regressor=None):
        """rot6d -> R, SMPL, (H36M regression), projection, R -> axis-angle, theta
        (lib/models/spin.py:263-291) in one tp_smpl_forward call on the IEF state [N,160]."""
        smpl = self.smpl
        p = smpl.packed()
        N = psc.shape[0]
        if (not is_train) and J_regressor is not None:
            jreg, src = smpl.h36m_tables(J_regressor)
        else:
            jreg, src = smpl._jreg_extra, smpl._src49
        pose = psc                      # columns   0..143, row stride 160
        betas = psc[:, 144:]            # columns 144..153
        cam = psc[:, 154:]              # columns 154..156
        verts, joints, kp2d, rotmat, theta = smpl_forward_native(
            p, pose, PSC, nv.POSE_ROT6D, betas, PSC, cam, PSC, N, jreg, src, want_theta=True,
            blend_mode=1 if self.precision == "bf16" else 0)
        nv.mark("k45_smpl")
        return [{'theta': theta, 'verts': verts, 'kp_2d': kp2d, 'kp_3d': joints, 'rotmat': rotmat}]

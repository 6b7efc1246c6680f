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

    # ------------------------------------------------------------------ forward
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

"""VIBE bootstrap model with the reference's module API (lib/models/vibe.py:27-117) on the sm_100a kernels.

evaluate.py:89-99,234 runs it once per sequence to seed the theta slots of the first TePose window:

  TemporalEncoder(n_layers=1, hidden_size=2048, add_linear=False, bidirectional=False, use_residual=True)
  VIBE(seqlen, batch_size=64, n_layers=1, hidden_size=2048, add_linear=False, bidirectional=False,
       use_residual=True, pretrained=...)
  VIBE.forward(input [B,T,2048], J_regressor=None) -> [ {theta [B,T,85], verts [B,T,6890,3], kp_2d, kp_3d, rotmat} ]

Same kernels as TePose: K1 input projection (tcgen05 / FFMA), K2 persistent recurrence writing EVERY
step's state, the relu + Linear as one GEMM over all T*B rows, the residual + TNF->NTF permute as one
pass (tp_unpack_rows_residual), then Regressor / SMPL over N = B*T rows.  nn.GRU / nn.Linear members
are parameter containers (same state_dict keys as the reference); their forward is never called.
"""
from __future__ import annotations

import os
import os.path as osp

import torch
import torch.nn as nn

from . import _native as nv
from .smpl import BASE_DATA_DIR
from .spin import Regressor
from .tepose import GruKernels, _round_up

FEAT = 2048


class TemporalEncoder(nn.Module, GruKernels):
    def __init__(self, n_layers=1, hidden_size=2048, add_linear=False, bidirectional=False, use_residual=True,
                 precision="fp32"):
        super().__init__()
        if hidden_size % 32 != 0:
            raise ValueError("tepose_b200 needs hidden_size to be a multiple of 32")
        self.gru = nn.GRU(input_size=FEAT, hidden_size=hidden_size, bidirectional=bidirectional, num_layers=n_layers)
        self.linear = None                                     # lib/models/vibe.py:45-49
        if bidirectional:
            self.linear = nn.Linear(hidden_size * 2, FEAT)
        elif add_linear:
            self.linear = nn.Linear(hidden_size, FEAT)
        self.use_residual = use_residual
        self.hidden_size = hidden_size
        self.n_layers = n_layers
        self.n_dir = 2 if bidirectional else 1
        self.precision = precision
        self._pack = None
        self._pack_key = None

    # ------------------------------------------------------------------ packing
    def _key(self):
        ts = list(self.parameters())
        return (self.precision,) + tuple((t.device, t.data_ptr(), t._version) for t in ts)

    def packed(self):
        key = self._key()
        if self._pack is not None and self._pack_key == key:
            return self._pack
        w0 = self.gru.weight_ih_l0
        nv.require_cuda(w0, "encoder parameters (call .cuda() first)")
        dev, H, D = w0.device, self.hidden_size, self.n_dir
        lp = self.precision == "bf16"
        wdt = torch.bfloat16 if lp else torch.float32
        g = lambda name: getattr(self.gru, name).detach().to(dev, torch.float32)
        sfx = ["", "_reverse"][:D]

        def pad_k(w, kp):
            out = torch.zeros(w.shape[0], kp, device=dev, dtype=wdt)
            out[:, :w.shape[1]] = w.to(wdt)
            return out

        layers = []
        for l in range(self.n_layers):
            kin = FEAT if l == 0 else D * H
            kp = _round_up(kin, 64) if lp else kin
            layers.append({
                "kp": kp,
                "w_ih": torch.cat([pad_k(g(f"weight_ih_l{l}{s}"), kp) for s in sfx], dim=0).contiguous(),    # [D*3H, kp]
                "b_ih": torch.cat([g(f"bias_ih_l{l}{s}") for s in sfx]).contiguous(),
                "w_hh": [self._pack_whh(g(f"weight_hh_l{l}{s}"), lp) for s in sfx],
                "b_hh": [g(f"bias_hh_l{l}{s}").contiguous() for s in sfx],
            })
        pk = {"layers": layers, "kp_out": _round_up(D * H, 64) if lp else D * H}
        if self.linear is not None:
            pk["w_lin"] = pad_k(self.linear.weight.detach().to(dev, torch.float32), pk["kp_out"]).contiguous()
            pk["b_lin"] = self.linear.bias.detach().to(dev, torch.float32).contiguous()
        self._pack, self._pack_key = pk, key
        return pk

    # ------------------------------------------------------------------ forward
    @nv.device_guard
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,T,2048] -> [B,T,F] (lib/models/vibe.py:52-65); F = 2048 with a linear layer, else D*H."""
        nv.require_cuda(x, "input")
        if x.dim() != 3 or x.shape[2] != FEAT:
            raise ValueError(f"expected input [B,T,{FEAT}], got {tuple(x.shape)}")
        pk = self.packed()
        L = nv.lib()
        dev, H, D, Ln = x.device, self.hidden_size, self.n_dir, self.n_layers
        B, T = x.shape[0], x.shape[1]
        lp = self.precision == "bf16"
        adt = torch.bfloat16 if lp else torch.float32
        prec = nv.PRECISIONS[self.precision]
        x = x.detach().float()
        if x.stride(2) != 1:
            x = x.contiguous()
        M = T * B
        src = torch.empty(M, FEAT, device=dev, dtype=adt)                       # time-major rows t*B + b
        nv.check(L.tp_pack_rows(nv.vp(x.data_ptr()), x.stride(0), x.stride(1), B, T, FEAT, nv.ptr(src), FEAT, prec, 0,
                                nv.stream()), "tp_pack_rows")
        nv.mark("pack")
        y = None
        for l in range(Ln):
            d = pk["layers"][l]
            last = l == Ln - 1
            gi = torch.empty(M, D * 3 * H, device=dev, dtype=torch.float32)
            self._input_proj(src, M, d["w_ih"], d["kp"], d["b_ih"], [(0, M, 0, D * 3 * H)], [gi])
            nv.mark(f"k1_input_proj_l{l}")
            kpn = pk["kp_out"] if last else pk["layers"][l + 1]["kp"]
            alloc = torch.zeros if kpn != D * H else torch.empty                # K padding must read as zero
            y = alloc(M, kpn, device=dev, dtype=torch.float32)
            y_lp = alloc(M, kpn, device=dev, dtype=torch.bfloat16) if (lp and not last) else None
            jobs = [self._job(dev, gi, 0, d["w_hh"][0], d["b_hh"][0], T, 0, 1, y=y, ycol=0, y_lp=y_lp)]
            if D == 2:      # the reverse direction walks t = T-1..0 and stores its state at the frame it read
                jobs.append(self._job(dev, gi, 3 * H, d["w_hh"][1], d["b_hh"][1], T, T - 1, -1, y=y, ycol=H, y_lp=y_lp,
                                      t_out0=T - 1, t_out_step=-1))
            self._recurrence(jobs, B)
            nv.mark(f"k2_recurrence_l{l}")
            src = y_lp if lp else y
        F = D * H
        if self.linear is not None:                                             # y = linear(relu(y))
            F = FEAT
            lin = torch.empty(M, FEAT, device=dev, dtype=torch.float32)
            kp = pk["kp_out"]
            if lp:
                a = torch.empty(M, kp, device=dev, dtype=torch.bfloat16)
                nv.check(L.tp_pack_rows(nv.ptr(y), kp, 0, M, 1, kp, nv.ptr(a), kp, prec, 1, nv.stream()), "tp_pack_rows")
                arr = (nv.GemmSeg * 1)(nv.GemmSeg(0, M, 0, FEAT, nv.ptr(lin), FEAT, nv.ptr(pk["b_lin"])))
                nv.check(L.tp_gemm_bf16_tc(nv.ptr(a), M, nv.ptr(pk["w_lin"]), FEAT, kp, arr, 1, nv.stream()), "tp_gemm_bf16_tc")
            else:
                nv.check(L.tp_gemm_f32(nv.ptr(y), kp, nv.ptr(pk["w_lin"]), kp, nv.ptr(pk["b_lin"]), nv.vp(0), 0, nv.ptr(lin),
                                       FEAT, M, FEAT, kp, 1.0, 0.0, 1, nv.stream()), "tp_gemm_f32")
            y = lin
            nv.mark("vibe_linear")
        res = x if (self.use_residual and F == FEAT) else None                  # lib/models/vibe.py:60-61
        out = torch.empty(B, T, F, device=dev, dtype=torch.float32)
        out_lp = torch.empty(B, T, F, device=dev, dtype=torch.bfloat16) if lp else None
        nv.check(L.tp_unpack_rows_residual(nv.ptr(y), y.shape[1], nv.vp(0 if res is None else res.data_ptr()),
                                           x.stride(0), x.stride(1), B, T, F, nv.ptr(out), nv.ptr(out_lp), nv.stream()),
                 "tp_unpack_rows_residual")
        nv.mark("vibe_residual")
        out._tp_bf16 = out_lp
        return out


class VIBE(nn.Module):
    def __init__(self, seqlen, batch_size=64, n_layers=1, hidden_size=2048, add_linear=False, bidirectional=False,
                 use_residual=True, pretrained=osp.join(BASE_DATA_DIR, 'spin_model_checkpoint.pth.tar'), precision="fp32"):
        super().__init__()
        if precision not in nv.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(nv.PRECISIONS)}")
        self.seqlen = seqlen
        self.batch_size = batch_size
        self.encoder = TemporalEncoder(n_layers=n_layers, hidden_size=hidden_size, bidirectional=bidirectional,
                                       add_linear=add_linear, use_residual=use_residual, precision=precision)
        self.regressor = Regressor(precision=precision)
        if pretrained and os.path.isfile(pretrained):          # lib/models/vibe.py:97-101
            pretrained_dict = torch.load(pretrained)['model']
            self.regressor.load_state_dict(pretrained_dict, strict=False)
            print(f'=> loaded pretrained model from \'{pretrained}\'')

    @nv.device_guard
    def forward(self, input, J_regressor=None):
        if self.training:
            raise NotImplementedError("tepose_b200.VIBE implements the inference path; call .eval() first")
        batch_size, seqlen = input.shape[:2]
        nv.mark("start")
        feature = self.encoder(input)
        lp = getattr(feature, "_tp_bf16", None)
        feature = feature.reshape(-1, feature.size(-1))
        if lp is not None:
            feature._tp_bf16 = lp.reshape(-1, lp.size(-1))
        smpl_output = self.regressor(feature, J_regressor=J_regressor)
        for s in smpl_output:                                  # lib/models/vibe.py:112-117
            s['theta'] = s['theta'].reshape(batch_size, seqlen, -1)
            s['verts'] = s['verts'].reshape(batch_size, seqlen, -1, 3)
            s['kp_2d'] = s['kp_2d'].reshape(batch_size, seqlen, -1, 2)
            s['kp_3d'] = s['kp_3d'].reshape(batch_size, seqlen, -1, 3)
            s['rotmat'] = s['rotmat'].reshape(batch_size, seqlen, -1, 3, 3)
        return smpl_output

"""CPU restatement of TePose's TRAIN-mode forward and its gradients (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product never does).

What it restates (file:line relative to the reference repository):
  * lib/models/tepose.py:71-87,121-147  encoder in train mode: the fwd-feature and the rec-feature are regressed
    separately -> 2 rows per sequence (`is_train=True`, `torch.cat((y_fwd[:, None], y_rec[:, None]), dim=1)`);
  * lib/models/spin.py:240-291          Regressor.forward under `.train()`: `drop1` / `drop2` (nn.Dropout(), p = 0.5)
    are ACTIVE between fc1 / fc2 / dec* in each of the 3 IEF iterations;
  * lib/core/trainer.py:203,235-237     `generator(inp, is_train=True)`, `zero_grad / backward / step`.
Dropout masks are INPUTS: mask tensors [n_iter, 2, N, 1024] of {0,1}; the layer output is
`a * mask / (1 - p)`, which is what nn.Dropout computes for the mask it draws.  Gradients come from torch.autograd over
the torch-CPU ops the reference itself runs (nn.GRU, F.linear, the LBS restatement of oracle/torch_ref.py).

Pinned by tests/golden/train_*.npz: outputs, loss and per-parameter gradient probes produced by the UNMODIFIED
reference modules in train mode (oracle/ref_harness.py:run_reference_train forces the same masks through forward hooks
on `regressor.drop1/drop2`).  The SMPL part goes through the smplx stand-in on both sides: parity unpinned there.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import synth, torch_ref

DROP_P = 0.5            # nn.Dropout() default, lib/models/spin.py:216-218
N_ITER = 3


def make_masks(seed: int, n_rows: int, n_iter: int = N_ITER) -> np.ndarray:
    """Bernoulli(1 - p) keep masks [n_iter, 2 (drop1, drop2), n_rows, 1024] as float32 {0,1} (numpy PCG64)."""
    g = synth._rng(seed, 900)
    return (g.random((n_iter, 2, n_rows, 1024)) >= DROP_P).astype(np.float32)


def make_targets(seed: int, n_rows: int) -> dict:
    """Synthetic regression targets for the outputs that carry gradient in TePoseLoss (lib/core/loss.py:93-131:
    kp_2d, kp_3d[:, 25:39], theta)."""
    g = synth._rng(seed, 901)
    return {"kp_2d": g.standard_normal((n_rows, 49, 2)).astype(np.float32) * 0.5,
            "kp_3d": g.standard_normal((n_rows, 14, 3)).astype(np.float32) * 0.3,
            "theta": g.standard_normal((n_rows, 85)).astype(np.float32) * 0.2}


def synthetic_loss(out: dict, tgt: dict) -> torch.Tensor:
    """Stand-in for TePoseLoss on the same outputs (the data-dependent terms need licensed datasets): MSE on kp_2d,
    on the 14 common 3-D joints (lib/core/loss.py:99 `[:, 25:39]`) and on theta."""
    t = lambda k: torch.as_tensor(tgt[k], dtype=out[k].dtype, device=out[k].device)
    n = out["kp_2d"].shape[0] * out["kp_2d"].shape[1] if out["kp_2d"].dim() == 4 else out["kp_2d"].shape[0]
    kp2 = out["kp_2d"].reshape(n, 49, 2)
    kp3 = out["kp_3d"].reshape(n, 49, 3)[:, 25:39]
    th = out["theta"].reshape(n, 85)
    return ((kp2 - t("kp_2d")) ** 2).mean() + ((kp3 - t("kp_3d")) ** 2).mean() * 10.0 + ((th - t("theta")) ** 2).mean()


def ief_forward_train(W: dict, feat: torch.Tensor, masks: torch.Tensor, init: tuple):
    """lib/models/spin.py:250-261 with the dropout layers applied as given masks."""
    pose, shape, cam = init
    scale = 1.0 / (1.0 - DROP_P)
    for i in range(masks.shape[0]):
        xc = torch.cat([feat, pose, shape, cam], 1)
        xc = F.linear(xc, W["fc1.weight"], W["fc1.bias"]) * masks[i, 0] * scale
        xc = F.linear(xc, W["fc2.weight"], W["fc2.bias"]) * masks[i, 1] * scale
        pose = F.linear(xc, W["decpose.weight"], W["decpose.bias"]) + pose
        shape = F.linear(xc, W["decshape.weight"], W["decshape.bias"]) + shape
        cam = F.linear(xc, W["deccam.weight"], W["deccam.bias"]) + cam
    return pose, shape, cam


class TrainOracle:
    """Holds torch-CPU parameters (requires_grad) named like the reference's state_dict and runs forward / backward."""

    def __init__(self, sd: dict, smpl_seed: int, n_layers: int, hidden: int):
        self.n_layers, self.hidden = n_layers, hidden
        self.m = torch_ref.SmplModel.synthetic(smpl_seed)
        self.gru_fwd = torch_ref.build_gru(sd, "gru_fwd", n_layers, hidden, False)
        self.gru_rec = torch_ref.build_gru(sd, "gru_rec", n_layers, hidden, True)
        self.lin = {}
        for k in ("encoder.linear_fwd.weight", "encoder.linear_fwd.bias", "encoder.linear_rec.weight", "encoder.linear_rec.bias"):
            self.lin[k] = torch.as_tensor(sd[k], dtype=torch.float32).clone().requires_grad_(True)
        self.reg = {}
        for n in ("fc1", "fc2", "decpose", "decshape", "deccam"):
            for s in ("weight", "bias"):
                self.reg[f"{n}.{s}"] = torch.as_tensor(sd[f"regressor.{n}.{s}"], dtype=torch.float32).clone().requires_grad_(True)
        self.init = tuple(torch.as_tensor(sd[f"regressor.init_{k}"], dtype=torch.float32) for k in ("pose", "shape", "cam"))

    def named_parameters(self):
        for name, mod in (("gru_fwd", self.gru_fwd), ("gru_rec", self.gru_rec)):
            for k, p in mod.named_parameters():
                yield f"encoder.{name}.{k}", p
        for k, p in self.lin.items():
            yield k, p
        for k, p in self.reg.items():
            yield f"regressor.{k}", p

    def forward(self, x: torch.Tensor, masks: torch.Tensor) -> dict:
        B = x.shape[0]
        xt = x.permute(1, 0, 2)
        y, _ = self.gru_fwd(xt)                                          # tepose.py:73
        y_rec, _ = self.gru_rec(torch.flip(xt, dims=[0]))                # tepose.py:75-76
        a = F.linear(F.relu(y[-1]), self.lin["encoder.linear_fwd.weight"], self.lin["encoder.linear_fwd.bias"])
        b = F.linear(F.relu(y_rec[0]), self.lin["encoder.linear_rec.weight"], self.lin["encoder.linear_rec.bias"])
        feat = torch.stack([a, b], dim=1).reshape(2 * B, -1)             # tepose.py:85,126
        N = feat.shape[0]
        init = tuple(t.expand(N, -1) for t in self.init)
        pose, shape, cam = ief_forward_train(self.reg, feat, masks, init)
        R = torch_ref.rot6d_to_rotmat(pose).reshape(N, 24, 3, 3)
        verts, joints, _ = torch_ref.smpl_forward(self.m, shape, R=R)
        kp2d = torch_ref.projection(joints, cam)
        aa = torch_ref.rotmat_to_angle_axis(R.reshape(-1, 3, 3)).reshape(N, 72)
        out = {"theta": torch.cat([cam, aa, shape], dim=1), "verts": verts, "kp_2d": kp2d, "kp_3d": joints, "rotmat": R}
        return {"theta": out["theta"].reshape(B, 2, -1), "verts": out["verts"].reshape(B, 2, -1, 3),
                "kp_2d": out["kp_2d"].reshape(B, 2, -1, 2), "kp_3d": out["kp_3d"].reshape(B, 2, -1, 3),
                "rotmat": out["rotmat"].reshape(B, 2, -1, 3, 3)}

    def loss_and_grads(self, x, masks, tgt):
        for _, p in self.named_parameters():
            p.grad = None
        out = self.forward(torch.as_tensor(x), torch.as_tensor(masks))
        loss = synthetic_loss(out, tgt)
        loss.backward()
        grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in self.named_parameters()}
        return {k: v.detach() for k, v in out.items()}, float(loss.detach()), grads


def grad_probe(g: torch.Tensor, n: int = 32) -> np.ndarray:
    """[sum, sum |.|, n evenly spaced entries] of a gradient tensor (what the golden files keep per parameter)."""
    f = g.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return np.concatenate([[float(f.sum()), float(f.abs().sum())], f[idx].numpy()]).astype(np.float64)

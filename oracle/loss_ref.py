"""CPU restatement of the generator loss's data terms (lib/core/loss.py) -- TEST INFRASTRUCTURE.
Plain torch ops, differentiated by torch.autograd; pinned by tests/golden/loss_*.npz, which hold the values and gradients of the
UNMODIFIED reference `TePoseLoss` (oracle/make_golden.py:make_loss_golden)."""
from __future__ import annotations

import numpy as np
import torch

from . import torch_ref


def make_loss_case(seed: int, batch: int = 6, n2d: int = 2, seqlen: int = 2):
    """Synthetic generator outputs + dataset dicts shaped like lib/core/trainer.py:155-221 feeds the loss: `n2d` rows from the 2-D
    dataset first, then `batch - n2d` rows from the 3-D dataset, `seqlen` = the [B, 2, ...] pair of the train-mode output."""
    g = np.random.Generator(np.random.PCG64([seed, 77]))
    f = lambda *s: torch.from_numpy(g.standard_normal(s).astype(np.float32))
    n3d = batch - n2d
    out = {"theta": f(batch, seqlen, 85) * 0.3, "kp_2d": f(batch, seqlen, 49, 2) * 0.5, "kp_3d": f(batch, seqlen, 49, 3) * 0.3}
    kp2 = lambda n: torch.cat([f(n, seqlen, 49, 2) * 0.5, torch.from_numpy(g.random((n, seqlen, 49, 1)).astype(np.float32))], dim=-1)
    data_2d = {"kp_2d": kp2(n2d)}
    w_3d = torch.from_numpy((g.random((n3d, seqlen)) > 0.3).astype(np.float32))
    w_smpl = torch.from_numpy(np.repeat((g.random((n3d, 1)) > 0.4), seqlen, axis=1).astype(np.float32))
    data_3d = {"kp_2d": kp2(n3d), "kp_3d": f(n3d, seqlen, 49, 3) * 0.3, "theta": f(n3d, seqlen, 85) * 0.3, "w_3d": w_3d, "w_smpl": w_smpl}
    pre_mosh = f(batch, 3, 85) * 0.3
    data_motion_mosh = {"theta": f(batch, 4, 85) * 0.3}
    return out, data_2d, data_3d, pre_mosh, data_motion_mosh


class StubDiscriminator(torch.nn.Module):
    """A small stand-in for the caller-supplied motion discriminator ([N, T, 72] -> [N, 1]): fixed weights from the seed."""

    def __init__(self, seed: int = 0):
        super().__init__()
        g = np.random.Generator(np.random.PCG64([seed, 78]))
        self.w = torch.nn.Parameter(torch.from_numpy(g.standard_normal((72, 1)).astype(np.float32)) * 0.05)

    def forward(self, x):
        return torch.tanh(x.mean(dim=1) @ self.w)


def data_terms(pred_j2d, real_2d, pred_j3d, real_3d, pred_theta, real_theta, weights=(60., 30., 1., 0.001)):
    """loss.py:106-126: (loss_kp_2d, loss_kp_3d, loss_pose, loss_shape), each already multiplied by its weight."""
    conf = real_2d[:, :, -1:].clone()                                                   # keypoint_loss, loss.py:179-192 (weights 1, 1)
    l2d = (conf * (pred_j2d - real_2d[:, :, :-1]) ** 2).mean() * weights[0] if len(real_2d) else torch.zeros(())
    if len(real_3d):                                                                    # keypoint_3d_loss, loss.py:194-217
        P, G = pred_j3d[:, 25:39], real_3d[:, 25:39]
        G = G - ((G[:, 2] + G[:, 3]) / 2)[:, None]
        P = P - ((P[:, 2] + P[:, 3]) / 2)[:, None]
        l3d = ((P - G) ** 2).mean() * weights[1]
    else:
        l3d = torch.zeros(())
    if len(pred_theta):                                                                 # smpl_losses, loss.py:219-231
        Rp = torch_ref.batch_rodrigues_quat(pred_theta[:, 3:75].reshape(-1, 3)).reshape(-1, 24, 3, 3)
        Rg = torch_ref.batch_rodrigues_quat(real_theta[:, 3:75].reshape(-1, 3)).reshape(-1, 24, 3, 3)
        lpose = ((Rp - Rg) ** 2).mean() * weights[2]
        lshape = ((pred_theta[:, 75:] - real_theta[:, 75:]) ** 2).mean() * weights[3]
    else:
        lpose = lshape = torch.zeros(())
    return l2d, l3d, lpose, lshape


def tepose_loss(generator_outputs, data_2d, data_3d, pre_mosh, data_motion_mosh, motion_discriminator,
                weights=(60., 30., 1., 0.001), d_motion_loss_weight=1.):
    """TePoseLoss.forward (loss.py:59-171) restated around data_terms(): returns (gen_loss, motion_dis_loss, loss_dict)."""
    reduce = lambda x: x.contiguous().view((x.shape[0] * x.shape[1],) + x.shape[2:])
    n2 = data_2d["kp_2d"].shape[0] if data_2d else 0
    real_2d = reduce(torch.cat((data_2d["kp_2d"], data_3d["kp_2d"]), 0) if data_2d else data_3d["kp_2d"])
    w_3d, w_smpl = data_3d["w_3d"].bool().reshape(-1), data_3d["w_smpl"].bool().reshape(-1)
    preds = generator_outputs[-1]
    l2d, l3d, lpose, lshape = data_terms(reduce(preds["kp_2d"]), real_2d, reduce(preds["kp_3d"][n2:])[w_3d], reduce(data_3d["kp_3d"])[w_3d],
                                         reduce(preds["theta"][n2:])[w_smpl], reduce(data_3d["theta"])[w_smpl], weights)
    d = {"loss_kp_2d": l2d, "loss_kp_3d": l3d}
    if int(w_smpl.sum()) > 0:
        d["loss_shape"], d["loss_pose"] = lshape, lpose
    thetas = torch.cat([o["theta"] for o in generator_outputs], 0)
    pm = torch.cat((pre_mosh, thetas.mean(dim=1, keepdim=True)), dim=1)
    pm = torch.cat((pm[:n2], pm[n2:][~w_smpl[::2]]), dim=0)
    rm = data_motion_mosh["theta"]
    rm = torch.cat((rm[:n2], rm[n2:][~w_smpl[::2]]), dim=0)
    if pm.shape[0] > 0:
        k = lambda v: v.shape[0]
        e = motion_discriminator(pm[:, :, 3:75])
        d["e_m_disc_loss"] = torch.sum((e - 1.0) ** 2) / k(e) * d_motion_loss_weight
        gen = torch.stack(list(d.values())).sum()
        fake, real = motion_discriminator(pm.detach()[:, :, 3:75]), motion_discriminator(rm[:, :, 3:75])
        la, lb = torch.sum((real - 1) ** 2) / k(real), torch.sum(fake ** 2) / k(fake)
        d["d_m_disc_real"], d["d_m_disc_fake"], d["d_m_disc_loss"] = la * d_motion_loss_weight, lb * d_motion_loss_weight, (la + lb) * d_motion_loss_weight
        return gen, d["d_m_disc_loss"], d
    return torch.stack(list(d.values())).sum(), torch.zeros(1), d

"""Drop-in for lib.core.loss.TePoseLoss (reference lib/core/loss.py:33-171): same constructor arguments, same forward
signature, same return value (gen_loss, motion_dis_loss, loss_dict).

What runs where:
  * the DATA terms (loss_kp_2d, loss_kp_3d, loss_pose, loss_shape: loss.py:106-126 with keypoint_loss :179-192, keypoint_3d_loss
    :194-217 and smpl_losses :219-231) are ONE native call, tp_tepose_loss (csrc/loss.cu), that returns the four weighted values and
    their gradient w.r.t. the generator outputs; `_DataTerms` hands that gradient to autograd, so `gen_loss.backward()` continues
    into the hand-written backward of the generator (tepose_b200/train.py) exactly like the reference's loss does into autograd;
  * the row selection by `w_3d` / `w_smpl` and the concatenations (loss.py:75-104) are index plumbing done with torch;
  * the adversarial terms (loss.py:128-160) call the `motion_discriminator` module the CALLER passes in, with torch ops, as the
    reference does -- the MS-G3D discriminator itself (lib/models/motion_discriminator_gcn.py) is not part of this package.  With
    `motion_discriminator=None` the adversarial terms are left out (the reference would fail there).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _native as nv


class _DataTerms(torch.autograd.Function):
    """losses[4] = tp_tepose_loss(kp2d, real2d, kp3d, real3d, theta, real_theta); the kernel also returns d(sum)/d(prediction)."""

    @staticmethod
    def forward(ctx, kp2d, real2d, kp3d, real3d, theta, real_theta, weights):
        dev = kp2d.device
        f = lambda t: t.detach().to(dev, torch.float32).contiguous()
        kp2d_c, real2d_c, kp3d_c, real3d_c, theta_c, real_theta_c = map(f, (kp2d, real2d, kp3d, real3d, theta, real_theta))
        n2, n3, ns = kp2d_c.shape[0], kp3d_c.shape[0], theta_c.shape[0]
        L = nv.lib()
        losses = torch.empty(4, device=dev, dtype=torch.float32)
        g2, g3, gs = torch.empty_like(kp2d_c), torch.empty_like(kp3d_c), torch.empty_like(theta_c)
        ws = nv.workspace(L.tp_tepose_loss_workspace_bytes(n2, n3, ns), dev)
        w6 = (C.c_float * 6)(*[float(v) for v in weights])
        P = lambda t: nv.vp(t.data_ptr()) if t.numel() else nv.vp(0)
        with torch.cuda.device(dev):
            nv.check(L.tp_tepose_loss(P(kp2d_c), P(real2d_c), n2, P(kp3d_c), P(real3d_c), n3, P(theta_c), P(real_theta_c), ns, w6,
                                      nv.ptr(losses), P(g2), P(g3), P(gs), nv.ptr(ws), ws.numel(), nv.stream()), "tp_tepose_loss")
        ctx.save_for_backward(g2, g3, gs)
        return losses

    @staticmethod
    def backward(ctx, g_losses):
        # the kernel's gradient arrays are separable per term except theta, which carries pose + shape together: those two terms
        # must arrive with the same upstream factor (they do: gen_loss is a plain sum, loss.py:145)
        g2, g3, gs = ctx.saved_tensors
        if gs.numel() and not bool(g_losses[2] == g_losses[3]):
            raise NotImplementedError("tepose_b200.TePoseLoss: loss_pose and loss_shape must be summed with equal weight (as loss.py:145 does)")
        return g2 * g_losses[0], None, g3 * g_losses[1], None, gs * g_losses[2], None, None


class TePoseLoss(nn.Module):
    def __init__(self, e_loss_weight=60., e_3d_loss_weight=30., e_pose_loss_weight=1., e_shape_loss_weight=0.001,
                 d_motion_loss_weight=1., device='cuda'):
        super().__init__()
        self.e_loss_weight = e_loss_weight
        self.e_3d_loss_weight = e_3d_loss_weight
        self.e_pose_loss_weight = e_pose_loss_weight
        self.e_shape_loss_weight = e_shape_loss_weight
        self.d_motion_loss_weight = d_motion_loss_weight
        self.device = device
        self.enc_loss = batch_encoder_disc_l2_loss
        self.dec_loss = batch_adv_disc_l2_loss

    def forward(self, generator_outputs, data_2d, data_3d, pre_mosh=None, data_body_mosh=None, data_motion_mosh=None,
                body_discriminator=None, motion_discriminator=None):
        reduce = lambda x: x.contiguous().view((x.shape[0] * x.shape[1],) + x.shape[2:])          # loss.py:71
        if data_2d:
            sample_2d_count = data_2d['kp_2d'].shape[0]
            real_2d = torch.cat((data_2d['kp_2d'], data_3d['kp_2d']), 0)
        else:
            sample_2d_count = 0
            real_2d = data_3d['kp_2d']
        real_2d = reduce(real_2d)
        real_3d = reduce(data_3d['kp_3d'])
        data_3d_theta = reduce(data_3d['theta'])
        w_3d = data_3d['w_3d'].type(torch.bool).reshape(-1)
        w_smpl = data_3d['w_smpl'].type(torch.bool).reshape(-1)
        total_predict_thetas = torch.cat([output['theta'] for output in generator_outputs], 0)
        preds = generator_outputs[-1]
        pred_j3d = reduce(preds['kp_3d'][sample_2d_count:])[w_3d]
        pred_theta = reduce(preds['theta'][sample_2d_count:])[w_smpl]
        pred_j2d = reduce(preds['kp_2d'])
        data_3d_theta = data_3d_theta[w_smpl]
        real_3d = real_3d[w_3d]

        losses = _DataTerms.apply(pred_j2d, real_2d, pred_j3d, real_3d, pred_theta, data_3d_theta,
                                  (self.e_loss_weight, self.e_3d_loss_weight, self.e_pose_loss_weight, self.e_shape_loss_weight, 1.0, 1.0))
        loss_dict = {'loss_kp_2d': losses[0], 'loss_kp_3d': losses[1]}
        if pred_theta.shape[0] > 0:
            loss_dict['loss_shape'] = losses[3]
            loss_dict['loss_pose'] = losses[2]

        pred_motion = real_motion = None
        if motion_discriminator is not None:
            pred_motion = torch.cat((pre_mosh, torch.unsqueeze(total_predict_thetas.mean(dim=1), 1)), dim=1)       # loss.py:128-133
            pred_motion = torch.cat((pred_motion[:sample_2d_count], pred_motion[sample_2d_count:][~w_smpl[::2]]), dim=0)
            real_motion = data_motion_mosh['theta']
            real_motion = torch.cat((real_motion[:sample_2d_count], real_motion[sample_2d_count:][~w_smpl[::2]]), dim=0)
        if pred_motion is not None and pred_motion.shape[0] > 0:
            start_idx, end_idx = 3, 75
            e_motion_disc_loss = self.enc_loss(motion_discriminator(pred_motion[:, :, start_idx:end_idx])) * self.d_motion_loss_weight
            fake_motion = pred_motion.detach()
            fake_disc_value = motion_discriminator(fake_motion[:, :, start_idx:end_idx])
            real_disc_value = motion_discriminator(real_motion[:, :, start_idx:end_idx])
            d_real, d_fake, d_loss = self.dec_loss(real_disc_value, fake_disc_value)
            loss_dict['e_m_disc_loss'] = e_motion_disc_loss
            gen_loss = torch.stack(list(loss_dict.values())).sum()
            loss_dict['d_m_disc_real'] = d_real * self.d_motion_loss_weight
            loss_dict['d_m_disc_fake'] = d_fake * self.d_motion_loss_weight
            loss_dict['d_m_disc_loss'] = d_loss * self.d_motion_loss_weight
            return gen_loss, loss_dict['d_m_disc_loss'], loss_dict
        gen_loss = torch.stack(list(loss_dict.values())).sum()
        return gen_loss, torch.zeros(1).float(), loss_dict


def batch_encoder_disc_l2_loss(disc_value):
    """lib/core/loss.py:234-240."""
    k = disc_value.shape[0]
    return torch.sum((disc_value - 1.0) ** 2) * 1.0 / k


def batch_adv_disc_l2_loss(real_disc_value, fake_disc_value):
    """lib/core/loss.py:243-251."""
    ka, kb = real_disc_value.shape[0], fake_disc_value.shape[0]
    lb, la = torch.sum(fake_disc_value ** 2) / kb, torch.sum((real_disc_value - 1) ** 2) / ka
    return la, lb, la + lb

"""Evaluation metrics of lib/utils/eval_utils.py computed where the predictions already are -- on the device
(SURVEY.md 8(f-3)).  Same function names / argument meaning; inputs are torch tensors (numpy is accepted and
copied to the GPU; CPU torch tensors are rejected like everywhere in this package), results are CUDA tensors (call .cpu().numpy() where the reference returned numpy).

  batch_compute_similarity_transform_torch(S1, S2)            eval_utils.py:287-337
  compute_error_accel_eval(joints_gt, joints_pred, vis=None)  eval_utils.py:110-138   (evaluate.py:442)
  compute_error_accel(joints_gt, joints_pred, vidlen_each, seqlen)   :79-108          (lib/core/tester.py:309)
  compute_accel(joints, vidlen_each, seqlen)                  :53-76                  (lib/core/tester.py:308)
  compute_error_verts(pred_verts, target_verts=None, target_theta=None, smpl=None)    :141-175
  compute_errors(gt3ds, preds)                                :354-378
  pose_metrics(pred_j3ds, target_j3ds, pelvis=(2, 3))         evaluate.py:420-443 in one pass
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nv


def _dev_f32(t, device=None) -> torch.Tensor:
    """fp32 + contiguous.  torch tensors stay where they are (a CPU tensor is rejected by the native layer, as
    everywhere in this package); numpy inputs are copied to `device` / the current GPU."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.asarray(t))
        if device is None and torch.cuda.is_available():
            device = torch.device("cuda", torch.cuda.current_device())
    t = t.detach().to(device=device if device is not None else t.device, dtype=torch.float32)
    nv.require_cuda(t, "metric input")
    return t.contiguous()


def _pelvis(pelvis):
    if pelvis is None:
        return -1, -1
    if isinstance(pelvis, int):
        return pelvis, -1
    return int(pelvis[0]), int(pelvis[1])


@nv.device_guard
def pose_metrics(pred_j3ds, target_j3ds, pelvis=(2, 3), want_aligned=False):
    """evaluate.py:420-443 for one sequence: root-align both joint sets, then per-frame MPJPE, Procrustes-aligned
    MPJPE and acceleration error (zeros at the first / last frame, evaluate.py:441-442), all in metres.
    pred / target [N,J,3].  pelvis: (2, 3) = hip midpoint of the 14-joint set, an int = that joint (mpii3d: J-3),
    None = no alignment."""
    P = _dev_f32(pred_j3ds)
    G = _dev_f32(target_j3ds, P.device)
    if P.shape != G.shape or P.dim() != 3 or P.shape[2] != 3:
        raise ValueError(f"expected matching [N,J,3] joint sets, got {tuple(P.shape)} and {tuple(G.shape)}")
    n, J = P.shape[0], P.shape[1]
    p0, p1 = _pelvis(pelvis)
    if p0 < 0 and isinstance(pelvis, int):
        p0 += J
    L = nv.lib()
    mpjpe = torch.empty(n, device=P.device)
    pa = torch.empty(n, device=P.device)
    aligned = torch.empty_like(P) if want_aligned else None
    nv.check(L.tp_pose_metrics(nv.ptr(P), nv.ptr(G), n, J, p0, p1, nv.ptr(aligned), nv.ptr(mpjpe), nv.ptr(pa), nv.stream()),
             "tp_pose_metrics")
    accel = torch.zeros(n, device=P.device)
    if n >= 3:
        nv.check(L.tp_accel_error(nv.ptr(P), nv.ptr(G), 1, n, J, p0, p1, nv.vp(accel.data_ptr() + 4), nv.stream()), "tp_accel_error")
    out = {"mpjpe": mpjpe, "mpjpe_pa": pa, "accel_err": accel}
    if want_aligned:
        out["aligned"] = aligned
    return out


def batch_compute_similarity_transform_torch(S1, S2):
    """S1, S2 [N,J,3] (or [N,3,J]) -> S1 mapped onto S2 by the best similarity transform (eval_utils.py:287-337)."""
    A = _dev_f32(S1)
    B = _dev_f32(S2, A.device)
    transposed = False
    if A.shape[0] != 3 and A.shape[0] != 2:          # the reference's own layout test (eval_utils.py:294)
        transposed = True                             # [N,J,3] as passed by evaluate.py / tester.py
    else:
        raise ValueError("batch_compute_similarity_transform_torch: pass [N,J,3] joint sets with N > 3")
    assert B.shape == A.shape
    n, J = A.shape[0], A.shape[1]
    out = torch.empty_like(A)
    nv.check(nv.lib().tp_pose_metrics(nv.ptr(A), nv.ptr(B), n, J, -1, -1, nv.ptr(out), nv.vp(0), nv.vp(0), nv.stream()),
             "tp_pose_metrics")
    return out if transposed else out.permute(0, 2, 1)


@nv.device_guard
def compute_error_accel_eval(joints_gt, joints_pred, vis=None):
    """[N,J,3] x 2 -> [N-2] (visible entries only when vis [N] is given: a frame counts if it and its two
    successors are visible, eval_utils.py:128-136)."""
    G = _dev_f32(joints_gt)
    P = _dev_f32(joints_pred, G.device)
    n, J = P.shape[0], P.shape[1]
    out = torch.empty(max(n - 2, 0), device=P.device)
    nv.check(nv.lib().tp_accel_error(nv.ptr(P), nv.ptr(G), 1, n, J, -1, -1, nv.ptr(out), nv.stream()), "tp_accel_error")
    if vis is not None:
        v = torch.as_tensor(np.asarray(vis) if not torch.is_tensor(vis) else vis).to(P.device).bool()
        keep = (v & torch.roll(v, -1) & torch.roll(v, -2))[:-2]
        out = out[keep]
    return out


def _masked_mean(normed: torch.Tensor, vidlen_each, lo: int, tail: int, extra: int):
    """sum_i sum(normed[i, lo : vidlen_i - tail]) / (sum(vidlen) - n * (seqlen + extra) + 1e-8), eval_utils.py:70-76,104-108."""
    vl = torch.as_tensor(np.asarray(vidlen_each) if not torch.is_tensor(vidlen_each) else vidlen_each).to(normed.device).reshape(-1)
    idx = torch.arange(normed.shape[1], device=normed.device)[None]
    mask = (idx >= lo) & (idx < (vl.long()[:, None] - tail))
    total = (normed * mask).sum()
    return total / (vl.sum() - vl.shape[0] * (lo + 1 + extra) + 1e-8)


@nv.device_guard
def compute_accel(joints, vidlen_each, seqlen):
    """joints [S,L,J,3] -> scalar mean acceleration norm over frames seqlen-1 .. vidlen-3 of every sequence."""
    P = _dev_f32(joints)
    S, Ln, J = P.shape[0], P.shape[1], P.shape[2]
    normed = torch.empty(S, max(Ln - 2, 0), device=P.device)
    nv.check(nv.lib().tp_accel_error(nv.ptr(P), nv.vp(0), S, Ln, J, -1, -1, nv.ptr(normed), nv.stream()), "tp_accel_error")
    return _masked_mean(normed, vidlen_each, seqlen - 1, 2, 1)


@nv.device_guard
def compute_error_accel(joints_gt, joints_pred, vidlen_each, seqlen, vis=None):
    """[S,L,J,3] x 2 -> scalar mean acceleration error over frames seqlen-1 .. vidlen-5 (eval_utils.py:79-108)."""
    if vis is not None:
        raise NotImplementedError("per-frame visibility is only supported by compute_error_accel_eval")
    G = _dev_f32(joints_gt)
    P = _dev_f32(joints_pred, G.device)
    S, Ln, J = P.shape[0], P.shape[1], P.shape[2]
    normed = torch.empty(S, max(Ln - 2, 0), device=P.device)
    nv.check(nv.lib().tp_accel_error(nv.ptr(P), nv.ptr(G), S, Ln, J, -1, -1, nv.ptr(normed), nv.stream()), "tp_accel_error")
    return _masked_mean(normed, vidlen_each, seqlen - 1, 4, 3)


@nv.device_guard
def compute_error_verts(pred_verts, target_verts=None, target_theta=None, device=None, smpl=None):
    """Mean per-vertex distance per body [N] (eval_utils.py:141-175).  Without target_verts the target mesh is
    built from target_theta [N,85] (cam | axis-angle pose | betas) with the SMPL forward (pose2rot=True), in
    chunks of 5000 bodies as the reference does."""
    P = _dev_f32(pred_verts, device)
    L = nv.lib()
    n, V = P.shape[0], P.shape[1]
    out = torch.empty(n, device=P.device)
    if target_verts is not None:
        T = _dev_f32(target_verts, P.device)
        assert T.shape == P.shape
        nv.check(L.tp_vertex_error(nv.ptr(T), nv.ptr(P), n, V, nv.ptr(out), nv.stream()), "tp_vertex_error")
        return out
    if smpl is None:
        from .smpl import SMPL, SMPL_MODEL_DIR
        smpl = SMPL(SMPL_MODEL_DIR, batch_size=1).to(P.device)
    th = _dev_f32(target_theta, P.device)
    assert th.shape[0] == n
    for lo in range(0, n, 5000):
        hi = min(n, lo + 5000)
        pose, betas = th[lo:hi, 3:75], th[lo:hi, 75:]
        tv = smpl(betas=betas, body_pose=pose[:, 3:], global_orient=pose[:, :3], pose2rot=True).vertices
        nv.check(L.tp_vertex_error(nv.ptr(tv), nv.vp(P.data_ptr() + 4 * lo * V * 3), hi - lo, V, nv.vp(out.data_ptr() + 4 * lo),
                                   nv.stream()), "tp_vertex_error")
    return out


def align_by_pelvis(joints):
    """[14,3] in LSP order: subtract the hip midpoint (eval_utils.py:340-351)."""
    joints = torch.as_tensor(joints)
    return joints - ((joints[2] + joints[3]) / 2.0)[None]


def compute_errors(gt3ds, preds):
    """MPJPE after pelvis alignment and after Procrustes, per frame (eval_utils.py:354-378)."""
    m = pose_metrics(preds, gt3ds, pelvis=(2, 3))
    return m["mpjpe"], m["mpjpe_pa"]

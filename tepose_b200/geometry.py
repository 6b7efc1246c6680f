"""Rotation utilities with the names / argument meaning of the reference's
lib/utils/geometry.py, executed by the sm_100a kernels in csrc/geometry.cu."""
from __future__ import annotations

import torch

from . import _native as nv


def _prep(x: torch.Tensor, width: int) -> torch.Tensor:
    nv.require_cuda(x, "input")
    return x.reshape(-1, width).contiguous().float()


@nv.device_guard
def rot6d_to_rotmat(x: torch.Tensor) -> torch.Tensor:
    """lib/utils/geometry.py:330-343.  [...,6k] -> [N,3,3] (x is viewed as (-1,3,2))."""
    flat = _prep(x, 6)
    out = torch.empty(flat.shape[0], 3, 3, device=flat.device, dtype=torch.float32)
    nv.check(nv.lib().tp_rot6d_to_rotmat(nv.ptr(flat), nv.ptr(out), flat.shape[0], nv.stream()), "tp_rot6d_to_rotmat")
    return out


@nv.device_guard
def rotation_matrix_to_angle_axis(rotation_matrix: torch.Tensor) -> torch.Tensor:
    """lib/utils/geometry.py:68-97.  [N,3,3] (or [N,3,4], last column ignored) -> [N,3]."""
    if rotation_matrix.shape[-2:] == (3, 4):
        rotation_matrix = rotation_matrix[..., :3]
    flat = _prep(rotation_matrix, 9)
    out = torch.empty(flat.shape[0], 3, device=flat.device, dtype=torch.float32)
    nv.check(nv.lib().tp_rotmat_to_angle_axis(nv.ptr(flat), nv.ptr(out), flat.shape[0], nv.stream()),
             "tp_rotmat_to_angle_axis")
    return out


@nv.device_guard
def batch_rodrigues(axisang: torch.Tensor, form: str = "quat") -> torch.Tensor:
    """form='quat': lib/utils/geometry.py:22-34 (returns [N,9] like the reference);
    form='smplx': smplx.lbs.batch_rodrigues (returns [N,3,3])."""
    flat = _prep(axisang, 3)
    out = torch.empty(flat.shape[0], 3, 3, device=flat.device, dtype=torch.float32)
    code = nv.RODRIGUES_QUAT if form == "quat" else nv.RODRIGUES_SMPLX
    nv.check(nv.lib().tp_batch_rodrigues(nv.ptr(flat), nv.ptr(out), flat.shape[0], code, nv.stream()), "tp_batch_rodrigues")
    return out.reshape(-1, 9) if form == "quat" else out


@nv.device_guard
def projection(pred_joints: torch.Tensor, pred_camera: torch.Tensor) -> torch.Tensor:
    """lib/models/spin.py:307-320.  joints [N,J,3], camera [N,3] -> [N,J,2]."""
    nv.require_cuda(pred_joints, "pred_joints")
    j = pred_joints.contiguous().float()
    c = pred_camera.contiguous().float()
    out = torch.empty(j.shape[0], j.shape[1], 2, device=j.device, dtype=torch.float32)
    nv.check(nv.lib().tp_projection(nv.ptr(j), nv.ptr(c), nv.ptr(out), j.shape[0], j.shape[1], nv.stream()), "tp_projection")
    return out

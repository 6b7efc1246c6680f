"""Shared test helpers: synthetic base_data directory, model construction, comparisons."""
import contextlib
import os
import tempfile

import numpy as np
import torch

from oracle import synth, torch_ref


@contextlib.contextmanager
def base_data_cwd(seed=0):
    """cwd with data/base_data/* (the reference's asset paths are cwd-relative:
    lib/core/config.py:31)."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        synth.write_base_data(os.path.join(tmp, "data", "base_data"), seed)
        os.chdir(tmp)
        try:
            yield tmp
        finally:
            os.chdir(old)


def build_product_model(seed, seqlen, n_layers, hidden, precision="fp32", device="cpu"):
    from tepose_b200.synthetic import build_synthetic_model
    return build_synthetic_model(seed, seqlen, n_layers, hidden, precision, device)


def oracle_forward(seed, sd, x, n_layers, hidden, is_train=False, use_h36m=False):
    m = torch_ref.SmplModel.synthetic(seed)
    return torch_ref.tepose_forward(sd, m, torch.as_tensor(x), n_layers, hidden, is_train=is_train,
                                    J_regressor=m.J_regressor_h36m if use_h36m else None), m


def aa_to_R(aa):
    return torch_ref.batch_rodrigues_smplx(torch.as_tensor(aa, dtype=torch.float32).reshape(-1, 3))


def compare_outputs(got, ref, vert_tol=1e-4, rot_tol=1e-5, kp2d_tol=1e-3, label=""):
    """Tolerances of SURVEY.md 8(c): verts / joints max-abs <= 1e-4 m (fp32), rotmat <= 1e-5,
    theta compared after mapping the axis-angle part through aa -> R, kp_2d <= 1e-3."""
    g = {k: (v.detach().cpu() if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in got.items()}
    r = {k: (v.detach().cpu() if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in ref.items()}
    errs = {}
    for k in ("verts", "kp_3d", "rotmat", "kp_2d"):
        assert g[k].shape == r[k].shape, (label, k, g[k].shape, r[k].shape)
        errs[k] = float((g[k] - r[k]).abs().max())
    tg, tr = g["theta"].reshape(-1, 85), r["theta"].reshape(-1, 85)
    errs["theta_cam_shape"] = float(torch.cat([tg[:, :3] - tr[:, :3], tg[:, 75:] - tr[:, 75:]], 1).abs().max())
    errs["theta_pose_as_R"] = float((aa_to_R(tg[:, 3:75]) - aa_to_R(tr[:, 3:75])).abs().max())
    assert errs["verts"] <= vert_tol, (label, errs)
    assert errs["kp_3d"] <= vert_tol, (label, errs)
    assert errs["rotmat"] <= rot_tol, (label, errs)
    assert errs["kp_2d"] <= kp2d_tol, (label, errs)
    assert errs["theta_cam_shape"] <= rot_tol * 10, (label, errs)
    assert errs["theta_pose_as_R"] <= rot_tol * 10, (label, errs)
    return errs

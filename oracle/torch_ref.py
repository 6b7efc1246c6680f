"""CPU oracle for the TePose per-sequence hot path -- TEST INFRASTRUCTURE ONLY.

A plain torch-CPU fp32 restatement of the reference algorithm, written from the
reference's behaviour (not copied): every function cites the reference lines it
follows (paths relative to /root/reference).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package; the product (``tepose_b200``) never does.

PARITY PIN STATUS
  * encoder / Regressor / rot6d / R->axis-angle / projection / output assembly are
    pinned against the reference's own modules run unmodified in the authoring
    container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
  * the SMPL body model (``smplx==0.1.13``, requirements.txt:7) is a third-party
    dependency that is NOT vendored in /root/reference and not installable offline.
    ``smpl_forward`` below restates smplx's published ``lbs`` algorithm (SURVEY.md
    App. A.6); the reference ships no test or golden vector for it, so that part is
    **parity unpinned** -- guarded only by the float64 cross-check in
    ``oracle/np64.py`` and the invariants in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import synth

# Joint bookkeeping of the SMPL wrapper (lib/models/smpl.py:14-52): the 49 output
# joints are picked from [24 posed joints | 21 vertex picks | 9 extra-regressed].
JOINT_SOURCE_49 = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7,
                   25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                   8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20,
                   47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27]
# lib/models/smpl.py:57-58
H36M_TO_J14 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10]


# --------------------------------------------------------------------------- rotations
def rot6d_to_rotmat(x: torch.Tensor) -> torch.Tensor:
    """lib/utils/geometry.py:330-343.  [N,144] or [N*24,6] -> [N*24,3,3].
    The 6 numbers are read as a (3,2) block: a1 = even entries, a2 = odd entries;
    Gram-Schmidt with F.normalize(eps=1e-6); result columns are (b1,b2,b3)."""
    p = x.reshape(-1, 3, 2)
    a1, a2 = p[..., 0], p[..., 1]
    b1 = a1 / a1.norm(dim=1, keepdim=True).clamp_min(1e-6)
    u = a2 - (b1 * a2).sum(dim=1, keepdim=True) * b1
    b2 = u / u.norm(dim=1, keepdim=True).clamp_min(1e-6)
    b3 = torch.linalg.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], dim=2)


def rotmat_to_quat(R: torch.Tensor) -> torch.Tensor:
    """lib/utils/geometry.py:153-233 on M = R^T; returns (w,x,y,z)*0.5/sqrt(t).
    The reference selects among four candidates with mask-multiply-sum (:220-231);
    all candidates are finite so that equals a plain select."""
    M = R.transpose(1, 2)
    m00, m01, m02 = M[:, 0, 0], M[:, 0, 1], M[:, 0, 2]
    m10, m11, m12 = M[:, 1, 0], M[:, 1, 1], M[:, 1, 2]
    m20, m21, m22 = M[:, 2, 0], M[:, 2, 1], M[:, 2, 2]
    d2 = m22 < 1e-6
    d01 = m00 > m11
    d0n1 = m00 < -m11
    t0 = 1 + m00 - m11 - m22
    t1 = 1 - m00 + m11 - m22
    t2 = 1 - m00 - m11 + m22
    t3 = 1 + m00 + m11 + m22
    q0 = torch.stack([m12 - m21, t0, m01 + m10, m20 + m02], -1)
    q1 = torch.stack([m20 - m02, m01 + m10, t1, m12 + m21], -1)
    q2 = torch.stack([m01 - m10, m20 + m02, m12 + m21, t2], -1)
    q3 = torch.stack([t3, m12 - m21, m20 - m02, m01 - m10], -1)
    c0 = (d2 & d01)[:, None]
    c1 = (d2 & ~d01)[:, None]
    c2 = (~d2 & d0n1)[:, None]
    q = torch.where(c0, q0, torch.where(c1, q1, torch.where(c2, q2, q3)))
    t = torch.where(c0[:, 0], t0, torch.where(c1[:, 0], t1, torch.where(c2[:, 0], t2, t3)))
    return (q / torch.sqrt(t)[:, None]) * 0.5


def quat_to_angle_axis(q: torch.Tensor) -> torch.Tensor:
    """lib/utils/geometry.py:100-150."""
    w, v = q[:, 0], q[:, 1:]
    s2 = (v * v).sum(-1)
    s = torch.sqrt(s2)
    two_theta = 2.0 * torch.where(w < 0, torch.atan2(-s, -w), torch.atan2(s, w))
    k = torch.where(s2 > 0, two_theta / s, torch.full_like(s, 2.0))
    return v * k[:, None]


def rotmat_to_angle_axis(R: torch.Tensor) -> torch.Tensor:
    """lib/utils/geometry.py:68-97; NaN -> 0 element-wise (:96)."""
    aa = quat_to_angle_axis(rotmat_to_quat(R.reshape(-1, 3, 3)))
    return torch.where(torch.isnan(aa), torch.zeros_like(aa), aa)


def batch_rodrigues_quat(aa: torch.Tensor) -> torch.Tensor:
    """lib/utils/geometry.py:22-65 (quaternion form, used by the loss): [N,3] -> [N,3,3]."""
    n = (aa + 1e-8).norm(dim=1, keepdim=True)
    axis = aa / n
    half = 0.5 * n
    q = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack([
        w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
        2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
        2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z,
    ], dim=1).reshape(-1, 3, 3)


def batch_rodrigues_smplx(aa: torch.Tensor) -> torch.Tensor:
    """smplx.lbs.batch_rodrigues (third-party, restated): angle = ||r + 1e-8||,
    k = r/angle, R = I + sin*K + (1-cos)*K@K.   [N,3] -> [N,3,3]."""
    angle = (aa + 1e-8).norm(dim=1, keepdim=True)
    k = aa / angle
    c, s = torch.cos(angle)[:, None], torch.sin(angle)[:, None]
    kx, ky, kz = k.unbind(1)
    z = torch.zeros_like(kx)
    K = torch.stack([z, -kz, ky, kz, z, -kx, -ky, kx, z], dim=1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=aa.dtype, device=aa.device)[None]
    return eye + s * K + (1 - c) * torch.bmm(K, K)


# --------------------------------------------------------------------------- SMPL
class SmplModel:
    """Tensors of an SMPL-format model, laid out as smplx.SMPL keeps them."""

    def __init__(self, model: dict, extra: dict, dtype=torch.float32):
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype)
        self.v_template = t(model["v_template"])                       # [V,3]
        self.shapedirs = t(model["shapedirs"])[:, :, :10]              # [V,3,10]
        V = self.v_template.shape[0]
        self.posedirs = t(model["posedirs"]).reshape(V * 3, -1).T.contiguous()  # [207,3V]
        self.J_regressor = t(model["J_regressor"])                     # [24,V]
        self.lbs_weights = t(model["weights"])                         # [V,24]
        parents = np.asarray(model["kintree_table"])[0].astype(np.int64).copy()
        parents[0] = -1
        self.parents = torch.as_tensor(parents)
        self.extra_vertex_ids = torch.as_tensor(synth.SMPL_EXTRA_VERTEX_IDS)
        self.J_regressor_extra = t(extra["J_regressor_extra"])         # [9,V]
        self.J_regressor_h36m = t(extra["J_regressor_h36m"])           # [17,V]
        self.faces = np.asarray(model["f"]).astype(np.int64)

    @classmethod
    def synthetic(cls, seed: int = 0, dtype=torch.float32):
        return cls(synth.make_smpl_model(seed), synth.make_extra_regressors(seed), dtype)


def smpl_lbs(m: SmplModel, betas: torch.Tensor, R: torch.Tensor):
    """smplx.lbs.lbs with pose2rot=False (third-party, restated).
    betas [N,10], R [N,24,3,3] -> verts [N,V,3], posed joints [N,24,3].
    Dense, un-fused, materialises T[N,V,4,4] like smplx does."""
    N = betas.shape[0]
    dt = betas.dtype
    v_shaped = m.v_template[None] + torch.einsum("bl,mkl->bmk", betas, m.shapedirs)
    J = torch.einsum("bik,ji->bjk", v_shaped, m.J_regressor)
    eye = torch.eye(3, dtype=dt, device=betas.device)
    pose_feature = (R[:, 1:] - eye).reshape(N, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, m.posedirs).reshape(N, -1, 3)
    # rigid chain (smplx.lbs.batch_rigid_transform)
    rel = J.clone()
    rel[:, 1:] = rel[:, 1:] - J[:, m.parents[1:]]
    G = torch.zeros(N, 24, 4, 4, dtype=dt, device=betas.device)
    G[:, :, :3, :3] = R
    G[:, :, :3, 3] = rel
    G[:, :, 3, 3] = 1
    chain = [G[:, 0]]
    for i in range(1, 24):
        chain.append(torch.matmul(chain[int(m.parents[i])], G[:, i]))
    W = torch.stack(chain, dim=1)
    posed_J = W[:, :, :3, 3]
    J_h = F.pad(J, [0, 1])[..., None]                                  # [N,24,4,1] (w=0)
    A = W - F.pad(torch.matmul(W, J_h), [3, 0])
    T = torch.matmul(m.lbs_weights[None].expand(N, -1, -1), A.reshape(N, 24, 16)).reshape(N, -1, 4, 4)
    v_h = F.pad(v_posed, [0, 1], value=1.0)[..., None]
    verts = torch.matmul(T, v_h)[:, :, :3, 0]
    return verts, posed_J


def smpl_forward(m: SmplModel, betas, R=None, pose_aa=None):
    """lib/models/smpl.py:72-84 on top of smplx.SMPL.forward.
    Returns verts [N,V,3], joints49 [N,49,3], R [N,24,3,3]."""
    if R is None:  # pose2rot=True callers, e.g. lib/utils/eval_utils.py:168
        R = batch_rodrigues_smplx(pose_aa.reshape(-1, 3)).reshape(-1, 24, 3, 3)
    verts, posed_J = smpl_lbs(m, betas, R)
    joints45 = torch.cat([posed_J, verts[:, m.extra_vertex_ids]], dim=1)
    extra = torch.einsum("bik,ji->bjk", verts, m.J_regressor_extra)   # vertices2joints
    joints54 = torch.cat([joints45, extra], dim=1)
    return verts, joints54[:, JOINT_SOURCE_49], R


def projection(joints: torch.Tensor, cam: torch.Tensor) -> torch.Tensor:
    """lib/models/spin.py:307-351: perspective projection, f=5000, identity rotation,
    t = (cam1, cam2, 2*5000/(224*cam0+1e-9)), then /112."""
    t = torch.stack([cam[:, 1], cam[:, 2], 2 * 5000.0 / (224.0 * cam[:, 0] + 1e-9)], dim=-1)
    p = joints + t[:, None]
    p = p / p[:, :, 2:3]
    return (5000.0 * p[:, :, :2]) / 112.0


# --------------------------------------------------------------------------- network
def _t(sd, key):
    return torch.as_tensor(sd[key], dtype=torch.float32)


def build_gru(sd: dict, name: str, n_layers: int, hidden: int, bidir: bool) -> torch.nn.GRU:
    """torch.nn.GRU configured as lib/models/tepose.py:53-64, loaded from ``sd``."""
    gru = torch.nn.GRU(input_size=synth.INPUT_SIZE, hidden_size=hidden,
                       bidirectional=bidir, num_layers=n_layers)
    pref = f"encoder.{name}."
    gru.load_state_dict({k[len(pref):]: _t(sd, k) for k in sd if k.startswith(pref)})
    return gru.eval()


def encoder_forward(sd: dict, x: torch.Tensor, n_layers: int, hidden: int, is_train: bool = False,
                    grus=None) -> torch.Tensor:
    """lib/models/tepose.py:71-87.  x [B,T,2133] -> [B,2048] (eval) / [B,2,2048] (train)."""
    gf, gr = grus if grus is not None else (build_gru(sd, "gru_fwd", n_layers, hidden, False),
                                            build_gru(sd, "gru_rec", n_layers, hidden, True))
    xt = x.permute(1, 0, 2)
    y, _ = gf(xt)
    y_rec, _ = gr(torch.flip(xt, dims=[0]))
    a = F.linear(F.relu(y[-1]), _t(sd, "encoder.linear_fwd.weight"), _t(sd, "encoder.linear_fwd.bias"))
    b = F.linear(F.relu(y_rec[0]), _t(sd, "encoder.linear_rec.weight"), _t(sd, "encoder.linear_rec.bias"))
    if is_train:
        return torch.stack([a, b], dim=1)
    return (a + b) / 2


def ief_forward(sd: dict, feat: torch.Tensor, n_iter: int = 3, init=None):
    """lib/models/spin.py:240-261 (eval: dropout is identity; there is NO activation
    between fc1 and fc2 -- SURVEY.md F7).  feat [N,2048] -> pose6d, shape, cam."""
    N = feat.shape[0]
    if init is None:
        pose = _t(sd, "regressor.init_pose").expand(N, -1)
        shape = _t(sd, "regressor.init_shape").expand(N, -1)
        cam = _t(sd, "regressor.init_cam").expand(N, -1)
    else:
        pose, shape, cam = init
    W = lambda n: (_t(sd, f"regressor.{n}.weight"), _t(sd, f"regressor.{n}.bias"))
    for _ in range(n_iter):
        u = F.linear(F.linear(torch.cat([feat, pose, shape, cam], 1), *W("fc1")), *W("fc2"))
        pose = F.linear(u, *W("decpose")) + pose
        shape = F.linear(u, *W("decshape")) + shape
        cam = F.linear(u, *W("deccam")) + cam
    return pose, shape, cam


def regressor_forward(sd: dict, m: SmplModel, feat: torch.Tensor, J_regressor=None,
                      is_train: bool = False, n_iter: int = 3) -> dict:
    """lib/models/spin.py:240-291."""
    N = feat.shape[0]
    pose, shape, cam = ief_forward(sd, feat, n_iter)
    R = rot6d_to_rotmat(pose).reshape(N, 24, 3, 3)
    verts, joints, _ = smpl_forward(m, shape, R=R)
    if (not is_train) and J_regressor is not None:                     # spin.py:275-278
        joints = torch.matmul(J_regressor[None].expand(N, -1, -1), verts)[:, H36M_TO_J14]
    kp2d = projection(joints, cam)
    aa = rotmat_to_angle_axis(R.reshape(-1, 3, 3)).reshape(N, 72)
    return {"theta": torch.cat([cam, aa, shape], dim=1), "verts": verts,
            "kp_2d": kp2d, "kp_3d": joints, "rotmat": R}


def tepose_forward(sd: dict, m: SmplModel, x: torch.Tensor, n_layers: int, hidden: int,
                   is_train: bool = False, J_regressor=None, grus=None) -> dict:
    """lib/models/tepose.py:121-147 (the list-of-one-dict is flattened to the dict)."""
    B = x.shape[0]
    with torch.no_grad():
        feat = encoder_forward(sd, x, n_layers, hidden, is_train, grus)
        out = regressor_forward(sd, m, feat.reshape(-1, feat.shape[-1]), J_regressor, is_train)
    lead = (B, 2) if is_train else (B,)
    return {
        "theta": out["theta"].reshape(*lead, -1),
        "verts": out["verts"].reshape(*lead, -1, 3),
        "kp_2d": out["kp_2d"].reshape(*lead, -1, 2),
        "kp_3d": out["kp_3d"].reshape(*lead, -1, 3),
        "rotmat": out["rotmat"].reshape(*lead, -1, 3, 3),
    }


# --------------------------------------------------------------------------- VIBE bootstrap (evaluate.py:89-99,234)
def vibe_encoder_forward(sd: dict, x: torch.Tensor, n_layers: int, hidden: int, add_linear: bool = False,
                         bidirectional: bool = False, use_residual: bool = True) -> torch.Tensor:
    """lib/models/vibe.py:52-65.  x [B,T,2048] -> [B,T,F]."""
    gru = torch.nn.GRU(input_size=2048, hidden_size=hidden, bidirectional=bidirectional, num_layers=n_layers)
    pref = "encoder.gru."
    gru.load_state_dict({k[len(pref):]: _t(sd, k) for k in sd if k.startswith(pref)})
    gru.eval()
    n, t, f = x.shape
    xt = x.permute(1, 0, 2)
    y, _ = gru(xt)
    if bidirectional or add_linear:
        y = F.linear(F.relu(y).reshape(-1, y.size(-1)), _t(sd, "encoder.linear.weight"), _t(sd, "encoder.linear.bias"))
        y = y.reshape(t, n, f)
    if use_residual and y.shape[-1] == 2048:
        y = y + xt
    return y.permute(1, 0, 2)


def vibe_forward(sd: dict, m: SmplModel, x: torch.Tensor, n_layers: int, hidden: int, add_linear: bool = False,
                 bidirectional: bool = False, use_residual: bool = True, J_regressor=None) -> dict:
    """lib/models/vibe.py:104-119 (list-of-one-dict flattened)."""
    B, T = x.shape[:2]
    with torch.no_grad():
        feat = vibe_encoder_forward(sd, x, n_layers, hidden, add_linear, bidirectional, use_residual)
        out = regressor_forward(sd, m, feat.reshape(-1, feat.shape[-1]), J_regressor, False)
    return {
        "theta": out["theta"].reshape(B, T, -1),
        "verts": out["verts"].reshape(B, T, -1, 3),
        "kp_2d": out["kp_2d"].reshape(B, T, -1, 2),
        "kp_3d": out["kp_3d"].reshape(B, T, -1, 3),
        "rotmat": out["rotmat"].reshape(B, T, -1, 3, 3),
    }


def windowed_stream(sd_vibe: dict, vibe_arch: dict, sd: dict, m: SmplModel, features: torch.Tensor, seqlen: int,
                    n_layers: int, hidden: int, J_regressor=None, theta_input=None) -> dict:
    """The live loop of evaluate.py:229-269 (theta ring seeded from VIBE as demo.py:237 does, or from
    ``theta_input`` [T-1,85] as evaluate.py:219 does).  features [B,N,2048] -> per-frame outputs [B,N,...]."""
    B, N, T = features.shape[0], features.shape[1], seqlen
    first = vibe_forward(sd_vibe, m, features[:, :T], J_regressor=J_regressor, **vibe_arch)
    ring = (first["theta"][:, :T - 1].clone() if theta_input is None
            else torch.as_tensor(theta_input, dtype=torch.float32).reshape(-1, T - 1, 85).expand(B, -1, -1).clone())
    frames = {k: [v[:, t] for t in range(T - 1)] for k, v in first.items()}
    grus = (build_gru(sd, "gru_fwd", n_layers, hidden, False), build_gru(sd, "gru_rec", n_layers, hidden, True))
    for i in range(N - T + 1):
        x = torch.zeros(B, T, 2048 + 85)
        x[:, :, :2048] = features[:, i:i + T]
        x[:, :T - 1, 2048:] = ring
        out = tepose_forward(sd, m, x, n_layers, hidden, False, J_regressor, grus)
        for k in frames:
            frames[k].append(out[k])
        ring[:, :T - 2] = ring[:, 1:T - 1].clone()
        ring[:, T - 2] = out["theta"]
    return {k: torch.stack(v, dim=1) for k, v in frames.items()}


# --------------------------------------------------------------------------- carried state
def encoder_causal_states(sd: dict, x: torch.Tensor, hidden: int, h0=None):
    """Live-stream oracle (SURVEY.md F3/F4, L=1 only): the encoder restated as three
    causal pieces driven by explicit hidden state -- torch.nn.GRU stepped with h0.
    Returns (hF, hB, hS) after consuming the frames of ``x`` [B,T,2133] in order."""
    B = x.shape[0]
    H = hidden
    xt = x.permute(1, 0, 2)

    def run(prefix, sfx, h_init, frames):
        g = torch.nn.GRU(synth.INPUT_SIZE, H)
        g.load_state_dict({
            "weight_ih_l0": _t(sd, f"encoder.{prefix}.weight_ih_l0{sfx}"),
            "weight_hh_l0": _t(sd, f"encoder.{prefix}.weight_hh_l0{sfx}"),
            "bias_ih_l0": _t(sd, f"encoder.{prefix}.bias_ih_l0{sfx}"),
            "bias_hh_l0": _t(sd, f"encoder.{prefix}.bias_hh_l0{sfx}"),
        })
        _, hn = g(frames, h_init)
        return hn[0]

    z = torch.zeros(1, B, H)
    hF0, hB0 = (z, z) if h0 is None else (h0[0][None], h0[1][None])
    with torch.no_grad():
        hF = run("gru_fwd", "", hF0, xt)
        hB = run("gru_rec", "_reverse", hB0, xt)
        hS = run("gru_rec", "", z, xt[-1:])
    return hF, hB, hS


def encoder_from_states(sd, hF, hB, hS):
    """linear heads of lib/models/tepose.py:79-83 applied to the causal states."""
    a = F.linear(F.relu(hF), _t(sd, "encoder.linear_fwd.weight"), _t(sd, "encoder.linear_fwd.bias"))
    b = F.linear(F.relu(torch.cat([hS, hB], dim=1)), _t(sd, "encoder.linear_rec.weight"),
                 _t(sd, "encoder.linear_rec.bias"))
    return (a + b) / 2


# --------------------------------------------------------------------------- evaluation metrics (lib/utils/eval_utils.py)
def procrustes_align(S1: torch.Tensor, S2: torch.Tensor) -> torch.Tensor:
    """batch_compute_similarity_transform_torch, lib/utils/eval_utils.py:287-337.  S1, S2 [N,J,3] -> S1_hat [N,J,3]
    (computed in the dtype of the inputs; pass float64 for a tight reference)."""
    A, B = S1.permute(0, 2, 1), S2.permute(0, 2, 1)
    mu1, mu2 = A.mean(dim=-1, keepdim=True), B.mean(dim=-1, keepdim=True)
    X1, X2 = A - mu1, B - mu2
    var1 = (X1 ** 2).sum(dim=(1, 2))
    K = X1.bmm(X2.permute(0, 2, 1))
    U, s, Vh = torch.linalg.svd(K)
    V = Vh.transpose(1, 2)
    Z = torch.eye(3, dtype=A.dtype).repeat(A.shape[0], 1, 1)
    Z[:, -1, -1] *= torch.sign(torch.det(U.bmm(V.permute(0, 2, 1))))
    R = V.bmm(Z.bmm(U.permute(0, 2, 1)))
    scale = torch.diagonal(R.bmm(K), dim1=1, dim2=2).sum(-1) / var1
    t = mu2 - scale[:, None, None] * R.bmm(mu1)
    return (scale[:, None, None] * R.bmm(A) + t).permute(0, 2, 1)


def align_pelvis(j: torch.Tensor, pelvis=(2, 3)) -> torch.Tensor:
    """evaluate.py:420-428: hip midpoint (2, 3), a single joint (int) or none."""
    if pelvis is None:
        return j
    if isinstance(pelvis, int):
        return j - j[..., [pelvis], :]
    return j - (j[..., [pelvis[0]], :] + j[..., [pelvis[1]], :]) / 2.0


def pose_metrics(pred: torch.Tensor, target: torch.Tensor, pelvis=(2, 3)) -> dict:
    """evaluate.py:420-443 for one sequence [N,J,3] (metres)."""
    P, G = align_pelvis(pred, pelvis), align_pelvis(target, pelvis)
    mpjpe = torch.sqrt(((P - G) ** 2).sum(-1)).mean(-1)
    hat = procrustes_align(P, G)
    pa = torch.sqrt(((hat - G) ** 2).sum(-1)).mean(-1)
    accel = torch.zeros(P.shape[0], dtype=P.dtype)
    if P.shape[0] >= 3:
        accel[1:-1] = accel_error(P, G)
    return {"mpjpe": mpjpe, "mpjpe_pa": pa, "accel_err": accel, "aligned": hat}


def accel_error(pred: torch.Tensor, target=None) -> torch.Tensor:
    """compute_error_accel_eval (eval_utils.py:110-138, vis=None) on [...,N,J,3] -> [...,N-2]; target None: the norm of
    pred's own acceleration (compute_accel, :60-63)."""
    a = pred[..., :-2, :, :] - 2 * pred[..., 1:-1, :, :] + pred[..., 2:, :, :]
    if target is not None:
        a = a - (target[..., :-2, :, :] - 2 * target[..., 1:-1, :, :] + target[..., 2:, :, :])
    return torch.linalg.norm(a, dim=-1).mean(-1)


def accel_summary(normed: torch.Tensor, vidlen_each: torch.Tensor, seqlen: int, tail: int, extra: int) -> torch.Tensor:
    """The reductions of compute_accel (tail 2, extra 1) / compute_error_accel (tail 4, extra 3), eval_utils.py:70-76,104-108."""
    out = torch.zeros((), dtype=normed.dtype)
    for i in range(normed.shape[0]):
        out = out + normed[i, seqlen - 1:int(vidlen_each[i]) - tail].sum()
    return out / (vidlen_each.sum() - vidlen_each.shape[0] * (seqlen + extra) + 1e-8)


def vertex_error(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """compute_error_verts, eval_utils.py:173-175: [N,V,3] x 2 -> [N]."""
    return torch.sqrt(((a - b) ** 2).sum(dim=2)).mean(dim=1)

"""Stage-by-stage statement of the SMPL / projection BACKWARD the CUDA kernels of csrc/train.cu implement (test
infrastructure; CPU, torch, no autograd inside).  tests/test_train_oracle.py checks it against torch.autograd through
oracle/torch_ref.py:smpl_forward + projection, so the kernel's decomposition is verified before it is written in CUDA.

Forward being differentiated (third-party smplx lbs restated in oracle/torch_ref.py:148-176, wrapper
lib/models/smpl.py:72-84, projection lib/models/spin.py:307-351):
    J = J_reg.(v_t + S b);  v_posed = v_t + S b + P^T vec(R[1:] - I);  chain W_i = W_parent(i) [R_i | J_i - J_parent];
    A_i = [RW_i | t_i - RW_i J_i];  vert_v = sum_k w_vk (A_k.R v_posed_v + A_k.t);  joints49 = select(posed J | vertex
    picks | J_extra . verts);  kp_2d = (f/112) (X + c1, Y + c2) / (Z + 2 f / (224 c0 + 1e-9)).
"""
from __future__ import annotations

import torch

F_LEN = 5000.0


def projection_backward(joints, cam, g_kp2d):
    """-> g_joints [N,J,3], g_cam [N,3]"""
    s = F_LEN / 112.0
    den = 224.0 * cam[:, 0] + 1e-9
    t = torch.stack([cam[:, 1], cam[:, 2], 2 * F_LEN / den], dim=-1)
    p = joints + t[:, None]
    gx = s * g_kp2d[..., 0] / p[..., 2]
    gy = s * g_kp2d[..., 1] / p[..., 2]
    gz = -s * (g_kp2d[..., 0] * p[..., 0] + g_kp2d[..., 1] * p[..., 1]) / (p[..., 2] ** 2)
    g_p = torch.stack([gx, gy, gz], dim=-1)
    g_t = g_p.sum(dim=1)
    g_cam = torch.stack([g_t[:, 2] * (-2 * F_LEN * 224.0 / den ** 2), g_t[:, 0], g_t[:, 1]], dim=-1)
    return g_p, g_cam


def chain_forward(m, betas, R):
    """J [N,24,3], RW [N,24,3,3], t [N,24,3], rel [N,24,3] of the kinematic chain."""
    v_shaped = m.v_template[None] + torch.einsum("bl,mkl->bmk", betas, m.shapedirs)
    J = torch.einsum("bik,ji->bjk", v_shaped, m.J_regressor)
    par = m.parents
    rel = J.clone()
    rel[:, 1:] = J[:, 1:] - J[:, par[1:]]
    RW, t = [R[:, 0]], [rel[:, 0]]
    for i in range(1, 24):
        p = int(par[i])
        RW.append(RW[p] @ R[:, i])
        t.append((RW[p] @ rel[:, i, :, None])[..., 0] + t[p])
    return J, torch.stack(RW, 1), torch.stack(t, 1), rel, v_shaped


def smpl_backward(m, joint_src, betas, R, cam, g_verts_in, g_kp3d, g_kp2d, g_R_extra=None):
    """joint_src: list of 49 indices into [24 posed | 21 vertex picks | 9 regressed] (oracle/torch_ref.JOINT_SOURCE_49).
    Returns g_R [N,24,3,3], g_betas [N,10], g_cam [N,3]."""
    N, V = betas.shape[0], m.v_template.shape[0]
    # ---- recompute what the forward produced
    J, RW, t, rel, v_shaped = chain_forward(m, betas, R)
    eye = torch.eye(3, dtype=R.dtype)
    pf = (R[:, 1:] - eye).reshape(N, -1)
    v_posed = v_shaped + (pf @ m.posedirs).reshape(N, V, 3)
    A_R = RW
    A_t = t - (RW @ J[..., None])[..., 0]
    Wk = m.lbs_weights                                              # [V,24] (dense here; the kernel keeps the top-k)
    T_R = torch.einsum("vj,njab->nvab", Wk, A_R)
    T_t = torch.einsum("vj,nja->nva", Wk, A_t)
    verts = (T_R @ v_posed[..., None])[..., 0] + T_t
    joints45 = torch.cat([t, verts[:, m.extra_vertex_ids]], dim=1)
    extra = torch.einsum("bik,ji->bjk", verts, m.J_regressor_extra)
    joints49 = torch.cat([joints45, extra], dim=1)[:, joint_src]
    # ---- stage 1: projection + joint scatter
    g_p, g_cam = projection_backward(joints49, cam, g_kp2d)
    g_j49 = g_kp3d + g_p
    g_posedJ = torch.zeros(N, 24, 3, dtype=R.dtype)
    g_verts = torch.zeros(N, V, 3, dtype=R.dtype) if g_verts_in is None else g_verts_in.clone()
    g_extra = torch.zeros(N, 9, 3, dtype=R.dtype)
    for k, s in enumerate(joint_src):
        if s < 24:
            g_posedJ[:, s] += g_j49[:, k]
        elif s < 45:
            g_verts[:, int(m.extra_vertex_ids[s - 24])] += g_j49[:, k]
        else:
            g_extra[:, s - 45] += g_j49[:, k]
    # ---- stage 2: regressed joints -> vertices
    g_verts = g_verts + torch.einsum("jv,njc->nvc", m.J_regressor_extra, g_extra)
    # ---- stage 3: skinning
    g_vposed = (T_R.transpose(-1, -2) @ g_verts[..., None])[..., 0]
    g_A_R = torch.einsum("vj,nva,nvb->njab", Wk, g_verts, v_posed)
    g_A_t = torch.einsum("vj,nva->nja", Wk, g_verts)
    # ---- stage 4: blend shapes (one GEMM against [posedirs ; shapedirs])
    g_pf = g_vposed.reshape(N, -1) @ m.posedirs.t()                                  # [N,207]
    g_betas = torch.einsum("nvc,vcl->nl", g_vposed, m.shapedirs)                     # [N,10]
    # ---- stage 5: kinematic chain
    gRW = g_A_R - g_A_t[..., :, None] * J[..., None, :]
    gt = g_A_t + g_posedJ
    gJ = -(RW.transpose(-1, -2) @ g_A_t[..., None])[..., 0]
    gRW, gt, gJ = [x.clone() for x in (gRW, gt, gJ)]
    g_R = torch.zeros_like(R)
    par = m.parents
    for i in range(23, 0, -1):
        p = int(par[i])
        gRW[:, p] += gRW[:, i] @ R[:, i].transpose(-1, -2) + gt[:, i, :, None] * rel[:, i, None, :]
        g_R[:, i] = RW[:, p].transpose(-1, -2) @ gRW[:, i]
        grel = (RW[:, p].transpose(-1, -2) @ gt[:, i, :, None])[..., 0]
        gt[:, p] += gt[:, i]
        gJ[:, i] += grel
        gJ[:, p] -= grel
    g_R[:, 0] = gRW[:, 0]
    gJ[:, 0] += gt[:, 0]
    g_R[:, 1:] += g_pf.reshape(N, 23, 3, 3)
    if g_R_extra is not None:
        g_R = g_R + g_R_extra
    # J = J_reg (v_t + S b): gradient reaches the betas through the folded table J_reg.S [24,3,10]
    j_shapedirs = torch.einsum("jv,vcl->jcl", m.J_regressor, m.shapedirs)
    g_betas = g_betas + torch.einsum("njc,jcl->nl", gJ, j_shapedirs)
    return g_R, g_betas, g_cam

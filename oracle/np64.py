"""float64 NumPy restatement of the arithmetic on the hot path -- TEST INFRASTRUCTURE.

Independent of ``oracle/torch_ref.py`` (loops instead of batched matmuls, float64
instead of float32) so the two can be checked against each other; this is the only
guard on the un-pinned smplx boundary.  Small sizes only.
Citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRU cell: gate rows ordered r|z|n;
    n = tanh(W_in x + b_in + r*(W_hn h + b_hn)); h' = (1-z)*n + z*h."""
    H = h.shape[-1]
    gi = x @ w_ih.T + b_ih
    gh = h @ w_hh.T + b_hh
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    r = sig(gi[:, :H] + gh[:, :H])
    z = sig(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1.0 - z) * n + z * h


def gru_layer(xs, w_ih, w_hh, b_ih, b_hh, reverse=False, h0=None):
    """One direction of one nn.GRU layer over xs [T,B,F]; returns ys [T,B,H]."""
    T, B, _ = xs.shape
    H = w_hh.shape[1]
    h = np.zeros((B, H)) if h0 is None else h0
    ys = np.zeros((T, B, H))
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        h = gru_cell(xs[t], h, w_ih, w_hh, b_ih, b_hh)
        ys[t] = h
    return ys


def encoder(sd, x, n_layers, hidden, is_train=False):
    """lib/models/tepose.py:71-87 with both GRUs written out as cell loops."""
    f = lambda k: np.asarray(sd[k], dtype=np.float64)
    xs = np.transpose(np.asarray(x, np.float64), (1, 0, 2))

    def stack(name, bidir, seq):
        cur = seq
        for l in range(n_layers):
            p = f"encoder.{name}."
            outs = [gru_layer(cur, f(p + f"weight_ih_l{l}"), f(p + f"weight_hh_l{l}"),
                              f(p + f"bias_ih_l{l}"), f(p + f"bias_hh_l{l}"))]
            if bidir:
                outs.append(gru_layer(cur, f(p + f"weight_ih_l{l}_reverse"), f(p + f"weight_hh_l{l}_reverse"),
                                      f(p + f"bias_ih_l{l}_reverse"), f(p + f"bias_hh_l{l}_reverse"),
                                      reverse=True))
            cur = np.concatenate(outs, axis=-1)
        return cur

    y = stack("gru_fwd", False, xs)
    y_rec = stack("gru_rec", True, xs[::-1])
    relu = lambda v: np.maximum(v, 0.0)
    a = relu(y[-1]) @ f("encoder.linear_fwd.weight").T + f("encoder.linear_fwd.bias")
    b = relu(y_rec[0]) @ f("encoder.linear_rec.weight").T + f("encoder.linear_rec.bias")
    return np.stack([a, b], 1) if is_train else (a + b) / 2


def ief(sd, feat, n_iter=3):
    """lib/models/spin.py:250-261."""
    f = lambda k: np.asarray(sd[k], dtype=np.float64)
    N = feat.shape[0]
    pose = np.repeat(f("regressor.init_pose"), N, 0)
    shape = np.repeat(f("regressor.init_shape"), N, 0)
    cam = np.repeat(f("regressor.init_cam"), N, 0)
    lin = lambda n, v: v @ f(f"regressor.{n}.weight").T + f(f"regressor.{n}.bias")
    for _ in range(n_iter):
        u = lin("fc2", lin("fc1", np.concatenate([feat, pose, shape, cam], 1)))
        pose, shape, cam = pose + lin("decpose", u), shape + lin("decshape", u), cam + lin("deccam", u)
    return pose, shape, cam


def rot6d_to_rotmat(x):
    """lib/utils/geometry.py:330-343."""
    p = np.asarray(x, np.float64).reshape(-1, 3, 2)
    out = np.zeros((p.shape[0], 3, 3))
    for i, blk in enumerate(p):
        a1, a2 = blk[:, 0], blk[:, 1]
        b1 = a1 / max(np.linalg.norm(a1), 1e-6)
        u = a2 - (b1 @ a2) * b1
        b2 = u / max(np.linalg.norm(u), 1e-6)
        out[i] = np.stack([b1, b2, np.cross(b1, b2)], axis=1)
    return out


def rodrigues(aa):
    """smplx.lbs.batch_rodrigues (restated): one [3] vector -> [3,3]."""
    aa = np.asarray(aa, np.float64)
    ang = np.linalg.norm(aa + 1e-8)
    k = aa / ang
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def rotmat_to_angle_axis(R):
    """lib/utils/geometry.py:68-233 for a single [3,3] matrix (branch-for-branch)."""
    M = np.asarray(R, np.float64).T
    if M[2, 2] < 1e-6:
        if M[0, 0] > M[1, 1]:
            t = 1 + M[0, 0] - M[1, 1] - M[2, 2]
            q = [M[1, 2] - M[2, 1], t, M[0, 1] + M[1, 0], M[2, 0] + M[0, 2]]
        else:
            t = 1 - M[0, 0] + M[1, 1] - M[2, 2]
            q = [M[2, 0] - M[0, 2], M[0, 1] + M[1, 0], t, M[1, 2] + M[2, 1]]
    else:
        if M[0, 0] < -M[1, 1]:
            t = 1 - M[0, 0] - M[1, 1] + M[2, 2]
            q = [M[0, 1] - M[1, 0], M[2, 0] + M[0, 2], M[1, 2] + M[2, 1], t]
        else:
            t = 1 + M[0, 0] + M[1, 1] + M[2, 2]
            q = [t, M[1, 2] - M[2, 1], M[2, 0] - M[0, 2], M[0, 1] - M[1, 0]]
    with np.errstate(all="ignore"):
        q = np.array(q) / np.sqrt(t) * 0.5
        w, v = q[0], q[1:]
        s2 = float(v @ v)
        s = np.sqrt(s2)
        two_theta = 2.0 * (np.arctan2(-s, -w) if w < 0 else np.arctan2(s, w))
        k = two_theta / s if s2 > 0 else 2.0
        aa = v * k
    aa[np.isnan(aa)] = 0.0
    return aa


def smpl(model, extra, betas, R, joint_source, extra_vertex_ids):
    """smplx SMPL forward + lib/models/smpl.py:72-84 wrapper for ONE body, as loops.
    model: pkl-style dict; betas [10]; R [24,3,3].  -> verts [V,3], joints49 [49,3]."""
    f = lambda a: np.asarray(a, np.float64)
    vt, sdirs, pdirs = f(model["v_template"]), f(model["shapedirs"]), f(model["posedirs"])
    Jr, W = f(model["J_regressor"]), f(model["weights"])
    parents = np.asarray(model["kintree_table"])[0].astype(np.int64).copy()
    parents[0] = -1
    betas, R = f(betas), f(R)
    V = vt.shape[0]
    v_shaped = vt + sdirs @ betas                                    # [V,3]
    J = Jr @ v_shaped                                                # [24,3]
    pf = (R[1:] - np.eye(3)).reshape(-1)                             # [207]
    v_posed = v_shaped + pdirs.reshape(V, 3, -1) @ pf
    world = [None] * 24
    for i in range(24):
        G = np.eye(4)
        G[:3, :3] = R[i]
        G[:3, 3] = J[i] - (J[parents[i]] if i > 0 else 0.0)
        world[i] = G if i == 0 else world[parents[i]] @ G
    A = []
    for i in range(24):
        Ai = world[i].copy()
        Ai[:3, 3] = world[i][:3, 3] - world[i][:3, :3] @ J[i]
        A.append(Ai)
    A = np.stack(A)                                                  # [24,4,4]
    T = np.tensordot(W, A, axes=(1, 0))                              # [V,4,4]
    verts = np.einsum("vij,vj->vi", T[:, :3, :3], v_posed) + T[:, :3, 3]
    posed_J = np.stack([w[:3, 3] for w in world])
    j54 = np.concatenate([posed_J, verts[extra_vertex_ids], f(extra["J_regressor_extra"]) @ verts])
    return verts, j54[joint_source], posed_J


def projection(joints, cam):
    """lib/models/spin.py:307-351 for one body: joints [J,3], cam [3] -> [J,2]."""
    t = np.array([cam[1], cam[2], 2 * 5000.0 / (224.0 * cam[0] + 1e-9)])
    p = np.asarray(joints, np.float64) + t
    return 5000.0 * (p[:, :2] / p[:, 2:3]) / 112.0

"""SMPL body model with the reference's wrapper API (lib/models/smpl.py) on the fused
sm_100a kernels of csrc/smpl.cu.

Mirrors: constants JOINT_MAP / JOINT_NAMES / JOINT_IDS / H36M_TO_J17 / H36M_TO_J14 /
SMPL_MODEL_DIR / SMPL_MEAN_PARAMS / JOINT_REGRESSOR_TRAIN_EXTRA (lib/models/smpl.py:14-58),
class SMPL (lib/models/smpl.py:61-84, which subclasses the third-party smplx.SMPL) and
get_smpl_faces (:87-89).  Buffer / Parameter names follow smplx so released TePose
checkpoints load with strict=True.
"""
from __future__ import annotations

import os
import os.path as osp
import pickle
from collections import namedtuple

import numpy as np
import ctypes as C

import torch
import torch.nn as nn

from . import _native as nv

BASE_DATA_DIR = "data/base_data"          # lib/core/config.py:31 (cwd-relative, like the reference)

JOINT_NAMES = [
    'OP Nose', 'OP Neck', 'OP RShoulder', 'OP RElbow', 'OP RWrist', 'OP LShoulder', 'OP LElbow', 'OP LWrist',
    'OP MidHip', 'OP RHip', 'OP RKnee', 'OP RAnkle', 'OP LHip', 'OP LKnee', 'OP LAnkle', 'OP REye', 'OP LEye',
    'OP REar', 'OP LEar', 'OP LBigToe', 'OP LSmallToe', 'OP LHeel', 'OP RBigToe', 'OP RSmallToe', 'OP RHeel',
    'Right Ankle', 'Right Knee', 'Right Hip', 'Left Hip', 'Left Knee', 'Left Ankle', 'Right Wrist',
    'Right Elbow', 'Right Shoulder', 'Left Shoulder', 'Left Elbow', 'Left Wrist', 'Neck (LSP)',
    'Top of Head (LSP)', 'Pelvis (MPII)', 'Thorax (MPII)', 'Spine (H36M)', 'Jaw (H36M)', 'Head (H36M)',
    'Nose', 'Left Eye', 'Right Eye', 'Left Ear', 'Right Ear',
]
_SOURCES = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
            8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20, 47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27]
JOINT_MAP = dict(zip(JOINT_NAMES, _SOURCES))   # name -> index into [24 posed | 21 picked | 9 regressed]
JOINT_IDS = {name: i for i, name in enumerate(JOINT_NAMES)}
JOINT_REGRESSOR_TRAIN_EXTRA = osp.join(BASE_DATA_DIR, 'J_regressor_extra.npy')
SMPL_MEAN_PARAMS = osp.join(BASE_DATA_DIR, 'smpl_mean_params.npz')
SMPL_MODEL_DIR = BASE_DATA_DIR
H36M_TO_J17 = [6, 5, 4, 1, 2, 3, 16, 15, 14, 11, 12, 13, 8, 10, 0, 7, 9]
H36M_TO_J14 = H36M_TO_J17[:14]

# smplx VertexJointSelector for SMPL: face (5), feet (6), finger tips (10)
EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                    2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]

NUM_BODY_JOINTS = 23
VERT_TILE = 128

SMPLOutput = namedtuple("SMPLOutput", ["vertices", "joints", "full_pose", "betas", "global_orient", "body_pose"])
SMPLOutput.__new__.__defaults__ = (None,) * 6


class _ChumpyShim:
    """Stands in for chumpy.Ch objects inside the licensed SMPL pickles (chumpy is not a
    dependency): keeps whatever state the pickle carries and exposes it as an array."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"x": state})

    def __array__(self, dtype=None):
        a = np.asarray(self.__dict__.get("x", self.__dict__.get("r")))
        return a.astype(dtype) if dtype is not None else a


class _SmplUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("chumpy"):
            return _ChumpyShim
        return super().find_class(module, name)


def load_smpl_file(path: str) -> dict:
    """Reads an SMPL model file (.pkl as distributed, or .npz) into numpy arrays."""
    if path.endswith(".npz"):
        return dict(np.load(path))
    with open(path, "rb") as fh:
        data = _SmplUnpickler(fh, encoding="latin1").load()
    out = {}
    for k, v in data.items():
        if hasattr(v, "toarray"):          # scipy.sparse J_regressor
            v = v.toarray()
        try:
            out[k] = np.asarray(v)
        except Exception:                  # non-array metadata
            out[k] = v
    return out


def joint_source_codes(sources) -> list:
    """Index into [24 posed | 21 picked vertices | regressed rows] -> TP_JSRC_* code."""
    codes = []
    for s in sources:
        if s < 24:
            codes.append(s)
        elif s < 45:
            codes.append(1000 + EXTRA_VERTEX_IDS[s - 24])
        else:
            codes.append(100 + (s - 45))
    return codes


def pack_regressor(J: torch.Tensor, vp: int, device) -> torch.Tensor:
    """[R, n_verts] joint regressor -> dense fp32 [R, vp] zero padded (tp_smpl_forward `jreg`)."""
    J = J.detach().to(device=device, dtype=torch.float32)
    out = torch.zeros(J.shape[0], vp, device=device, dtype=torch.float32)
    out[:, :J.shape[1]] = J
    return out


class PackedSmpl:
    """Device-resident tables behind tp_smpl_model (layouts documented in include/tepose_b200.h)."""

    def __init__(self, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, device):
        f64 = lambda t: t.detach().to(device=device, dtype=torch.float64)
        V = v_template.shape[0]
        vp = (V + VERT_TILE - 1) // VERT_TILE * VERT_TILE
        self.n_verts, self.vp = V, vp
        blend = torch.zeros(218, 3, vp, device=device, dtype=torch.float32)
        blend[:207, :, :V] = posedirs.detach().to(device).float().reshape(207, V, 3).permute(0, 2, 1)
        blend[207:217, :, :V] = shapedirs.detach().to(device).float()[:, :, :10].permute(2, 1, 0)
        blend[217, :, :V] = v_template.detach().to(device).float().t()
        self.blend = blend.contiguous()
        # J = J_regressor.(v_template + shapedirs.beta): fold the regressor in (float64, then round)
        Jr = f64(J_regressor)
        self.j_template = (Jr @ f64(v_template)).float().contiguous()                                   # [24,3]
        self.j_shapedirs = torch.einsum("jv,vcl->jcl", Jr, f64(shapedirs)[:, :, :10]).float().contiguous()  # [24,3,10]
        par = parents.detach().to("cpu", torch.int64).clone()
        par[0] = -1
        if not bool((par[1:] < torch.arange(1, par.numel())).all()):
            raise ValueError("SMPL kinematic tree must satisfy parents[i] < i")
        self.parents = par.to(device=device, dtype=torch.int32).contiguous()
        W = lbs_weights.detach().to(device).float()
        ks = int((W != 0).sum(dim=1).max().item())
        ks = max(1, min(ks, W.shape[1]))
        vals, idx = torch.topk(W.abs(), ks, dim=1)
        wsel = torch.gather(W, 1, idx)
        self.ks = ks
        self.skin_idx = torch.zeros(vp, ks, device=device, dtype=torch.int32)
        self.skin_w = torch.zeros(vp, ks, device=device, dtype=torch.float32)
        self.skin_idx[:V] = idx.to(torch.int32)
        self.skin_w[:V] = wsel
        self.device = device
        # tensor-core blend tables (bf16 mode): [3*vp, 256] rows ((v//16)*3 + c)*16 + v%16
        self.blend_tc = self.template_pad = self.blend_km = self.blend_um = self.skin_um = self.w_dense = None
        self._folds = {}
        if ks <= 4:
            S = shapedirs.detach().to(device).float()[:, :, :10]                       # [V,3,10]
            S_hi = S.to(torch.bfloat16).float()
            cols = torch.zeros(vp, 3, 256, device=device, dtype=torch.float32)
            cols[:V, :, :207] = posedirs.detach().to(device).float().reshape(207, V, 3).permute(1, 2, 0)
            cols[:V, :, 207:217] = S_hi
            cols[:V, :, 217:227] = S_hi
            cols[:V, :, 227:237] = S - S_hi
            rows = cols.reshape(vp // 16, 16, 3, 256).permute(0, 2, 1, 3).reshape(vp * 3, 256).contiguous()
            self.blend_tc = torch.empty(nv.lib().tp_pack_mma_a_bytes(vp * 3, 256), dtype=torch.uint8, device=device)
            nv.check(nv.lib().tp_pack_mma_a_bf16(nv.ptr(rows), 256, vp * 3, 256, nv.ptr(self.blend_tc), nv.stream()),
                     "tp_pack_mma_a_bf16")
            self.blend_km = cols.reshape(vp * 3, 256).to(torch.bfloat16).contiguous()        # row v*3 + c (tcgen05 GEMM operand)
            self.template_pad = torch.zeros(vp, 3, device=device, dtype=torch.float32)
            self.template_pad[:V] = v_template.detach().to(device).float()
            # A-operand image of the fused tcgen05 blend + skinning kernel: [tile 128 v][plane][K block 64][row][8 chunks of 8 bf16],
            # chunk q of row r stored at position q ^ (r & 7) (the 128-byte swizzle the shared-memory staging is read back with)
            img = cols.to(torch.bfloat16).reshape(vp // 128, 128, 3, 4, 8, 8).permute(0, 2, 3, 1, 4, 5)     # t, c, kb, r, q, e
            sw = (torch.arange(8, device=device)[None, :] ^ (torch.arange(128, device=device) % 8)[:, None])   # [r, q'] -> source chunk
            self.blend_um = torch.gather(img, 4, sw[None, None, None, :, :, None].expand(vp // 128, 3, 4, 128, 8, 8)).contiguous()
            # A operand of the skinning MMA: per vertex row 64 bf16 = W_hi (24 joints) | 0 x 8 | W_lo (24) | 0 x 8, W = hi + lo
            Wd = torch.zeros(vp, 24, device=device, dtype=torch.float32)
            Wd[:V] = W
            self.w_dense = Wd
            w_hi = Wd.to(torch.bfloat16)
            w_lo = (Wd - w_hi.float()).to(torch.bfloat16)
            row = torch.zeros(vp, 64, device=device, dtype=torch.bfloat16)
            row[:, :24], row[:, 32:56] = w_hi, w_lo
            rimg = row.reshape(vp // 128, 128, 8, 8)
            self.skin_um = torch.gather(rimg, 2, sw[None, :, :, None].expand(vp // 128, 128, 8, 8)).contiguous()
        self.c_model = nv.SmplModel(nv.ptr(self.blend), nv.ptr(self.j_template), nv.ptr(self.j_shapedirs),
                                    nv.ptr(self.parents), nv.ptr(self.skin_idx), nv.ptr(self.skin_w),
                                    ks, V, vp, nv.ptr(self.blend_tc), nv.ptr(self.template_pad), nv.ptr(self.blend_km),
                                    nv.ptr(self.blend_um), nv.ptr(self.skin_um))


LARGE_BATCH = 1024          # bodies from which tp_smpl_forward takes the large-batch tcgen05 path (csrc/smpl.cu split_min_bodies)


class RegFold:
    """tp_smpl_regfold of one joint regressor: the regressor folded through the skinning weights and the (bf16) blend matrix in
    float64 at pack time, so that the large-batch path gets its regressed joints from one small GEMM over all bodies."""

    def __init__(self, packed: "PackedSmpl", jreg: torch.Tensor):
        R, vp = int(jreg.shape[0]), packed.vp
        G = (jreg.double()[:, None, :] * packed.w_dense.double().t()[None]).reshape(R * 24, vp)       # [(r, j), v]
        cols = packed.blend_km.double().reshape(vp, 3 * 256)                                            # [v, (c, k)] as the kernels read it
        M = (G @ cols).reshape(R * 24 * 3, 256)                                                         # row (r*24 + j)*3 + c
        self.nreg, self.nq_pad = R, (R * 72 + 15) // 16 * 16
        hi = M.float().to(torch.bfloat16)
        lo = (M - hi.double()).float().to(torch.bfloat16)
        self.m_km = torch.zeros(self.nq_pad, 512, device=jreg.device, dtype=torch.bfloat16)       # [M_hi | M_lo] against [coef | coef]
        self.m_km[:R * 72, :256], self.m_km[:R * 72, 256:] = hi, lo
        self.q_bias = torch.zeros(self.nq_pad, device=jreg.device, dtype=torch.float32)
        self.q_bias[:R * 72] = (G @ packed.template_pad.double()).reshape(-1).float()
        self.g0 = G.sum(dim=1).float().contiguous()
        self.c_fold = nv.SmplRegFold(nv.ptr(self.m_km), nv.ptr(self.q_bias), nv.ptr(self.g0), self.nreg, self.nq_pad)


def _fold_for(packed: "PackedSmpl", jreg):
    if jreg is None or packed.skin_um is None:
        return None
    key = (jreg.data_ptr(), jreg._version, tuple(jreg.shape))
    hit = packed._folds.get(key)
    if hit is None:
        hit = RegFold(packed, jreg)
        if len(packed._folds) > 4:
            packed._folds.clear()
        packed._folds[key] = hit
    return hit


def smpl_forward_native(packed: PackedSmpl, pose: torch.Tensor, ld_pose: int, pose_kind: int,
                        betas: torch.Tensor, ld_betas: int, cam, ld_cam: int, n: int,
                        jreg, joint_src: torch.Tensor, want_theta: bool, want_rotmat: bool = True,
                        blend_mode: int = 0):
    """One tp_smpl_forward call.  `pose`/`betas`/`cam` may be column views into a wider row
    (e.g. the IEF state [N,160]); pointers + row strides are passed as they are."""
    dev = packed.device
    nj = int(joint_src.numel())
    nreg = 0 if jreg is None else int(jreg.shape[0])
    # all outputs are views of ONE allocation (256-byte aligned pieces): a caller that ships every output to the host
    # (PipelinedTePose) does it with a single copy
    shapes = [("verts", (n, packed.n_verts, 3)), ("joints", (n, nj, 3))]
    if cam is not None:
        shapes.append(("kp2d", (n, nj, 2)))
    if want_rotmat:
        shapes.append(("rotmat", (n, 24, 3, 3)))
    if want_theta:
        shapes.append(("theta", (n, 85)))
    offs, total = {}, 0
    for name, shp in shapes:
        offs[name] = total
        total += (4 * int(torch.Size(shp).numel()) + 255) // 256 * 256
    flat = torch.empty(max(total, 256), device=dev, dtype=torch.uint8)
    piece = lambda name, shp: flat[offs[name]:offs[name] + 4 * int(torch.Size(shp).numel())].view(torch.float32).view(shp)
    outs = {name: piece(name, shp) for name, shp in shapes}
    verts, joints = outs["verts"], outs["joints"]
    kp2d, rotmat, theta = outs.get("kp2d"), outs.get("rotmat"), outs.get("theta")
    L = nv.lib()
    if packed.blend_tc is None:
        blend_mode = 0
    nbytes = L.tp_smpl_workspace_bytes(packed.c_model, n, nreg, blend_mode)
    ws = nv.workspace(nbytes, dev)
    P = lambda t: nv.vp(0) if t is None else nv.vp(t.data_ptr())
    fold = _fold_for(packed, jreg) if (blend_mode == 1 and n >= LARGE_BATCH) else None
    nv.check(L.tp_smpl_forward_ex(packed.c_model, n, P(pose), ld_pose, pose_kind, P(betas), ld_betas, P(cam), ld_cam,
                                  P(jreg), nreg, None if fold is None else C.byref(fold.c_fold), P(joint_src), nj, P(verts), P(joints),
                                  P(kp2d), P(rotmat), P(theta), blend_mode, P(ws), ws.numel(), nv.stream()), "tp_smpl_forward")
    return verts, joints, kp2d, rotmat, theta


class SMPL(nn.Module):
    """Drop-in for lib.models.smpl.SMPL (reference lib/models/smpl.py:61-84).

    SMPL(model_path, batch_size=1, create_transl=True, gender='neutral')
    forward(betas=, body_pose=, global_orient=, pose2rot=True, transl=None, **kw) -> SMPLOutput
    with .vertices [N,6890,3] and .joints [N,49,3] (24 posed | 21 picked | 9 regressed, then
    re-indexed by JOINT_MAP order).
    """

    NUM_JOINTS = 23
    NUM_BODY_JOINTS = 23

    def __init__(self, model_path, batch_size=1, create_transl=True, gender='neutral', dtype=torch.float32, **kwargs):
        super().__init__()
        if osp.isdir(model_path):
            cands = [osp.join(model_path, f"SMPL_{gender.upper()}.{ext}") for ext in ("pkl", "npz")]
            found = [c for c in cands if osp.isfile(c)]
            if not found:
                raise FileNotFoundError(f"no SMPL_{gender.upper()}.pkl/.npz under {model_path}")
            model_path = found[0]
        data = load_smpl_file(model_path)
        self.gender = gender
        self.batch_size = batch_size
        self.faces = np.asarray(data["f"]).astype(np.int64)
        t = lambda a: torch.tensor(np.asarray(a, dtype=np.float32))
        V = np.asarray(data["v_template"]).shape[0]
        self.register_buffer("faces_tensor", torch.tensor(self.faces, dtype=torch.long))
        self.register_buffer("v_template", t(data["v_template"]))
        self.register_buffer("shapedirs", t(np.asarray(data["shapedirs"])[:, :, :10]))
        self.register_buffer("J_regressor", t(data["J_regressor"]))
        self.register_buffer("posedirs", t(np.asarray(data["posedirs"]).reshape(V * 3, -1).T))   # [207, 3V]
        parents = np.asarray(data["kintree_table"])[0].astype(np.int64).copy()
        parents[0] = -1
        self.register_buffer("parents", torch.tensor(parents, dtype=torch.long))
        self.register_buffer("lbs_weights", t(data["weights"]))
        self.vertex_joint_selector = nn.Module()
        self.vertex_joint_selector.register_buffer("extra_joints_idxs", torch.tensor(EXTRA_VERTEX_IDS, dtype=torch.long))
        self.betas = nn.Parameter(torch.zeros(batch_size, 10, dtype=dtype))
        self.global_orient = nn.Parameter(torch.zeros(batch_size, 3, dtype=dtype))
        self.body_pose = nn.Parameter(torch.zeros(batch_size, 69, dtype=dtype))
        if create_transl:
            self.transl = nn.Parameter(torch.zeros(batch_size, 3, dtype=dtype))
        extra = np.load(JOINT_REGRESSOR_TRAIN_EXTRA)                      # lib/models/smpl.py:66-68
        self.register_buffer("J_regressor_extra", torch.tensor(extra, dtype=torch.float32))
        self.joint_map = torch.tensor(_SOURCES, dtype=torch.long)         # plain attribute, like the reference
        self._packed = None
        self._pack_key = None
        self._jreg_cache = {}
        # "fp32": strict FFMA blend (default, reference numerics); "bf16": tensor-core blend (<= 1 mm)
        self.blend_precision = "fp32"

    # ------------------------------------------------------------------ packing
    def _key(self):
        bufs = (self.v_template, self.shapedirs, self.posedirs, self.J_regressor, self.lbs_weights, self.J_regressor_extra)
        return tuple((b.device, b.data_ptr(), b._version) for b in bufs)

    def packed(self) -> PackedSmpl:
        key = self._key()
        if self._packed is None or self._pack_key != key:
            dev = self.v_template.device
            nv.require_cuda(self.v_template, "SMPL buffers (call .cuda() / .to(device) first)")
            self._packed = PackedSmpl(self.v_template, self.shapedirs, self.posedirs, self.J_regressor,
                                      self.parents, self.lbs_weights, dev)
            self._jreg_extra = pack_regressor(self.J_regressor_extra, self._packed.vp, dev)
            self._src49 = torch.tensor(joint_source_codes(_SOURCES), dtype=torch.int32, device=dev)
            self._pack_key = key
            self._jreg_cache = {}
        return self._packed

    def h36m_tables(self, J_regressor: torch.Tensor):
        """Packed [17,vp] regressor + joint codes for lib/models/spin.py:275-278."""
        p = self.packed()
        key = (J_regressor.device, J_regressor.data_ptr(), J_regressor._version, tuple(J_regressor.shape))
        hit = self._jreg_cache.get(key)
        if hit is None:
            jreg = pack_regressor(J_regressor, p.vp, p.device)
            src = torch.tensor([100 + j for j in H36M_TO_J14], dtype=torch.int32, device=p.device)
            hit = (jreg, src)
            self._jreg_cache = {key: hit}
        return hit

    # ------------------------------------------------------------------ forward
    @nv.device_guard
    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None, pose2rot=True, **kwargs):
        p = self.packed()
        betas = self.betas if betas is None else betas
        body_pose = self.body_pose if body_pose is None else body_pose
        global_orient = self.global_orient if global_orient is None else global_orient
        if transl is None and hasattr(self, "transl"):
            transl = self.transl
        n = max(betas.shape[0], body_pose.shape[0], global_orient.shape[0])
        dev = p.device
        betas_c = betas.detach().to(dev, torch.float32).expand(n, -1).contiguous()
        if pose2rot:
            full = torch.cat([global_orient.reshape(-1, 3), body_pose.reshape(-1, 69)], dim=1)
            kind, width = nv.POSE_AXIS_ANGLE, 72
        else:
            full = torch.cat([global_orient.reshape(-1, 1, 3, 3), body_pose.reshape(-1, 23, 3, 3)], dim=1)
            kind, width = nv.POSE_ROTMAT, 216
        flat = full.detach().to(dev, torch.float32).reshape(-1, width).expand(n, -1).contiguous()
        verts, joints, _, _, _ = smpl_forward_native(p, flat, width, kind, betas_c, 10, None, 0, n,
                                                     self._jreg_extra, self._src49, want_theta=False, want_rotmat=False,
                                                     blend_mode=1 if self.blend_precision == "bf16" else 0)
        if transl is not None:
            tr = transl.detach().to(dev, torch.float32)
            verts = verts + tr[:, None]
            joints = joints + tr[:, None]
        return SMPLOutput(vertices=verts, joints=joints, full_pose=full, betas=betas,
                          global_orient=global_orient, body_pose=body_pose)


def get_smpl_faces():
    """lib/models/smpl.py:87-89."""
    return SMPL(SMPL_MODEL_DIR, batch_size=1, create_transl=False).faces

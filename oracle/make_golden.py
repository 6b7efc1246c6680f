"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules
(/root/reference/lib/models/{tepose,spin,smpl}.py, lib/utils/geometry.py) on CPU
through oracle/ref_harness.py.  Authoring container only.

    python -m oracle.make_golden            # rewrites tests/golden/

Fixtures hold seeds + configuration + the reference's outputs; weights, SMPL-shaped
assets and inputs are regenerated from the seed by oracle/synth.py (numpy PCG64, not
torch RNG), so a fixture is a few hundred KB instead of hundreds of MB.
SMPL outputs come from the smplx stand-in (parity unpinned, see oracle/torch_ref.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import ref_harness, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> configuration of a TePose forward
FORWARD_CASES = {
    "fwd_L1_H64_B2_T4": dict(seed=11, batch=2, seqlen=4, n_layers=1, hidden=64),
    "fwd_L1_H128_B3_T16_h36m": dict(seed=12, batch=3, seqlen=16, n_layers=1, hidden=128, use_h36m=True),
    "fwd_L2_H64_B2_T6": dict(seed=13, batch=2, seqlen=6, n_layers=2, hidden=64),
    "fwd_L1_H64_B2_T5_train": dict(seed=14, batch=2, seqlen=5, n_layers=1, hidden=64, is_train=True),
    "fwd_L2_H96_B1_T3_train": dict(seed=15, batch=1, seqlen=3, n_layers=2, hidden=96, is_train=True),
}

# name -> configuration of a VIBE bootstrap forward (lib/models/vibe.py; evaluate.py:89-99 uses L2, add_linear)
VIBE_CASES = {
    "vibe_L2_H64_B2_T3_linear_h36m": dict(seed=21, batch=2, seqlen=3, n_layers=2, hidden=64, add_linear=True, use_h36m=True),
    "vibe_L1_H32_B1_T4_bidir": dict(seed=22, batch=1, seqlen=4, n_layers=1, hidden=32, bidirectional=True),
    "vibe_L1_H2048_B1_T2_plain": dict(seed=23, batch=1, seqlen=2, n_layers=1, hidden=2048),
}

# name -> configuration of a live-loop run (VIBE bootstrap + one TePose window per frame, thetas fed back)
STREAM_CASES = {
    "stream_T4_N9_L1_H64_vibeL2_H32": dict(seed=31, batch=1, frames=9, seqlen=4, n_layers=1, hidden=64,
                                           vibe_layers=2, vibe_hidden=32),
}


def edge_rotations() -> np.ndarray:
    """Rotation matrices that hit every branch of lib/utils/geometry.py:191-233:
    identity (s == 0), exact/near 180-degree turns about each axis (trace = -1),
    tiny angles, and random ones."""
    g = np.random.Generator(np.random.PCG64(77))
    mats = [np.eye(3), np.diag([1.0, -1, -1]), np.diag([-1.0, 1, -1]), np.diag([-1.0, -1, 1])]

    def rod(v):
        a = np.linalg.norm(v)
        k = v / a
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K

    for ang in (1e-7, 1e-4, 1e-2, 1.0, 2.0, 3.0, 3.14, 3.1415, np.pi - 1e-6):
        for _ in range(6):
            ax = g.standard_normal(3)
            mats.append(rod(ax / np.linalg.norm(ax) * ang))
    for _ in range(64):
        mats.append(rod(g.standard_normal(3) * 1.5))
    return np.stack(mats).astype(np.float32)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, cfg in VIBE_CASES.items():
        out = ref_harness.run_reference_vibe(**cfg)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"),
                            cfg=np.array(repr(cfg)), **{k: v.astype(np.float32) for k, v in out.items()})
        print(name, {k: v.shape for k, v in out.items()})
    for name, cfg in STREAM_CASES.items():
        out = ref_harness.run_reference_stream(**cfg)
        out["verts"] = out["verts"][:, ::4]            # every 4th frame's mesh keeps the fixture small
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"),
                            cfg=np.array(repr(cfg)), **{k: v.astype(np.float32) for k, v in out.items()})
        print(name, {k: v.shape for k, v in out.items()})
    if "--vibe-only" in sys.argv:
        return
    for name, cfg in FORWARD_CASES.items():
        out = ref_harness.run_reference(**cfg)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"),
                            cfg=np.array(repr(cfg)), **{k: v.astype(np.float32) for k, v in out.items()})
        print(name, {k: v.shape for k, v in out.items()})

    with ref_harness.reference_env(0) as mods:
        geo = mods.geometry
        g = np.random.Generator(np.random.PCG64(78))
        x6 = g.standard_normal((96, 6)).astype(np.float32)
        x6[0] = [1, 0, 0, 1, 0, 0]
        x6[1] = 0.0                                   # degenerate: both norms below eps
        x6[2] = [1e-7, 0, 0, 0, 0, 0]
        x6[3] = [1, 1, 2, 2, 3, 3]                    # a2 parallel to a1
        R = edge_rotations()
        aa = np.concatenate([g.standard_normal((60, 3)) * 0.7, np.zeros((2, 3)),
                             g.standard_normal((2, 3)) * 1e-9]).astype(np.float32)
        with torch.no_grad():
            rot6d = geo.rot6d_to_rotmat(torch.from_numpy(x6.copy())).numpy()
            r2aa = geo.rotation_matrix_to_angle_axis(torch.from_numpy(R.copy())).numpy()
            rod_q = geo.batch_rodrigues(torch.from_numpy(aa.copy())).numpy().reshape(-1, 3, 3)
            joints = torch.from_numpy(g.standard_normal((5, 49, 3)).astype(np.float32))
            cam = torch.from_numpy((np.array([0.9, 0, 0]) + 0.2 * g.standard_normal((5, 3))).astype(np.float32))
            proj = mods.spin.projection(joints, cam).numpy()
        np.savez_compressed(os.path.join(GOLDEN_DIR, "geometry.npz"),
                            x6=x6, rot6d=rot6d, R=R, r2aa=r2aa, aa=aa, rod_q=rod_q,
                            joints=joints.numpy(), cam=cam.numpy(), proj=proj)
        print("geometry", rot6d.shape, r2aa.shape, rod_q.shape, proj.shape)


if __name__ == "__main__":
    main()

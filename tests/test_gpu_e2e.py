"""GPU end-to-end parity: tepose_b200.TePose (CUDA kernels through the C ABI) against the
golden vectors produced by the unmodified reference and against the CPU oracle."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import synth, torch_ref
from tests.helpers import build_product_model, compare_outputs, oracle_forward

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


@pytest.mark.parametrize("fname", sorted(f for f in os.listdir(GOLD) if f.startswith("fwd_")))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_against_reference_golden(fname, precision):
    z = np.load(os.path.join(GOLD, fname))
    cfg = ast.literal_eval(str(z["cfg"]))
    model, sd = build_product_model(cfg["seed"], cfg["seqlen"], cfg["n_layers"], cfg["hidden"], precision, DEV)
    x = torch.from_numpy(synth.make_input(cfg["seed"], cfg["batch"], cfg["seqlen"])).to(DEV)
    Jr = torch_ref.SmplModel.synthetic(cfg["seed"]).J_regressor_h36m.to(DEV) if cfg.get("use_h36m") else None
    out = model(x, is_train=cfg.get("is_train", False), J_regressor=Jr)[-1]
    gold = {k: z[k] for k in ("theta", "verts", "kp_2d", "kp_3d", "rotmat")}
    if precision == "fp32":   # north_star: <= 1e-4 m on verts / joints in fp32
        errs = compare_outputs(out, gold, label=fname)
    else:                     # north_star: <= 1 mm in bf16
        errs = compare_outputs(out, gold, vert_tol=1e-3, rot_tol=2e-3, kp2d_tol=5e-2, label=fname)
    print(fname, precision, errs)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("L,H,T,B", [(1, 2048, 16, 32), (2, 1024, 6, 32), (1, 2048, 16, 1)])
def test_forward_full_size_against_oracle(L, H, T, B, precision):
    """BASELINE configs at full size (defaults L=1,H=2048,T=16; released L=2,H=1024,T=6)."""
    seed = 40 + L
    model, sd = build_product_model(seed, T, L, H, precision, DEV)
    x = synth.make_input(seed, B, T)
    ref, m = oracle_forward(seed, sd, x, L, H)
    out = model(torch.from_numpy(x).to(DEV))[-1]
    if precision == "fp32":
        errs = compare_outputs(out, ref, label=f"L{L}H{H}")
    else:
        errs = compare_outputs(out, ref, vert_tol=1e-3, rot_tol=2e-3, kp2d_tol=5e-2, label=f"L{L}H{H}")
        mpjpe = float((out["kp_3d"].cpu() - ref["kp_3d"]).norm(dim=-1).mean())
        assert mpjpe < 1e-3, mpjpe        # <= 1 mm MPJPE delta
    print(L, H, T, B, precision, errs)

"""Error structure of k_gru_umma against the float64 GRU cell (debug aid; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tepose_b200._native as nv
from tests.test_gpu_kernels import _gru_case, _dev_whh, _dev_whh_umma, cu, DEV


def run(B, T, H, seed=5):
    L = nv.lib()
    gi, w, b, h0, ys, hT = _gru_case(B, T, H, "bf16", seed, False, False)
    d = dict(gi=gi.to(DEV), w=_dev_whh(w, "bf16"), wu=_dev_whh_umma(w), b=cu(b), y=torch.zeros(T, B, H, device=DEV))
    j = nv.GruJob()
    j.gi, j.ldg, j.w_hh, j.b_hh = d["gi"].data_ptr(), 3 * H, d["w"].data_ptr(), d["b"].data_ptr()
    j.w_hh_umma = d["wu"].data_ptr()
    j.y, j.ldy, j.steps, j.t_in0, j.t_in_step, j.t_out0, j.t_out_step = d["y"].data_ptr(), H, T, 0, 1, 0, 1
    arr = (nv.GruJob * 1)(j)
    ws = nv.workspace(L.tp_gru_workspace_bytes(1, B, H), DEV)
    nv.check(L.tp_gru_recurrence(arr, 1, B, H, nv.PRECISION_BF16, nv.ptr(ws), ws.numel(), nv.stream()))
    torch.cuda.synchronize()
    e = (d["y"].cpu().double() - ys).abs()
    print(f"B={B} T={T} H={H}: max err {float(e.max()):.3e}; per step {[f'{float(x):.1e}' for x in e.amax(dim=(1, 2))]}")
    if float(e.max()) > 2e-4 and T >= 2:
        e1 = e[1]                                   # first step with a matmul
        per_u = e1.amax(dim=0).reshape(-1, 8).amax(dim=1)          # per 8-unit group
        print("  step 1, per 8-unit group (first 16):", [f"{float(x):.1e}" for x in per_u[:16]])
        print("  step 1, per batch row:", [f"{float(x):.1e}" for x in e1.amax(dim=1)])


if __name__ == "__main__":
    for (B, T, H) in [(32, 3, 128), (32, 3, 256), (32, 3, 1024), (32, 4, 2048), (4, 3, 2048)]:
        try:
            run(B, T, H)
        except Exception as ex:  # noqa: BLE001
            print(f"B={B} T={T} H={H}: FAILED {ex}")

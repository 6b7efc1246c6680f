"""Per-phase SM-clock breakdown of k_gru_umma at B=32,T=16,H=2048 (2 directions + the single-step direction)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tepose_b200._native as nv
from tepose_b200 import synthetic as synth
from tests.helpers import build_product_model
B, T, H = int(os.environ.get("B", 32)), 16, 2048
model, _ = build_product_model(0, T, 1, H, "bf16", "cuda:0")
x = torch.from_numpy(synth.make_input(0, B, T)).cuda()
enc = model.encoder
for _ in range(3):
    enc.encode_states(x)
torch.cuda.synchronize()
grid = 148
trace = torch.zeros(grid * T * 8, dtype=torch.int64, device="cuda")
nv.lib().tp_gru_set_trace(nv.vp(trace.data_ptr()))
enc.encode_states(x)
torch.cuda.synchronize()
nv.lib().tp_gru_set_trace(nv.vp(0))
tr = trace.cpu().numpy().reshape(grid, T, 8).astype(np.float64)
# slots: 0 step top (epi), 1 barrier seen (prod), 2 fences done (prod), 3 first h block landed (mma), 4 all MMAs issued (mma),
#        5 accumulator complete (epi), 6 partials exchanged (epi), 7 about to arrive on the grid barrier (epi)
if os.environ.get("TP_UM_TRACESET") == "2":
    for cta in (0, 1, 64, 65):
        rows = []
        for s in range(2, T - 1):
            t = tr[cta, s]
            rows.append([t[i] - t[0] for i in range(8)])
        m = np.mean(rows, axis=0)
        print(f"cta {cta:3d} MMA warp, cycles after stage 0 landed: " + ", ".join(f"st{i//2} {'issued' if i%2 else 'landed'}={m[i]:5.0f}" for i in range(8)))
    sys.exit(0)
if os.environ.get("TP_UM_TRACESET") == "1":
    for cta in (0, 1, 64, 65):
        rows = []
        for s in range(2, T - 1):
            t = tr[cta, s]
            rows.append([t[1] - t[5], t[2] - t[1], t[3] - t[2], t[6] - t[3], t[4] - t[6], t[7] - t[4]])
        m = np.mean(rows, axis=0)
        print(f"cta {cta:3d} epilogue: tmem ld={m[0]:5.0f}, remote stores+arrive={m[1]:5.0f}, local stores+arrive={m[2]:5.0f}, wait peer={m[3]:5.0f}, "
              f"gate math+h stores={m[4]:5.0f}, proxy fence+bar={m[5]:5.0f}")
    sys.exit(0)
names = ["top->barrier seen", "fences", "h block 0 lands", "mma issue", "issue->acc complete", "ld+dsmem exchange", "gates+stores+bar"]
for cta in (0, 1, 62, 63, 64, 65, 127):
    rows = []
    for s in range(2, T - 1):
        t = tr[cta, s]
        rows.append([t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6]])
    m = np.mean(rows, axis=0)
    tot = np.mean([tr[cta, s + 1, 0] - tr[cta, s, 0] for s in range(2, T - 2)])
    print(f"cta {cta:3d}: " + ", ".join(f"{n}={v:6.0f}" for n, v in zip(names, m)) + f" | step-to-step {tot:7.0f} cycles")
print("kernel span per CTA 0 (step 0 top -> last step top):", tr[0, T - 1, 0] - tr[0, 0, 0])
for cta in (0, 65, 127):
    t = tr[cta, 0]
    print(f"cta {cta:3d} prologue: setup+tmem alloc={t[2]-t[1]:7.0f}, TMEM fill={t[3]-t[2]:7.0f}, cluster sync={t[4]-t[3]:6.0f}, pdl wait={t[5]-t[4]:6.0f}, "
          f"smem W landed at {t[6]-t[1]:7.0f}, step 0 top at {t[0]-t[1]:7.0f}, kernel end at {tr[cta, T-1, 7]-t[1]:8.0f} cycles after start")

"""Debug: per-phase SM-clock breakdown of the bf16 GRU recurrence at B=32,T=16,H=2048 (2 jobs + 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tepose_b200._native as nv
from tepose_b200 import synthetic as synth
from tests.helpers import build_product_model
B, T, H = 32, 16, 2048
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
names = ["h staged (from step start)", "mma loop", "red write+sync", "gates+stores (to barrier entry)", "barrier"]
for cta in (0, 1, 63, 64, 100, 127):
    rows = []
    for s in range(1, T - 1):
        t = tr[cta, s]
        rows.append([t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]])
    m = np.mean(rows, axis=0)
    print(f"cta {cta:3d}: " + ", ".join(f"{n}={v:7.0f}" for n, v in zip(names, m)) + f"  | step total={m.sum():7.0f} cycles")
# spread of barrier entry across CTAs at a middle step
s = 8
ent = tr[:128, s, 4]; ext = tr[:128, s, 5]
print("step 8: barrier entry spread (cycles, not comparable across SMs exactly):", ent.max() - ent.min(), " exit spread:", ext.max() - ext.min())

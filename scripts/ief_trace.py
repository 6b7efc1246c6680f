"""Debug: per-layer SM-clock breakdown of the fused heads + IEF kernel at B=32 (14 layers: 3 head slices, fc1 feature
part, 3 x {fc1 state part, fc2, dec})."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tepose_b200._native as nv
from tepose_b200 import synthetic as synth
model, _ = synth.build_synthetic_model(0, 16, 1, 2048, "bf16", "cuda:0")
x = torch.from_numpy(synth.make_input(0, 32, 16)).cuda()
for _ in range(3):
    model(x)
torch.cuda.synchronize()
NL, G = 16, 128
trace = torch.zeros(max(G * NL * 8, 148 * 16 * 8), dtype=torch.int64, device="cuda")
h_fwd, h_rec = model.encoder.encode_states(x)
torch.cuda.synchronize()
nv.lib().tp_gru_set_trace(nv.vp(trace.data_ptr()))
model.regress_states(h_fwd, h_rec)
torch.cuda.synchronize()
nv.lib().tp_gru_set_trace(nv.vp(0))
tr = trace.cpu().numpy()[:G * NL * 8].reshape(G, NL, 8).astype(np.float64)
names = ["head k0", "head k1", "head k2", "fc1 x"] + [f"it{i} {n}" for i in range(3) for n in ("fc1 p", "fc2", "dec")]
for cta in (0, 5, 100):
    print(f"cta {cta}")
    for l in range(13):
        t = tr[cta, l]
        if t[0] == 0:
            continue
        print(f"   {names[l]:9s}: loads/stage={t[1]-t[0]:7.0f}  mma+red+epilogue={t[2]-t[1]:7.0f}  prefetch+barrier={t[3]-t[2]:7.0f}   layer total={t[3]-t[0]:7.0f}")
print("kernel span cycles (cta0, first layer start -> last layer end):", tr[0, 12, 3] - tr[0, 0, 0])

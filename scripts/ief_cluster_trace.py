"""Debug: time of the heads + IEF stage at B=32 with the IEF iterations on the 16-CTA cluster (k_ief_cluster) and inside the
grid-barrier kernel, and the SM-clock stamps of the cluster kernel's phases (CL_TRACE)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tepose_b200._native as nv
from tepose_b200 import synthetic as synth
B = int(os.environ.get("B", "32"))
model, _ = synth.build_synthetic_model(0, 16, 1, 2048, "bf16", "cuda:0")
x = torch.from_numpy(synth.make_input(0, B, 16)).cuda()
for _ in range(3):
    model(x)
torch.cuda.synchronize()
L = nv.lib()
h_fwd, h_rec = model.encoder.encode_states(x)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L.tp_set_pdl(0)
for cluster in (1, 0, 1):
    L.tp_set_ief_cluster(cluster)
    ts = []
    for rep in range(12):
        flush.zero_()
        model.encoder._sync_tail = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        psc = model._psc_from_states(h_fwd, h_rec)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"cluster={cluster}: heads+IEF (eager, PDL off, incl. barrier memset + launch gaps) us: min {min(ts):.1f} median {sorted(ts)[len(ts)//2]:.1f}")
L.tp_set_ief_cluster(1)
trace = torch.zeros(148 * 16 * 8, dtype=torch.int64, device="cuda")
L.tp_gru_set_trace(nv.vp(trace.data_ptr()))
model.encoder._sync_tail = None
model._psc_from_states(h_fwd, h_rec)
torch.cuda.synchronize()
L.tp_gru_set_trace(nv.vp(0))
tr = trace.cpu().numpy()
hb = tr[:128 * 16].reshape(128, 16).astype(np.float64)
hn = ["start(after pdl wait)", "hcat staged", "heads mma done", "group sync 1", "reduce + grid barrier", "feat staged", "fc1x mma + store", "group sync 2 + base"]
for cta in (0, 7, 64, 127):
    row = [f"{hn[i]} +{hb[cta, i] - hb[cta, i - 1]:.0f}" for i in range(1, 8)]
    print(f"k_heads_base cta {cta}: total {hb[cta, 7] - hb[cta, 0]:.0f} | " + " | ".join(row))
cl = tr[16384:16384 + 16 * 64].reshape(16, 64).astype(np.float64)
names = {1: "S wait done", 2: "fc1p + push", 3: "fc2 first 4 slices in", 4: "fc2 done", 5: "dec pushed", 6: "P wait done", 7: "owner done"}
for c in (0, 9, 15):
    print(f"k_ief_cluster cta {c}: total {cl[c,63]-cl[c,0]:.0f} cycles")
    for it in range(3):
        prev = cl[c, 0] if it == 0 else last
        row = []
        for s in range(1, 8):
            v = cl[c, s + it * 8]
            if v:
                row.append(f"{names[s]} +{v - prev:.0f}")
                prev = v
        last = prev
        print(f"   it{it}: " + " | ".join(row))

"""Debug: per-layer SM-clock breakdown of the fused IEF kernel at N=32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tepose_b200._native as nv
from tests.helpers import build_product_model
model, _ = build_product_model(0, 16, 1, 2048, "bf16", "cuda:0")
reg = model.regressor
feat = torch.randn(32, 2048, device="cuda")
feat._tp_bf16 = feat.to(torch.bfloat16)
for _ in range(3):
    reg(feat)
torch.cuda.synchronize()
trace = torch.zeros(64 * 12 * 8, dtype=torch.int64, device="cuda")
nv.lib().tp_gru_set_trace(nv.vp(trace.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); reg(feat); e1.record()
torch.cuda.synchronize()
nv.lib().tp_gru_set_trace(nv.vp(0))
print("regressor forward (IEF + SMPL) ms:", e0.elapsed_time(e1))
tr = trace.cpu().numpy().reshape(64, 12, 8).astype(np.float64)
for cta in (0, 5, 63):
    print(f"cta {cta}")
    for l in range(10):
        t = tr[cta, l]
        print(f"   layer {l}: stage A={t[1]-t[0]:7.0f} [issue {t[4]-t[0]:6.0f} | first half landed {t[5]-t[4]:6.0f} | warp0 done {t[6]-t[5]:6.0f} | sync {t[1]-t[6]:6.0f}]  mma+red+epilogue={t[2]-t[1]:7.0f}  prefetch+barrier={t[3]-t[2]:7.0f}")
print("kernel span cycles (cta0):", tr[0, 9, 3] - tr[0, 0, 0])

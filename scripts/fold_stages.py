"""Stage times (CUDA events between kernels, eager) with fold_linear on, B=32,T=16,H=2048 bf16."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tepose_b200._native as nv
from tepose_b200 import synthetic as synth
for B in (32, 1):
    model, _ = synth.build_synthetic_model(0, 16, 1, 2048, "bf16", "cuda:0")
    model.fold_linear = True
    x = torch.from_numpy(synth.make_input(0, B, 16)).cuda()
    acc = {}
    with torch.no_grad():
        for i in range(13):
            torch.cuda._sleep(20_000_000)
            nv.start_marks(); model(x); mk = nv.stop_marks(); torch.cuda.synchronize()
            if i >= 3:
                for (n0, a), (n1, b) in zip(mk[:-1], mk[1:]):
                    acc.setdefault(n1, []).append(a.elapsed_time(b))
    print("B", B, {k: round(float(np.mean(v)), 4) for k, v in acc.items()})

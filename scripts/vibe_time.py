"""Times the released VIBE bootstrap config (evaluate.py:89-99: L=2, H=1024, add_linear, residual) on cuda:0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tepose_b200 import _native as nv
from tepose_b200.synthetic import build_synthetic_vibe, make_vibe_input

for precision in ("bf16", "fp32"):
    model, _ = build_synthetic_vibe(7, 16, 2, 1024, True, False, True, precision, "cuda:0")
    for B in (1, 32):
        x = torch.from_numpy(make_vibe_input(7, B, 16)).cuda()
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            model(x)
        e1.record()
        torch.cuda.synchronize()
        nv.start_marks()
        model(x)
        torch.cuda.synchronize()
        mk = nv.stop_marks()
        stages = {b[0]: round(a[1].elapsed_time(b[1]), 4) for a, b in zip(mk[:-1], mk[1:])}
        print(f"vibe {precision} B={B} T=16: {e0.elapsed_time(e1) / n:.3f} ms per call ({B * 16} bodies)  stages {stages}")

"""Timing of the metric kernels at evaluation scale (N = 65 536 bodies / frames)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tepose_b200 import eval_utils as eu

def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

N = 65536
P = torch.randn(N, 14, 3, device="cuda"); G = P + 0.05 * torch.randn_like(P)
ms = timed(lambda: eu.pose_metrics(P, G))
print(f"pose_metrics (mpjpe + PA-mpjpe + accel) N={N}: {ms:.3f} ms  ({N / ms / 1e3:.1f} M frames/s)")
Nv = 16384
va = torch.randn(Nv, 6890, 3, device="cuda"); vb = va + 0.01
ms = timed(lambda: eu.compute_error_verts(va, target_verts=vb))
gb = 2 * Nv * 6890 * 12 / 1e9
print(f"vertex_error N={Nv}: {ms:.3f} ms  {gb / (ms * 1e-3):.0f} GB/s algorithmic (2 x 82 680 B per body)")

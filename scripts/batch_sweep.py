"""Forward time vs batch size (bf16, T=16, H=2048), eager with warm-up -- which kernels the larger batches land on."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tepose_b200 import synthetic as synth
from tepose_b200.graph import GraphedTePose
model, _ = synth.build_synthetic_model(0, 16, 1, 2048, "bf16", "cuda:0")
for B in (32, 48, 64, 96, 128):
    x = torch.from_numpy(synth.make_input(0, B, 16)).cuda()
    g = GraphedTePose(model, B, 16)
    for _ in range(3):
        g(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B:4d}: {ms:.3f} ms per forward  {B / ms:.1f} k frames/s  launches {g.launches_per_replay}")

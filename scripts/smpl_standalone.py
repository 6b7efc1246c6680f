"""Config 4 driver: SMPL forward (axis-angle in) for N bodies, times it; used under ncu too.
usage: smpl_standalone.py [N] [blend=bf16|fp32] [reps] [skinning=random|coherent]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tepose_b200 import synthetic as synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
blend = sys.argv[2] if len(sys.argv) > 2 else "bf16"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
skin = sys.argv[4] if len(sys.argv) > 4 else "random"
model, _ = synth.build_synthetic_model(0, 16, 1, 64, blend, "cuda:0", skinning=skin)
smpl = model.regressor.smpl
smpl.blend_precision = blend
b = synth.make_bodies(3, N)
aa = torch.from_numpy(b["pose_aa"]).cuda(); betas = torch.from_numpy(b["betas"]).cuda()
with torch.no_grad():
    for _ in range(2):
        out = smpl(betas=betas, body_pose=aa[:, 3:], global_orient=aa[:, :3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = smpl(betas=betas, body_pose=aa[:, 3:], global_orient=aa[:, :3])
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"smpl standalone N={N} blend={blend} skinning={skin}: {ms:.3f} ms  {N / ms / 1e3:.2f} M bodies/s  {N * 85780 / ms / 1e6:.0f} GB/s algorithmic")

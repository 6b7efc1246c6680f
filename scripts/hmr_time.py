"""Debug: one HMR.feature_extractor pass on 32 crops (for an ncu launch list: which layers cost what)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tepose_b200 import synthetic as psynth
m = psynth.build_synthetic_hmr(0, "cuda:0")
x = torch.from_numpy(psynth.make_image_batch(0, 32)).cuda()
with torch.no_grad():
    for _ in range(2):
        m.feature_extractor(x)
    torch.cuda.synchronize()
    m.feature_extractor(x)
torch.cuda.synchronize()

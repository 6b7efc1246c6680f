"""Golden outputs of the UNMODIFIED reference HMR (lib/models/spin.py:59-204: ResNet-50 feature_extractor + IEF + SMPL) on CPU fp32
-- TEST INFRASTRUCTURE, authoring container only (needs /root/reference).  The weights are not stored: tepose_b200.synthetic
.make_hmr_state(seed) regenerates them for both sides; the fixture holds the outputs only (tests/golden/hmr_N2.npz).
usage: python -m oracle.make_golden_hmr"""
import os

import numpy as np
import torch

from . import ref_harness
from tepose_b200 import synthetic as psynth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED, N = 11, 2


def run_reference_hmr(seed=SEED, n=N):
    x = psynth.make_image_batch(seed, n)
    with ref_harness.reference_env(seed) as mods:
        spin = mods.spin
        model = spin.hmr(pretrained=False)                     # HMR(Bottleneck, [3, 4, 6, 3], SMPL_MEAN_PARAMS)
        own = model.state_dict()
        sd = psynth.make_hmr_state({k: v.shape for k, v in own.items()}, seed)
        missing = [k for k in own if k not in sd and not k.startswith("smpl.") and not k.startswith("init_")]
        assert not missing, missing
        for k, v in sd.items():
            assert tuple(own[k].shape) == tuple(v.shape), k
            own[k] = torch.as_tensor(v)
        model.load_state_dict(own, strict=True)
        model.eval()
        with torch.no_grad():
            xf = model.feature_extractor(torch.from_numpy(x))
            xf2, out = model(torch.from_numpy(x), return_features=True)
        assert torch.equal(xf, xf2)
        res = {"xf": xf.numpy().copy(), "keys": np.array(sorted(k for k in own if not k.startswith("smpl.")))}
        res.update({k: v.detach().numpy().copy() for k, v in out[0].items()})
    return res


def main():
    res = run_reference_hmr()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "hmr_N2.npz"), cfg=np.array(repr({"seed": SEED, "n": N})),
                        **{k: (v if k == "keys" else v.astype(np.float32)) for k, v in res.items()})
    print({k: (v.shape, float(np.abs(v).mean())) for k, v in res.items() if k != "keys"})


if __name__ == "__main__":
    main()

"""CUDA-graph capture of a fixed-shape TePose forward: the ~20 kernel launches of one step are
replayed as a single graph launch (launch-bound inner loop)."""
from __future__ import annotations

import torch

from . import _native as nv

OUTPUT_KEYS = ("theta", "verts", "kp_2d", "kp_3d", "rotmat")


class GraphedTePose:
    """model: tepose_b200.TePose on a CUDA device, in eval mode.

    g = GraphedTePose(model, batch, seqlen, J_regressor=None)
    out = g(x)            # x [B,T,2133] float32, CUDA or pinned host; returns the static output dict
    g.replay()            # re-run on whatever is in g.static_input
    """

    def __init__(self, model, batch, seqlen, J_regressor=None, is_train=False, warmup=2, input_dtype=torch.float32):
        p = next(model.parameters())
        nv.require_cuda(p, "model parameters")
        self.model, self.device = model, p.device
        if input_dtype not in (torch.float32, torch.float16):
            raise ValueError("input_dtype must be torch.float32 or torch.float16")
        self.static_input = torch.zeros(batch, seqlen, 2133, device=self.device, dtype=input_dtype)
        self.J_regressor = None if J_regressor is None else J_regressor.to(self.device)
        self.is_train = is_train
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):     # packs weights, sets func attributes, warms the allocator
                self.model(self.static_input, is_train=is_train, J_regressor=self.J_regressor)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        before = nv.lib().tp_launch_count()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_output = self.model(self.static_input, is_train=is_train, J_regressor=self.J_regressor)[-1]
        self.launches_per_replay = int(nv.lib().tp_launch_count() - before)

    @nv.device_guard
    def replay(self):
        self.graph.replay()
        return self.static_output

    @nv.device_guard
    def __call__(self, x: torch.Tensor):
        self.static_input.copy_(x, non_blocking=True)
        return self.replay()


class GraphedHMRFeatures:
    """`HMR.feature_extractor` on a fixed [batch,3,size,size] input as ONE CUDA graph (the 76 launches of the ResNet-50 otherwise
    pay the host's launch cadence): `g = GraphedHMRFeatures(hmr, 32); xf = g(images)`; `g.static_input`, `g.replay()` as above."""

    def __init__(self, model, batch, size=224, warmup=2):
        p = next(model.parameters())
        nv.require_cuda(p, "model parameters")
        self.model, self.device = model, p.device
        self.static_input = torch.zeros(batch, 3, size, size, device=self.device, dtype=torch.float32)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):     # folds / packs the weights, warms the allocator
                self.model.feature_extractor(self.static_input)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        before = nv.lib().tp_launch_count()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_output = self.model.feature_extractor(self.static_input)
        self.launches_per_replay = int(nv.lib().tp_launch_count() - before)

    @nv.device_guard
    def replay(self):
        self.graph.replay()
        return self.static_output

    @nv.device_guard
    def __call__(self, x: torch.Tensor):
        self.static_input.copy_(x, non_blocking=True)
        return self.replay()

"""Live-stream (causal, carried-state) mode of the TePose hot path -- SURVEY.md F3/F4, section 8 a15.

The reference's live loop (evaluate.py:247-269, demo.py:238-252) re-runs the whole T-frame window
from h0 = 0 for every new frame (O(N*T) GRU steps) and feeds each prediction's theta back into the
inputs of later windows.  For n_layers == 1 the encoder is causal (F3): gru_fwd and the *_reverse
direction of gru_rec both consume frames in arrival order, and the remaining direction is a single
step on the newest frame.  This module keeps the two causal hidden states on the device and advances
them ONE committed step per frame (O(N)):

    committed state C_{t-1}  = state after frames <= t-1, each fed with its own predicted theta
    frame t arrives          : C_{t-1}   = step(C_{t-2}, [feat_{t-1}, theta_{t-1}])   (commit)
                               tentative = step(C_{t-1}, [feat_t, 0])                 (newest frame: theta slot 0,
                                                                                        evaluate.py:252)
                               hS        = step(0, [feat_t, 0]) with gru_rec's forward weights
                               output_t  = Regressor(heads(tentative, [hS, tentative_B]))

This is the reference's computation for a window that starts at the first frame of the stream
instead of T-1 frames ago; it is a NEW mode (F4), pinned against torch.nn.GRU stepped with explicit
h0 (oracle/torch_ref.encoder_causal_states), not against evaluate.py's windowed outputs.
"""
from __future__ import annotations

import torch

from . import _native as nv

FEAT = 2048


class LiveTePose:
    """stream = LiveTePose(model, batch=1); out = stream.step(features_t [B,2048]) per frame."""

    def __init__(self, model, batch=1, J_regressor=None, use_graph=True):
        if model.encoder.n_layers != 1:
            raise ValueError("live-stream carried state needs n_layers == 1 (SURVEY.md H5)")
        p = next(model.parameters())
        nv.require_cuda(p, "model parameters")
        self.model, self.device, self.B = model, p.device, batch
        self.H = model.encoder.hidden_size
        self.J_regressor = None if J_regressor is None else J_regressor.to(self.device)
        dev, B, H = self.device, batch, self.H
        self.x2 = torch.zeros(B, 2, 2133, device=dev)            # [previous frame + its theta | newest frame + 0]
        self.hF = torch.zeros(B, H, device=dev)                   # committed states C_{t-2} (before the commit step)
        self.hB = torch.zeros(B, H, device=dev)
        self.frames = 0
        self.use_graph = use_graph
        self._graph = None
        self._out = None

    def reset(self):
        self.x2.zero_(); self.hF.zero_(); self.hB.zero_()
        self.frames = 0

    # one frame, eager: first frame of a stream has no previous frame to commit
    def _first(self):
        enc = self.model.encoder
        h_fwd, h_rec = enc.encode_states(self.x2[:, 1:2])
        return self._decode(h_fwd, h_rec)

    def _steady(self):
        enc = self.model.encoder
        h_fwd, h_rec, seq_f, seq_b = enc.encode_states(self.x2, h0=(self.hF, self.hB), return_states=True)
        out = self._decode(h_fwd, h_rec)
        # carry: committed state after the previous frame; the newest frame becomes "previous" with its theta
        self.hF.copy_(seq_f[0]); self.hB.copy_(seq_b[0])
        return out

    def _decode(self, h_fwd, h_rec):
        out = self.model.regress_states(h_fwd, h_rec, J_regressor=self.J_regressor)[-1]
        self.x2[:, 0, :FEAT].copy_(self.x2[:, 1, :FEAT])
        self.x2[:, 0, FEAT:].copy_(out["theta"])
        return out

    @torch.no_grad()
    @nv.device_guard
    def step(self, features: torch.Tensor):
        """features [B,2048] (CUDA or pinned host).  Returns {theta, verts, kp_2d, kp_3d, rotmat} for
        the newest frame (tensors are reused by the next call when graphs are on)."""
        self.x2[:, 1, :FEAT].copy_(features, non_blocking=True)
        if self.frames == 0:
            out = self._first()
        elif not self.use_graph:
            out = self._steady()
        else:
            if self._graph is None:
                # warm up on a side stream (packs weights, sets kernel attributes), restoring the state after
                keep = (self.x2.clone(), self.hF.clone(), self.hB.clone())
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    self._steady()
                torch.cuda.current_stream(self.device).wait_stream(side)
                torch.cuda.synchronize(self.device)
                for dst, src in zip((self.x2, self.hF, self.hB), keep):
                    dst.copy_(src)
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._out = self._steady()
                for dst, src in zip((self.x2, self.hF, self.hB), keep):
                    dst.copy_(src)
            self._graph.replay()
            out = self._out
        self.frames += 1
        return out

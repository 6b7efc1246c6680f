#!/usr/bin/env python
"""Headline benchmark: output frames/s of the TePose per-sequence inference hot path at
B=32, T=16 (BASELINE.json configs[1]: encoder + IEF Regressor + SMPL LBS, L=1, H=2048).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--impl reference]

One "step" = one forward of one batch of B sequences = B output frames.
  value  : whole-job frames/s with inputs resident in HBM, CUDA-event timed per step, L2 flushed
           between steps (outside the event pairs), max over ranks.
  e2e    : the same metric through the public API with HOST buffers: pinned H2D of x, forward,
           D2H of all five outputs inside the timed region.
  roofline / cpu_baseline : see DESIGN.md "Measurement".
Under torchrun every rank runs the same per-GPU batch (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B, T, L, H = 32, 16, 1, 2048
SEED = 0
METRIC = "output frames/s at B=32,T=16 (TePose encoder + IEF Regressor + SMPL)"
TRAIN_METRIC = "training sequences/s at B=32,T=16 per GPU (forward + backward + gradient all-reduce + Adam)"
TRAIN_WORKLOAD = "BASELINE.json configs[4]: training step B=32,T=16, L=1,H=2048 (encoder + Regressor + SMPL), data-parallel replicas"
WORKLOAD = "BASELINE.json configs[1]: 3DPW-eval-shaped batched inference B=32,T=16, L=1,H=2048, 2133-d inputs"


def active_switches():
    """TP_* environment switches that change which kernel runs (recorded in the bench line's config)."""
    return {k: v for k, v in sorted(os.environ.items()) if k.startswith("TP_")}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML every ~2 ms; the timed
    region of this benchmark is only tens of milliseconds, too short for polling nvidia-smi)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mask, self.max_mhz, self._stop_evt = index, [], 0, None, threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self.sm.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
                    self.mask |= int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                else:
                    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                          "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    a, b_ = [float(v) for v in out.split(",")]
                    self.sm.append(a); self.max_mhz = b_
            except Exception:
                pass
            self._stop_evt.wait(0.002)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=6)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def oracle_step_fn(threads):
    """The CPU restatement of the reference path (oracle/torch_ref.py, kind 'port'): same
    torch.nn.GRU / Linear / smplx-style LBS ops the reference executes on CPU."""
    from oracle import synth, torch_ref
    torch.set_num_threads(threads)
    sd = synth.make_state_dict(SEED, L, H)
    m = torch_ref.SmplModel.synthetic(SEED)
    grus = (torch_ref.build_gru(sd, "gru_fwd", L, H, False), torch_ref.build_gru(sd, "gru_rec", L, H, True))
    sd_t = {k: torch.as_tensor(v) for k, v in sd.items()}
    xs = [torch.from_numpy(synth.make_input(SEED + i, B, T)) for i in range(2)]

    def step(i):
        return torch_ref.tepose_forward(sd_t, m, xs[i % 2], L, H, grus=grus)
    return step


def cpu_other_configs(threads):
    """the same CPU port on BASELINE configs[0] (one sequence, B=1,T=16: the live-loop frame) and on the
    SMPL-standalone shape (4096 bodies, axis-angle in, as lib/utils/eval_utils.py:164 chunks them); best of 5 after 2 warm-ups."""
    from oracle import synth, torch_ref
    torch.set_num_threads(threads)
    sd = {k: torch.as_tensor(v) for k, v in synth.make_state_dict(SEED, L, H).items()}
    m = torch_ref.SmplModel.synthetic(SEED)
    grus = (torch_ref.build_gru(sd, "gru_fwd", L, H, False), torch_ref.build_gru(sd, "gru_rec", L, H, True))
    x1 = torch.from_numpy(synth.make_input(SEED, 1, T))
    bod = synth.make_bodies(SEED, 4096)
    aa, betas = torch.from_numpy(bod["pose_aa"]), torch.from_numpy(bod["betas"])

    def best(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts), float(np.median(ts))
    with torch.no_grad():
        b1 = best(lambda: torch_ref.tepose_forward(sd, m, x1, L, H, grus=grus))
        sm = best(lambda: torch_ref.smpl_forward(m, betas, pose_aa=aa), reps=3, warm=1)
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            model = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), "")
    except OSError:
        pass
    return {"B1_T16_ms": {"best": 1e3 * b1[0], "median": 1e3 * b1[1]},
            "smpl_4096_bodies": {"best_ms": 1e3 * sm[0], "median_ms": 1e3 * sm[1], "bodies_per_s": 4096 / sm[1]},
            "cpu_model": model, "threads": threads}


def time_cpu(step, budget_s, min_iters=2, warm=1):
    for i in range(warm):
        step(i)
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < min_iters or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        step(len(times))
        times.append(time.perf_counter() - t0)
        if len(times) >= 200:
            break
    return times


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port) with all
    host threads, on the same config / metric.  Rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    step = oracle_step_fn(cores)
    for i in range(max(2, args.warmup)):   # oneDNN primitive creation makes the first calls 5x slower
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    dt = time.perf_counter() - t0
    fps = B * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch": B, "seqlen": T, "n_layers": L, "hidden": H},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} forwards of the full B=32,T=16 batch (oracle/torch_ref.py, torch {torch.__version__} CPU)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_reference_gpu(args, rank):
    """Informative second baseline: the reference's own op sequence as PyTorch executes it on the
    B200 (cuDNN GRU, cuBLAS Linear, ~40 ATen launches of LBS), fp32, no CUDA graph.  Not a bench arm."""
    if rank != 0:
        return
    from oracle import synth, torch_ref
    dev = torch.device("cuda", 0)
    sd = synth.make_state_dict(SEED, L, H)
    m = torch_ref.SmplModel.synthetic(SEED)
    for k, v in vars(m).items():
        if torch.is_tensor(v):
            setattr(m, k, v.to(dev))
    grus = (torch_ref.build_gru(sd, "gru_fwd", L, H, False).to(dev), torch_ref.build_gru(sd, "gru_rec", L, H, True).to(dev))
    sd_t = {k: torch.as_tensor(v).to(dev) for k, v in sd.items()}
    x = torch.from_numpy(synth.make_input(SEED, B, T)).to(dev)
    import oracle.torch_ref as tr
    for _ in range(max(3, args.warmup)):
        tr.tepose_forward(sd_t, m, x, L, H, grus=grus)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tr.tepose_forward(sd_t, m, x, L, H, grus=grus)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "reference-gpu", "metric": METRIC, "value": B / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
                      "steps": args.steps, "dtype": "f32", "note": "torch ops of the reference path on cuda:0 (cuDNN/cuBLAS/ATen), eager"}),
          flush=True)


def run_train_reference(args, rank):
    """--config train --impl reference: the training step of the reference path on the host CPU (oracle/train_ref.py: the torch
    ops the reference runs + torch.autograd + torch.optim.Adam, lib/core/trainer.py:203,235-237), all host threads.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import synth, train_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(SEED, L, H)
    orc = train_ref.TrainOracle(sd, SEED, L, H)
    params = [p for _, p in orc.named_parameters()]
    opt = torch.optim.Adam(params, lr=5e-5)
    x = synth.make_input(SEED, B, T)
    masks = train_ref.make_masks(SEED, 2 * B)
    tgt = train_ref.make_targets(SEED, 2 * B)

    def step():
        orc.loss_and_grads(x, masks, tgt)
        opt.step()
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    n = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    dt = (time.perf_counter() - t0) / n
    line = {"impl": "reference", "metric": TRAIN_METRIC, "value": B / dt, "unit": "sequences/s", "n_gpus": args.gpus, "steps": n,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": TRAIN_WORKLOAD, "batch": B, "seqlen": T, "n_layers": L, "hidden": H},
            "cpu_baseline": {"value": B / dt, "unit": "sequences/s", "cores": cores, "kind": "port",
                             "sample": f"{n} training steps of the full B=32,T=16 batch (oracle/train_ref.py, torch {torch.__version__} CPU autograd + Adam)"},
            "e2e": {"value": B / dt, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measure_training(args, rank, world, dev, dist, steps, warmup, precision="fp32", tensor_core_grads=False):
    """BASELINE.json configs[4]: forward + backward + (for world > 1) NCCL gradient all-reduce + Adam step, B=32,T=16 per GPU.
    Returns a dict (rank-aggregated: max over ranks of the CUDA-event time).  The step is measured twice when world > 1:
    with the all-reduce and with it switched off, which gives the exposed-communication share."""
    from tepose_b200 import synthetic as synth
    from tepose_b200.synthetic import build_synthetic_model
    from tepose_b200.train import DataParallel, path_parameters, make_dropout_masks
    from tepose_b200 import shard as _sh
    import tepose_b200._native as nv
    model, _ = build_synthetic_model(SEED, T, L, H, precision, dev)
    model.train()
    model.train_tensor_core_grads = bool(tensor_core_grads)
    params = [p for _, p in path_parameters(model)]
    opt = torch.optim.Adam(params, lr=5e-5, fused=True)          # lib/utils/utils.py:145-152, lr of configs/*.yaml
    dp = DataParallel(model, opt)
    xs = [torch.from_numpy(synth.make_input(SEED + 100 * rank + i, B, T)).pin_memory() for i in range(2)]
    tgt = {k: torch.from_numpy(v).to(dev) for k, v in synth.make_train_targets(SEED + rank, 2 * B).items()}
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    loss_fn = lambda out: synth.synthetic_train_loss(out, tgt)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(n, use_sync):
        real_world = dp.sync.world
        if not use_sync:
            dp.sync.world = 1
        launches0 = nv.lib().tp_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        last = None
        for i in range(n):
            x = xs[i % 2].to(dev, non_blocking=True)            # host -> device copy of the step's input inside the timed region
            last = dp.step(x, loss_fn, make_dropout_masks(2 * B, dev, generator=gen))
        loss_host = float(last)                                  # device -> host read of the step's result
        e1.record()
        barrier()
        dp.sync.world = real_world
        ms = e0.elapsed_time(e1)
        return _sh.max_over_ranks([ms], dev)[0] / n, loss_host, int(nv.lib().tp_launch_count() - launches0) // max(n, 1)

    timed(max(warmup, 3), True)
    ms_sync, loss, launches = timed(steps, True)
    out = {"ms_per_step": ms_sync, "sequences_per_s": B * world / (ms_sync * 1e-3), "loss": loss, "launches_per_step": launches,
           "precision": precision, "tensor_core_weight_grads": bool(tensor_core_grads), "optimizer": "torch.optim.Adam(fused=True), lr 5e-5",
           "params": int(sum(p.numel() for p in params)), "h2d_bytes_per_step": int(xs[0].numel() * 4), "d2h_bytes_per_step": 4}
    if world > 1:
        ms_nosync, _, _ = timed(steps, False)
        out.update(allreduce_bytes_per_step=int(dp.sync.bytes) if dp.sync.bytes else int(sum(p.numel() for p in params) * 4),
                   ms_per_step_without_allreduce=ms_nosync, exposed_comm_frac=max(0.0, (ms_sync - ms_nosync) / ms_sync),
                   collective="ncclAllReduce(sum) per gradient bucket on a side stream, issued from inside the backward; 1/world scale")
    del dp, opt, model
    torch.cuda.empty_cache()
    return out


def _shard_max(v, dev):
    from tepose_b200 import shard as _sh
    return _sh.max_over_ranks([v], dev)[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-gpu"],
                    help="reference: CPU port of the reference path (the reference arm); reference-gpu: the same torch ops "
                         "(nn.GRU -> cuDNN, F.linear -> cuBLAS, smplx-style LBS) on cuda:0, informative only")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-live", action="store_true", help="skip the live-stream latency measurement")
    ap.add_argument("--no-fold", action="store_true", help="skip the fold_linear measurement")
    ap.add_argument("--no-smpl", action="store_true", help="skip the SMPL-standalone (65,536 bodies) measurement")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step block of the default line")
    ap.add_argument("--config", default="infer", choices=["infer", "train"],
                    help="train: BASELINE.json configs[4] (training step, NCCL gradient all-reduce when launched with torchrun) "
                         "as the headline of the printed line")
    ap.add_argument("--train-precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--tc-grads", action="store_true", help="training: weight-gradient GEMMs in bf16 on the tcgen05 GEMM")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.config == "train":
            run_train_reference(args, rank)
        else:
            run_reference(args, rank, world)
        return
    if args.impl == "reference-gpu":
        run_reference_gpu(args, rank)
        return

    import torch.distributed as dist
    # the product arm touches tepose_b200 only (oracle/ is used by the cpu_baseline leg further down)
    from tepose_b200 import synthetic as synth
    from tepose_b200.synthetic import build_synthetic_model as build_product_model
    import tepose_b200._native as nv
    from tepose_b200.graph import GraphedTePose, OUTPUT_KEYS

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path for the product)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # rank 0 prints exactly ONE line on stdout: NCCL writes its version banner to fd 1 when the communicator is
        # created, so fd 1 points at stderr until the first collective is through
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    if args.config == "train":
        sampler = ClockSampler(local_rank); sampler.start()
        tr = measure_training(args, rank, world, dev, dist, args.steps, args.warmup, args.train_precision, args.tc_grads)
        clocks = sampler.finish()
        if rank == 0:
            cpu_baseline = None
            if world == 1 and args.cpu_budget > 0:
                from oracle import synth as osynth, train_ref
                cores = os.cpu_count() or 1
                torch.set_num_threads(cores)
                orc = train_ref.TrainOracle(osynth.make_state_dict(SEED, L, H), SEED, L, H)
                opt = torch.optim.Adam([p for _, p in orc.named_parameters()], lr=5e-5)
                xo, mo, to = osynth.make_input(SEED, B, T), train_ref.make_masks(SEED, 2 * B), train_ref.make_targets(SEED, 2 * B)
                times = time_cpu(lambda i: (orc.loss_and_grads(xo, mo, to), opt.step()), min(args.cpu_budget, 15.0))
                cpu_baseline = {"value": B / float(np.median(times)), "unit": "sequences/s", "cores": cores, "kind": "port",
                                "sample": f"{len(times)} training steps of the full batch, median {1e3 * float(np.median(times)):.0f} ms "
                                          f"(oracle/train_ref.py: torch {torch.__version__} CPU autograd + Adam)"}
            line = {"metric": TRAIN_METRIC, "value": tr["sequences_per_s"], "unit": "sequences/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": tr["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f32" if args.train_precision == "fp32" else "bf16", "data": "synthetic",
                    "config": {"workload": TRAIN_WORKLOAD, "batch_per_gpu": B, "seqlen": T, "n_layers": L, "hidden": H,
                               "precision": args.train_precision, "parallelism": f"dp{world}", "switches": active_switches()},
                    "e2e": {"value": tr["sequences_per_s"], "unit": "sequences/s", "h2d_bytes_per_step": tr["h2d_bytes_per_step"],
                            "d2h_bytes_per_step": tr["d2h_bytes_per_step"],
                            "note": "the timed step copies its input from pinned host memory and reads the loss back"},
                    "gpu_launches": tr["launches_per_step"] * args.steps, "clocks": clocks, "training": tr, "cpu_baseline": cpu_baseline}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    model, _ = build_product_model(SEED, T, L, H, args.precision, dev)
    n_inputs = 4
    xs_host = [torch.from_numpy(synth.make_input(SEED + 100 * rank + i, B, T)).pin_memory() for i in range(n_inputs)]
    xs_dev = [x.to(dev) for x in xs_host]
    graphed = GraphedTePose(model, B, T)
    run = (lambda: model(graphed.static_input)[-1]) if args.no_graph else graphed.replay
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    lib = nv.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------------------------------------------------------- device-resident throughput
    for i in range(args.warmup):
        graphed.static_input.copy_(xs_dev[i % n_inputs]); run()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    launches0 = lib.tp_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    with torch.no_grad():
        for i in range(args.steps):
            graphed.static_input.copy_(xs_dev[i % n_inputs])
            flush.zero_()                                   # evict weights / inputs from L2 (not timed)
            evs[i][0].record()
            run()
            evs[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    dev_ms_total = float(step_ms.sum())
    eager_launches = int(lib.tp_launch_count() - launches0)
    launches = eager_launches if args.no_graph else graphed.launches_per_replay * args.steps

    # ---------------------------------------------------------------- end to end (host buffers)
    # public host-buffer API: pinned x -> H2D -> forward -> D2H of all five outputs, every step; copies of
    # neighbouring steps overlap compute through a 3-slot ring (tepose_b200/pipeline.py)
    from tepose_b200.pipeline import PipelinedTePose
    pipe = PipelinedTePose(model, B, T, depth=3)
    h2d_bytes, d2h_bytes = pipe.h2d_bytes, pipe.d2h_bytes
    with torch.no_grad():
        for i in range(args.warmup):
            pipe.result(pipe.submit(xs_host[i % n_inputs]))
        pipe.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        checksum = 0.0
        e0.record()
        t_e2e0 = time.perf_counter()
        tickets = []
        for i in range(args.steps):
            tickets.append(pipe.submit(xs_host[i % n_inputs]))
            if len(tickets) >= pipe.depth:                       # consume results as they complete
                checksum += float(pipe.result(tickets.pop(0))["theta"][0, 0])
        for tk in tickets:
            checksum += float(pipe.result(tk)["theta"][0, 0])
        pipe.drain()
        e1.record()
        barrier()
        t_e2e = time.perf_counter() - t_e2e0
    e2e_ms_total = max(e0.elapsed_time(e1), 1e3 * t_e2e)     # host-visible completion: the slower of the two clocks
    clocks = sampler.finish()

    # ---------------------------------------------------------------- per-kernel timing (roofline)
    stage_ms = {}
    lib.tp_set_pdl(0)             # events between kernels need plain stream order: no programmatic overlap in this pass
    with torch.no_grad():
        for i in range(min(args.steps, 20)):
            flush.zero_()
            torch.cuda._sleep(20_000_000)         # ~10 ms: let the CPU run ahead so events bracket GPU work only
            nv.start_marks()
            model(xs_dev[i % n_inputs])
            marks = nv.stop_marks()
            torch.cuda.synchronize(dev)
            for (n0, a), (n1, b_) in zip(marks[:-1], marks[1:]):
                stage_ms.setdefault(n1, []).append(a.elapsed_time(b_))
    lib.tp_set_pdl(1)
    stage_avg = {k: float(np.median(v)) for k, v in stage_ms.items()}     # median: an allocator cudaMalloc in one iteration must not skew a stage
    hbm_peak, tf_peak, peak_src = peaks()
    wbytes = 2 if args.precision == "bf16" else 4
    k2_ms = stage_avg.get("k2_recurrence_l0", float("nan"))
    k2_bytes = 2 * 3 * H * H * wbytes                       # W_hh of the two full directions, read once
    k1_ms = stage_avg.get("k1_input_proj_l0", float("nan"))
    k1_flops = 2.0 * (2 * B * T + B) * 2133 * 3 * H
    roofline = {
        "kernel": "k_gru_bf16_dual (K2 recurrence)" if args.precision == "bf16" else "k_gru_f32 (K2 recurrence)",
        "bound": "hbm", "achieved": k2_bytes / (k2_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / hbm_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        # (profiles/r02_ncu_k_gru_bf16_dual.txt: 76.4 MB + 2.4 MB; = W_hh once + the K1 gate pre-activations once)
        "traffic": 78.9e6 if args.precision == "bf16" else None, "peak_source": peak_src,
        # what actually bounds it: W_hh is re-streamed from L2 every step (it cannot stay on chip: 50.3 MB vs 33.6 MB of
        # shared memory).  L2 -> SM bytes per launch from the same capture (lts__t_sectors_srcunit_tex_op_read x 32 B)
        # against the L2 streaming rate scripts/micro/l2bw.cu reaches with 128 CTAs (profiles/r01_l2bw_micro.txt)
        "l2_stream": ({"bytes_per_launch": 1.0325e9, "achieved_GBps": 1.0325e9 / (k2_ms * 1e-3) / 1e9, "micro_ceiling_GBps": 6900.0,
                       "frac": 1.0325e9 / (k2_ms * 1e-3) / 1e9 / 6900.0} if args.precision == "bf16" else None),
        "algorithmic_bytes": k2_bytes, "avg_ms": k2_ms,
        "secondary": {"kernel": "k_gemm_bf16_tc (K1 input projection)" if args.precision == "bf16" else "k_gemm_f32 (K1)",
                      "bound": "tensor", "achieved": k1_flops / (k1_ms * 1e-3) / 1e12, "peak": tf_peak,
                      "unit": "TFLOP/s", "frac": k1_flops / (k1_ms * 1e-3) / 1e12 / tf_peak, "avg_ms": k1_ms},
    }
    # every stage against the roofline that bounds it: algorithmic bytes / flops per launch over its average time
    def _hbm(name, nbytes, what):
        ms = stage_avg.get(name)
        return None if not ms else {"stage": name, "bound": "hbm", "what": what, "algorithmic_bytes": nbytes, "avg_ms": ms,
                                    "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak}
    k3_w = (3 * H * 2048 + 2048 * 1024 + 3 * (160 * 1024 + 1024 * 1024 + 1024 * 160)) * wbytes     # heads + fc1 feature part + 3 IEF iterations
    k3_ms = (stage_avg.get("k3_heads_ief") or (stage_avg.get("k3_heads", 0.0) + stage_avg.get("k3_ief", 0.0))) or None
    roofline["stages"] = [x for x in (
        _hbm("pack", B * T * 2133 * 4 + B * T * 2176 * wbytes, "x fp32 in, padded time-major operand out"),
        {"stage": "k1_input_proj_l0", "bound": "tensor", "achieved": k1_flops / (k1_ms * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
         "frac": k1_flops / (k1_ms * 1e-3) / 1e12 / tf_peak, "avg_ms": k1_ms, "what": "minimal FLOPs of the three directions' input projection"},
        {"stage": "k2_recurrence_l0", "bound": "hbm", "achieved": k2_bytes / (k2_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
         "frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / hbm_peak, "avg_ms": k2_ms, "what": "W_hh of both directions read once (see l2_stream)"},
        None if not k3_ms else {"stage": "k3_heads_ief", "bound": "hbm", "what": "every Linear weight of the heads and the IEF read once (a 14-layer dependent chain)",
                                "algorithmic_bytes": k3_w, "avg_ms": k3_ms, "achieved": k3_w / (k3_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                "unit": "GB/s", "frac": k3_w / (k3_ms * 1e-3) / 1e9 / hbm_peak},
        _hbm("k45_smpl", B * 85780, "916 B in + 84 864 B out per body (SURVEY 8d)"),
    ) if x]

    # ---------------------------------------------------------------- live-stream latency (config 3, rank 0)
    live = None
    if rank == 0 and not args.no_live:
        from tepose_b200.live import LiveTePose
        stream = LiveTePose(model, batch=1)
        feats = torch.from_numpy(synth.make_input(SEED + 7, 1, 64)[0, :, :2048]).to(dev)
        n_live = 400
        lev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_live)]
        for i in range(50):
            stream.step(feats[i % 64][None])
        torch.cuda.synchronize(dev)
        for i in range(n_live):
            lev[i][0].record()
            stream.step(feats[i % 64][None])          # device-resident feature row in, theta fed back on device
            lev[i][1].record()
        torch.cuda.synchronize(dev)
        lat = np.array([a.elapsed_time(b_) for a, b_ in lev])
        live = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)), "frames": n_live,
                "mode": "carried-state causal step, B=1, theta feedback on device, CUDA graph per frame"}
        # the reference's own loop (evaluate.py:247-269): a full T-frame window from h0 = 0 per frame, thetas fed back
        from tepose_b200.stream import WindowedTePose
        ws = WindowedTePose(model, None, batch=1)
        ws.set_theta(torch.zeros(T - 1, 85))
        wfeats = torch.from_numpy(synth.make_input(SEED + 8, 1, 64 + T)[:, :, :2048]).to(dev)
        for i in range(50):
            ws.step(wfeats[:, i % 64:i % 64 + T])
        torch.cuda.synchronize(dev)
        for i in range(n_live):
            lev[i][0].record()
            ws.step(wfeats[:, i % 64:i % 64 + T])
            lev[i][1].record()
        torch.cuda.synchronize(dev)
        lat = np.array([a.elapsed_time(b_) for a, b_ in lev])
        live["windowed"] = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
                            "mode": f"reference loop: {T}-frame window per frame, B=1, window + theta ring in HBM, "
                                    "CUDA graph per frame"}

    # ---------------------------------------------------------------- SMPL standalone (config 4 shape, this rank's shard)
    smpl_sa = None
    if not args.no_smpl:
        from tepose_b200 import shard as _shard
        n_total = 65536
        lo, hi = _shard.partition(n_total, world, rank)
        bodies = synth.make_bodies(SEED + rank, hi - lo)
        smpl = model.regressor.smpl
        smpl.blend_precision = "bf16" if args.precision == "bf16" else "fp32"     # K4 on tensor cores in bf16 mode
        aa = torch.from_numpy(bodies["pose_aa"]).to(dev)
        betas = torch.from_numpy(bodies["betas"]).to(dev)
        with torch.no_grad():
            smpl(betas=betas, body_pose=aa[:, 3:], global_orient=aa[:, :3])
            torch.cuda.synchronize(dev)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(3):
                smpl(betas=betas, body_pose=aa[:, 3:], global_orient=aa[:, :3])
            s1.record()
            torch.cuda.synchronize(dev)
        sm_ms = s0.elapsed_time(s1) / 3
        bps, sm_ms_max = _shard.aggregate_throughput(hi - lo, sm_ms, dev)
        smpl.blend_precision = "fp32"
        smpl_sa = {"bodies": n_total, "blend": args.precision, "bodies_per_s": bps, "ms": sm_ms_max, "algorithmic_GBps_per_gpu":
                   (hi - lo) * 85780 / (sm_ms * 1e-3) / 1e9, "hbm_frac": (hi - lo) * 85780 / (sm_ms * 1e-3) / 1e9 / hbm_peak}
        del aa, betas

    # ---------------------------------------------------------------- opt-in: heads + IEF folded into one affine map
    folded = None
    if not args.no_graph and not args.no_fold:
        model.fold_linear = True
        g2 = GraphedTePose(model, B, T)
        fev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        with torch.no_grad():
            for i in range(args.warmup):
                g2.static_input.copy_(xs_dev[i % n_inputs]); g2.replay()
            barrier()
            for i in range(args.steps):
                g2.static_input.copy_(xs_dev[i % n_inputs])
                flush.zero_()
                fev[i][0].record()
                g2.replay()
                fev[i][1].record()
            barrier()
        f_ms = float(np.sum([a.elapsed_time(b_) for a, b_ in fev]))
        f_ms_max = _shard_max(f_ms, dev)
        folded = {"ms_per_step": f_ms_max / args.steps, "value": B * args.steps * world / (f_ms_max * 1e-3), "unit": "frames/s",
                  "launches_per_step": g2.launches_per_replay,
                  "what": "TePose(fold_linear=True): eval heads + 3 IEF iterations pre-composed in float64 into one "
                          "[3H,160] fp32 GEMM (same outputs within the mode's tolerance; NOT the headline)"}
        model.fold_linear = False
        del g2

    # ---------------------------------------------------------------- second shape: the released checkpoint's architecture
    released = None
    if rank == 0 and not args.no_graph and not args.no_fold:
        Lr, Hr, Tr = 2, 1024, 6
        m2, _ = build_product_model(SEED + 1, Tr, Lr, Hr, args.precision, dev)
        g3 = GraphedTePose(m2, B, Tr)
        xr = torch.from_numpy(synth.make_input(SEED + 9, B, Tr)).to(dev)
        rev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        with torch.no_grad():
            for i in range(args.warmup):
                g3.static_input.copy_(xr); g3.replay()
            torch.cuda.synchronize(dev)
            for i in range(args.steps):
                g3.static_input.copy_(xr)
                flush.zero_()
                rev[i][0].record()
                g3.replay()
                rev[i][1].record()
            torch.cuda.synchronize(dev)
        r_ms = float(np.mean([a.elapsed_time(b_) for a, b_ in rev]))
        released = {"config": {"batch": B, "seqlen": Tr, "n_layers": Lr, "hidden": Hr}, "ms_per_step": r_ms,
                    "frames_per_s": B / (r_ms * 1e-3), "launches_per_step": g3.launches_per_replay}
        del g3, m2

    # ---------------------------------------------------------------- training step (configs[4]; all ranks: it holds the all-reduce)
    training = None
    if not args.no_train:
        training = measure_training(args, rank, world, dev, dist, min(args.steps, 10), 3, "fp32", False)

    # ---------------------------------------------------------------- aggregate over ranks
    from tepose_b200 import shard as _sh
    dev_ms_max, e2e_ms_max = _sh.max_over_ranks([dev_ms_total, e2e_ms_total], dev)
    frames = B * args.steps * world
    value = frames / (dev_ms_max * 1e-3)
    e2e_value = frames / (e2e_ms_max * 1e-3)

    cpu_baseline = None
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        times = time_cpu(oracle_step_fn(cores), args.cpu_budget)
        cpu_fps = B / float(np.median(times))
        cpu_baseline = {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                        "sample": f"{len(times)} forwards of the full B=32,T=16 batch, median "
                                  f"{1e3 * float(np.median(times)):.1f} ms (oracle/torch_ref.py on torch {torch.__version__} CPU)"}
        if args.cpu_budget >= 4:
            cpu_baseline["other_configs"] = cpu_other_configs(cores)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "seqlen": T, "n_layers": L, "hidden": H,
                       "precision": args.precision, "cuda_graph": not args.no_graph, "switches": active_switches(),
                       "l2": "256 MiB memset between steps, outside the per-step event pairs",
                       "parallelism": f"batch-sharded x{world}, no collective"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms_max / args.steps, "pipeline_depth": 3},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "stages_ms": stage_avg,
            "live": live,
            "folded": folded,
            "released_config": released,
            "smpl_standalone": smpl_sa,
            "training": training,
            "step_ms": {"min": float(step_ms.min()), "median": float(np.median(step_ms)), "max": float(step_ms.max())},
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

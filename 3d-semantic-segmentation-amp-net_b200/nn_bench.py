"""bench.py hooks for the PointNet-attention workloads (BASELINE.json configs[0] and configs[2]).

  fwd    segmentation forward of 32 x 2048-point blocks (eval mode), the loop of
         test_pointnet_att_segmen.py:160-181 / train_pointnet-attention.py:396-450 at W = 1
  train  one training step of train_pointnet-attention.py:396-470: forward (train mode, dropout 0.3),
         CE(weight, ignore -1) + 0.001 reg, backward, 2 x Adam

Weights are random-init from the reference constructors' layout (oracle-independent: torch default init under
torch.manual_seed), inputs are the synthetic blocks of SURVEY 8(d) generated here with numpy.
"""
import json
import os

import numpy as np
import torch

NN_BATCH, NN_POINTS, NN_DIMS, NN_CLASSES = 32, 2048, 9, 5
FWD_FLOP_PER_POINT = 413143.0           # algorithmic forward FLOPs per point at N=2048, W=1
TRAIN_FLOP_PER_POINT = 3 * FWD_FLOP_PER_POINT


def _peaks():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d["bf16_tflops_sustained"], "measured (sustained cuBLAS bf16)"
    return 1400.0, "fallback"


def synthetic_blocks(rank, B=NN_BATCH, N=NN_POINTS):
    rng = np.random.default_rng(2000 + rank)
    x = rng.random((B, N, NN_DIMS), dtype=np.float32)
    x[:, :, :2] = x[:, :, :2] * 2 - 1
    x[:, :, 2] *= 0.3
    cent = x[:, :, :2].mean(1, keepdims=True)          # [B, 1, 2]
    tg = rng.integers(0, NN_CLASSES, (B, N)).astype(np.int64)
    return x, cent, tg


def build_modules(amp, device, dropout=0.3):
    torch.manual_seed(0)
    enc = amp.BasePointNet(point_dimension=3, return_local_features=True, global_feat_dim=256, device=device).to(device)
    seg = amp.SegmentationWithAttention(256, 8, num_classes=NN_CLASSES, local_dim=64, dropout=dropout, device=device).to(device)
    return enc, seg


def forward_pass(enc, seg, x, cent):
    """One window-block pass exactly as the scripts drive the modules (W = 1)."""
    out, ft = enc(x)
    local_feat = out[:, :, -64:]
    global_feat = out[:, 0, :-64].view(-1, 1, 256)
    gl = torch.transpose(global_feat, 0, 1)
    logits, _ = seg(gl, local_feat, cent, [local_feat.shape[1]], None)
    return logits, ft


def _timed(dist, fn, steps, warmup, flush):
    for _ in range(warmup):
        flush(); fn()
    dist.barrier()
    evs = []
    for _ in range(steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    dist.barrier()
    return dist.max_over_ranks(float(sum(a.elapsed_time(b) for a, b in evs)))


class _Flush:
    def __init__(self, device, mib=256):
        self.buf = torch.empty(mib << 20, dtype=torch.uint8, device=device)

    def __call__(self):
        self.buf.fill_(1)


def bench_fwd(dist, amp, steps, warmup, with_cpu, precision="fp32"):
    dev = dist.device
    enc, seg = build_modules(amp, dev)
    enc.eval(); seg.eval()
    enc.precision = seg.precision = precision
    x_np, c_np, _ = synthetic_blocks(dist.rank)
    x_host = torch.from_numpy(x_np).pin_memory()
    c_host = torch.from_numpy(c_np).pin_memory()
    x, cent = x_host.to(dev), c_host.to(dev)
    flush = _Flush(dev)
    keep = {}

    def step():
        keep["logits"], _ = forward_pass(enc, seg, x, cent)

    n0 = amp._lib.launch_count()
    ms = _timed(dist, step, steps, warmup, flush)
    launches = (amp._lib.launch_count() - n0) // (steps + warmup) * steps
    logits_host = torch.empty((NN_BATCH, NN_CLASSES, NN_POINTS), dtype=torch.float32).pin_memory()

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        cd = c_host.to(dev, non_blocking=True)
        lg, _ = forward_pass(enc, seg, xd, cd)
        logits_host.copy_(lg, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e_ms = _timed(dist, step_e2e, steps, warmup, flush)
    pts = NN_BATCH * NN_POINTS * dist.world * steps
    peak, src = _peaks()
    ach = (NN_BATCH * NN_POINTS * FWD_FLOP_PER_POINT) / (ms / steps * 1e-3) / 1e12
    res = {
        "value": pts / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms / steps, "gpu_launches": int(launches),
        "e2e": {"value": pts / (e_ms * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": int(x_host.numel() * 4 + c_host.numel() * 4), "d2h_bytes_per_step": int(logits_host.numel() * 4)},
        "roofline": {"bound": "tensor", "kernel": "whole forward (pw_linear_kernel chain, fp32 CUDA cores)" if precision == "fp32"
                     else "whole forward (tc_chain_kernel x 4: tcgen05 bf16 chains + fp32 per-cloud FC / attention kernels)",
                     "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": src,
                     "model": "413 143 algorithmic FLOP per point (SURVEY 8d) x 65 536 points / step time"},
        "config": {"workload": "configs[0]: segmentation forward, batch %d x %d points, 9 channels, eval, random-init weights"
                               % (NN_BATCH, NN_POINTS), "l2": "flushed between steps (256 MiB write)", "precision": precision},
        "dtype": "f32" if precision == "fp32" else "bf16",
    }
    if with_cpu:
        from oracle import nn_bench as onb
        res["cpu_baseline"] = onb.cpu_forward(sample_steps=3)
    return res


def bench_train(dist, amp, steps, warmup, with_cpu):
    dev = dist.device
    enc, seg = build_modules(amp, dev)
    enc.train(); seg.train()
    params = list(enc.parameters()) + list(seg.parameters())
    if dist.pg:
        for p in params:
            torch.distributed.broadcast(p.data, 0)
    opt_e = torch.optim.Adam(enc.parameters(), lr=1e-3)
    opt_s = torch.optim.Adam(seg.parameters(), lr=1e-3)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=dev), reduction="mean", ignore_index=-1)
    x_np, c_np, t_np = synthetic_blocks(dist.rank)
    x_host, c_host, t_host = (torch.from_numpy(a).pin_memory() for a in (x_np, c_np, t_np))
    x, cent, tg = x_host.to(dev), c_host.to(dev), t_host.to(dev)
    eye = torch.eye(64, device=dev)
    flush = _Flush(dev)
    keep = {}
    flat = None
    if dist.pg:
        from .parallel import GradAllReduce
        flat = GradAllReduce(params, dist.world)

    def train_step(xd, cd, td):
        opt_e.zero_grad(set_to_none=True); opt_s.zero_grad(set_to_none=True)
        logits, ft = forward_pass(enc, seg, xd, cd)
        loss = ce(logits, td) + 0.001 * torch.norm(eye - torch.bmm(ft, ft.transpose(2, 1)))
        loss.backward()
        if flat is not None:
            flat.all_reduce()
        opt_e.step(); opt_s.step()
        keep["loss"] = loss.detach()

    n0 = amp._lib.launch_count()
    ms = _timed(dist, lambda: train_step(x, cent, tg), steps, warmup, flush)
    launches = (amp._lib.launch_count() - n0) // (steps + warmup) * steps
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_e2e():
        train_step(x_host.to(dev, non_blocking=True), c_host.to(dev, non_blocking=True), t_host.to(dev, non_blocking=True))
        loss_host.copy_(keep["loss"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e_ms = _timed(dist, step_e2e, steps, warmup, flush)
    pts = NN_BATCH * NN_POINTS * dist.world * steps
    peak, src = _peaks()
    ach = (NN_BATCH * NN_POINTS * TRAIN_FLOP_PER_POINT) / (ms / steps * 1e-3) / 1e12
    res = {
        "value": pts / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms / steps, "gpu_launches": int(launches),
        "e2e": {"value": pts / (e_ms * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": int(x_host.numel() * 4 + c_host.numel() * 4 + t_host.numel() * 8), "d2h_bytes_per_step": 4},
        "roofline": {"bound": "tensor", "kernel": "whole step (pw_linear / wgrad chains, fp32 CUDA cores)", "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": src,
                     "model": "3 x 413 143 algorithmic FLOP per point (SURVEY 8d) x 65 536 points / step time"},
        "config": {"workload": "configs[2]: training step fwd+loss+bwd+2xAdam, batch %d x %d points per GPU, dropout 0.3%s"
                               % (NN_BATCH, NN_POINTS, ", NCCL gradient all-reduce" if dist.pg else ""),
                   "l2": "flushed between steps (256 MiB write)", "precision": "fp32"},
        "dtype": "f32", "final_loss": float(keep["loss"]),
    }
    if with_cpu:
        from oracle import nn_bench as onb
        res["cpu_baseline"] = onb.cpu_train(sample_steps=2)
    return res


def bench_fwd_bf16(dist, amp, steps, warmup, with_cpu):
    return bench_fwd(dist, amp, steps, warmup, with_cpu, precision="bf16")


def hooks():
    return {"fwd": bench_fwd, "fwd_bf16": bench_fwd_bf16, "train": bench_train}

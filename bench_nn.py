"""bench.py workloads of the PointNet-attention stages (BASELINE.json configs[0], configs[2], configs[3]); lives next to bench.py
(not in the product package: its cpu_baseline legs import oracle/).

  fwd    segmentation forward of 32 x 2048-point blocks (eval mode), the loop of
         test_pointnet_att_segmen.py:160-181 / train_pointnet-attention.py:396-450 at W = 1
  train  one training step of train_pointnet-attention.py:396-470: forward (train mode, dropout 0.3),
         CE(weight, ignore -1) + 0.001 reg, backward, 2 x Adam

Weights are random-init from the reference constructors' layout (oracle-independent: torch default init under
torch.manual_seed), inputs are the synthetic blocks of SURVEY 8(d) generated here with numpy.
"""
import json
import os
import sys

import numpy as np
import torch

NN_BATCH, NN_POINTS, NN_DIMS, NN_CLASSES = 32, 2048, 9, 5
FWD_FLOP_PER_POINT = 413143.0           # algorithmic forward FLOPs per point at N=2048, W=1
TRAIN_FLOP_PER_POINT = 3 * FWD_FLOP_PER_POINT


def _peaks():
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")
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


FWD_CHAIN32_DRAM_BYTES = 42.6e6


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

    # the forward is ~30 short dependent launches: capture it once in a CUDA graph (static shapes, the library only enqueues
    # on the capturing stream and owns no memory) and replay it; eager calls are measured next to it
    n0 = amp._lib.launch_count()
    for _ in range(2):
        keep["logits"], _ = forward_pass(enc, seg, x, cent)
    launches_per_step = (amp._lib.launch_count() - n0) // 2
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        forward_pass(enc, seg, x, cent)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep["logits"], _ = forward_pass(enc, seg, x, cent)

    ms = _timed(dist, graph.replay, steps, warmup, flush)
    eager_ms = _timed(dist, lambda: forward_pass(enc, seg, x, cent), steps, warmup, flush)
    launches = launches_per_step * steps
    logits_host = torch.empty((NN_BATCH, NN_CLASSES, NN_POINTS), dtype=torch.float32).pin_memory()

    # end to end: pinned host inputs -> the static input tensors -> the same captured forward -> logits to pinned host memory,
    # the three copies captured in the same graph as memcpy nodes (one replay + one stream synchronise per step)
    graph_e2e = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph_e2e):
        x.copy_(x_host, non_blocking=True)
        cent.copy_(c_host, non_blocking=True)
        lg_e2e, _ = forward_pass(enc, seg, x, cent)
        logits_host.copy_(lg_e2e, non_blocking=True)

    def step_e2e():
        graph_e2e.replay()
        torch.cuda.current_stream().synchronize()

    serial_ms = _timed(dist, step_e2e, steps, warmup, flush)
    del graph_e2e

    # end to end as a serving loop runs it: amp.StreamedForward keeps 3 batches in flight (copy-in / run / copy-out streams, one
    # captured forward per slot). Every step still copies ITS inputs from pinned host memory and ITS logits back inside the
    # timed region; the inputs rotate over a host pool larger than L2 (64 batches = 151 MB) and every slot has its own
    # activation buffers, so nothing a step reads is left over from the step before (no flush kernel in this region: it would
    # sit on the run stream between the graphs). AMP_BENCH_SERIAL_E2E=1 reports the serial loop above instead.
    pool_n = 64
    pool = []
    for i in range(pool_n):
        xp, cp, _ = synthetic_blocks(1000 + dist.rank * pool_n + i)
        pool.append((torch.from_numpy(xp).pin_memory(), torch.from_numpy(cp).pin_memory()))
    depth = 3
    pipe = amp.StreamedForward(lambda a, b: forward_pass(enc, seg, a, b)[0], pool[0], device=dev, depth=depth)

    def pipelined(n_steps, first):
        pend = []
        for i in range(n_steps):
            if len(pend) == depth:
                pipe.result(pend.pop(0))
            pend.append(pipe.submit(*pool[(first + i) % pool_n]))
        for t in pend:
            pipe.result(t)

    pipelined(max(warmup, depth), 0)
    torch.cuda.synchronize(dev)
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(pipe.s_in)
    pipelined(steps, warmup)
    ev1.record(pipe.s_out)
    ev1.synchronize()
    torch.cuda.synchronize(dev)
    dist.barrier()
    piped_ms = dist.max_over_ranks(float(ev0.elapsed_time(ev1)))
    serial = os.environ.get("AMP_BENCH_SERIAL_E2E") == "1"
    e_ms = serial_ms if serial else piped_ms
    pts = NN_BATCH * NN_POINTS * dist.world * steps
    peak, src = _peaks()
    ach = (NN_BATCH * NN_POINTS * FWD_FLOP_PER_POINT) / (ms / steps * 1e-3) / 1e12
    res = {
        "value": pts / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms / steps, "gpu_launches": int(launches),
        "e2e": {"value": pts / (e_ms * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": int(x_host.numel() * 4 + c_host.numel() * 4), "d2h_bytes_per_step": int(logits_host.numel() * 4)},
        "roofline": {"bound": "tensor", "kernel": "whole forward: tc_chain32_kernel x 5 (split-bf16 3-MMA tcgen05 chains, activations in TMEM) + fp32 per-cloud FC / attention kernels" if precision == "fp32"
                     else "whole forward: tc_chain_kernel x 4 (tcgen05 bf16 chains) + fp32 per-cloud FC / attention kernels",
                     "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s", "frac": ach / peak,
                     # dram__bytes_read + write summed over the five tc_chain32 launches of one forward, ncu --set full
                     # (profiles/r02_tc_chain32_ncu.txt; the outputs stay in the 126 MB L2, so writes are ~0)
                     "traffic": FWD_CHAIN32_DRAM_BYTES if precision == "fp32" else None, "peak_source": src,
                     # what the tensor pipe executes: three bf16 MMAs per product on the fp32-class path, one on the bf16 path
                     "executed_frac": (3.0 if precision == "fp32" else 1.0) * ach / peak,
                     "model": "413 143 algorithmic FLOP per point (SURVEY 8d) x 65 536 points / step time; the fp32 path executes 3 MMAs per product"},
        "config": {"workload": "configs[0]: segmentation forward, batch %d x %d points, 9 channels, eval, random-init weights" % (NN_BATCH, NN_POINTS)},
        "notes": {"l2": "flushed between steps (256 MiB write)", "precision": precision,
                  "launch": "CUDA graph replay of the two module calls (eager: %.3f ms per step)" % (eager_ms / steps),
                  "e2e": ("serial loop: copy in, forward, copy out, synchronise; L2 flushed between steps" if serial else
                          "StreamedForward, 3 batches in flight, per-step H2D + D2H timed, 151 MB pinned input pool (> L2), no "
                          "flush; serial loop (L2 flushed): %.0f points/s" % (pts / (serial_ms * 1e-3)))},
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
    # the script's two Adam optimizers (train_pointnet-attention.py:141-142), torch's fused multi-tensor implementation
    # the script's two Adam optimizers (train_pointnet-attention.py:141-142) as amp.FusedAdam: one launch per optimizer step
    # (AMP_BENCH_TORCH_ADAM=1: torch's fused multi-tensor implementation instead)
    torch_adam = os.environ.get("AMP_BENCH_TORCH_ADAM") == "1"
    if torch_adam:
        opt_e = torch.optim.Adam(enc.parameters(), lr=1e-3, fused=True, capturable=True)
        opt_s = torch.optim.Adam(seg.parameters(), lr=1e-3, fused=True, capturable=True)
    else:
        opt_e = amp.FusedAdam(enc.parameters(), lr=1e-3)
        opt_s = amp.FusedAdam(seg.parameters(), lr=1e-3)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=dev), reduction="mean", ignore_index=-1)
    x_np, c_np, t_np = synthetic_blocks(dist.rank)
    x_host, c_host, t_host = (torch.from_numpy(a).pin_memory() for a in (x_np, c_np, t_np))
    x, cent, tg = x_host.to(dev), c_host.to(dev), t_host.to(dev)
    eye = torch.eye(64, device=dev)
    flush = _Flush(dev)
    keep = {}
    flat = None
    if dist.pg:
        GradAllReduce = amp.GradAllReduce
        flat = GradAllReduce(params, dist.world, zero_copy=True)      # .grad = slices of one flat buffer, written in place

    def train_step(xd, cd, td):
        if flat is None:                                              # (zero-copy gradients are overwritten by every backward)
            opt_e.zero_grad(set_to_none=True); opt_s.zero_grad(set_to_none=True)
        logits, ft = forward_pass(enc, seg, xd, cd)
        loss = ce(logits, td) + 0.001 * torch.norm(eye - torch.bmm(ft, ft.transpose(2, 1)))
        loss.backward()
        if flat is not None:
            flat.all_reduce()
        opt_e.step(); opt_s.step()
        keep["loss"] = loss.detach()

    n0 = amp._lib.launch_count()
    eager_ms = _timed(dist, lambda: train_step(x, cent, tg), steps, warmup, flush)
    launches = (amp._lib.launch_count() - n0) // (steps + warmup) * steps
    # The step is ~180 short dependent kernels: capture zero_grad + forward + loss + backward + (all-reduce) + 2 x Adam once
    # in a CUDA graph (amp.GraphedStep: static tensors, a device-side dropout offset so that every replay draws a new mask)
    # and replay it. The eager step -- what the reference's train_loop does -- is measured next to it.
    graphed, ms = None, eager_ms
    if os.environ.get("AMP_BENCH_EAGER_TRAIN") != "1":
        try:
            graphed = amp.GraphedStep(lambda: train_step(x, cent, tg), device=dev)
            ms = _timed(dist, graphed, steps, warmup, flush)
        except Exception as e:                                   # capture refused (e.g. a collective that cannot be captured)
            sys.stderr.write("bench_train: CUDA graph capture failed (%s); eager numbers\n" % (str(e).splitlines()[0],))
            graphed, ms = None, eager_ms
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_e2e():
        if graphed is not None:                                      # pinned host inputs -> the graph's static tensors -> replay
            x.copy_(x_host, non_blocking=True); cent.copy_(c_host, non_blocking=True); tg.copy_(t_host, non_blocking=True)
            graphed()
        else:
            train_step(x_host.to(dev, non_blocking=True), c_host.to(dev, non_blocking=True), t_host.to(dev, non_blocking=True))
        loss_host.copy_(keep["loss"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e_ms = _timed(dist, step_e2e, steps, warmup, flush)
    if graphed is not None:
        graphed.close()
    pts = NN_BATCH * NN_POINTS * dist.world * steps
    peak, src = _peaks()
    ach = (NN_BATCH * NN_POINTS * TRAIN_FLOP_PER_POINT) / (ms / steps * 1e-3) / 1e12
    res = {
        "value": pts / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms / steps, "gpu_launches": int(launches),
        "e2e": {"value": pts / (e_ms * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": int(x_host.numel() * 4 + c_host.numel() * 4 + t_host.numel() * 8), "d2h_bytes_per_step": 4},
        "roofline": {"bound": "tensor", "kernel": "whole step (tc_layer_kernel / tc_wgrad_kernel: split-bf16 3-MMA tcgen05 GEMMs at fp32-class accuracy, few-row fp32 kernels, torch Adam)", "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "peak_source": src,
                     "model": "3 x 413 143 algorithmic FLOP per point (SURVEY 8d) x 65 536 points / step time"},
        "config": {"workload": "configs[2]: training step fwd+loss+bwd+2xAdam, batch %d x %d points per GPU" % (NN_BATCH, NN_POINTS)},
        "notes": {"l2": "flushed between steps (256 MiB write)", "precision": "fp32", "dropout": 0.3, "adam": "torch fused" if torch_adam else "amp.FusedAdam (one launch per optimizer)",
                  "launch": ("CUDA graph replay of the whole step (eager: %.3f ms per step)" % (eager_ms / steps)) if graphed is not None
                            else "eager",
                  "collective": "NCCL gradient all-reduce (AVG, one flat buffer)" if dist.pg else "none (1 GPU)"},
        "dtype": "f32", "final_loss": float(keep["loss"]),
    }
    if with_cpu:
        from oracle import nn_bench as onb
        res["cpu_baseline"] = onb.cpu_train(sample_steps=2)
    return res


def bench_train_w9(dist, amp, steps, warmup, with_cpu):
    """The training step at the shape the script really runs (train_pointnet-attention.py:337-475 on a collate_seq_padd batch):
    W = 9 windows of 2048 points per sample, batch 32: device-side batch assembly + augmentation (one H2D, one kernel), 9 encoder
    calls (one BatchNorm group each), attention + head over 9 x 2048 points per sample, CE(ignore -1) + reg, backward, 2 x Adam.
    Both the device-timed value and e2e start from the collated batch in pinned HOST memory (the assembly is part of the step)."""
    dev = dist.device
    W, B, N = 9, NN_BATCH, NN_POINTS
    enc, seg = build_modules(amp, dev)
    enc.train(); seg.train()
    params = list(enc.parameters()) + list(seg.parameters())
    if dist.pg:
        for p in params:
            torch.distributed.broadcast(p.data, 0)
    opt_e = amp.FusedAdam(enc.parameters(), lr=1e-3)
    opt_s = amp.FusedAdam(seg.parameters(), lr=1e-3)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=dev), reduction="mean", ignore_index=-1)
    rng = np.random.default_rng(5000 + dist.rank)
    pc = rng.random((B, N, NN_DIMS, W), dtype=np.float32)
    pc[:, :, :2, :] = pc[:, :, :2, :] * 2 - 1
    pc[:, :, 2, :] *= 0.3
    tg = rng.integers(0, NN_CLASSES, (B, N, W)).astype(np.int64)
    real = rng.integers(3, W + 1, B)                                 # trailing windows: replicas with targets -1 (collate_fns.py:42-44)
    for b in range(B):
        pc[b, :, :, real[b]:] = pc[b, :, :, real[b] - 1:real[b]]
        tg[b, :, real[b]:] = -1
    pc_host, tg_host = torch.from_numpy(pc).pin_memory(), torch.from_numpy(tg).pin_memory()
    cent = torch.from_numpy(pc[:, :, :2, :].mean(1).transpose(0, 2, 1).copy()).to(dev)      # [B, W, 2]
    eye = torch.eye(64, device=dev)
    flush = _Flush(dev)
    keep = {}
    flat = amp.GradAllReduce(params, dist.world, zero_copy=True) if dist.pg else None
    np.random.seed(7)

    def train_step():
        opt_e.zero_grad(set_to_none=True); opt_s.zero_grad(set_to_none=True)      # as the script does (:372-373)
        x, targets_pc = amp.assemble_windows(pc_host, tg_host, train=True, device=dev)
        lo, gl, npc, ft = [], [], [], None
        for w in range(W):
            out, ft = enc(x[w])
            lo.append(out[:, :, -64:]); gl.append(out[:, 0, :-64].view(-1, 1, 256)); npc.append(N)
        logits, _ = seg(torch.transpose(torch.cat(gl, 1), 0, 1), torch.cat(lo, 1), cent, npc, None)
        loss = ce(logits, targets_pc) + 0.001 * torch.norm(eye - torch.bmm(ft, ft.transpose(2, 1)))
        loss.backward()
        if flat is not None:
            flat.all_reduce()
        opt_e.step(); opt_s.step()
        keep["loss"] = loss.detach()

    n0 = amp._lib.launch_count()
    ms = _timed(dist, train_step, steps, warmup, flush)
    launches = (amp._lib.launch_count() - n0) // (steps + warmup) * steps
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_e2e():
        train_step()
        loss_host.copy_(keep["loss"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e_ms = _timed(dist, step_e2e, steps, warmup, flush)
    pts = B * W * N * dist.world * steps
    peak, src = _peaks()
    ach = (B * W * N * TRAIN_FLOP_PER_POINT) / (ms / steps * 1e-3) / 1e12
    res = {
        "value": pts / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms / steps, "gpu_launches": int(launches),
        "e2e": {"value": pts / (e_ms * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": int(pc_host.numel() * 4 + tg_host.numel() * 8 + W * N * 4 + W * 4), "d2h_bytes_per_step": 4},
        "roofline": {"bound": "tensor", "kernel": "whole step", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     "traffic": None, "peak_source": src, "model": "3 x 413 143 FLOP per point x 589 824 points / step time"},
        "config": {"workload": "train_w9: training step on a collated batch, 32 samples x 9 windows x 2048 points per GPU"},
        "notes": {"l2": "flushed between steps (256 MiB write)", "h2d": "the collated batch is copied from pinned host memory inside the timed step"},
        "dtype": "f32", "final_loss": float(keep["loss"]),
    }
    if with_cpu:
        from oracle import nn_bench as onb
        res["cpu_baseline"] = onb.cpu_train_w9()
    return res


def bench_fwd_bf16(dist, amp, steps, warmup, with_cpu):
    return bench_fwd(dist, amp, steps, warmup, with_cpu, precision="bf16")


# ------------------------------------------------------------------------------------------------
# configs[3]: arbitrary-scale inference of one 1M-point tile (k-means split into 2048-point blocks, blocks sharded)
# ------------------------------------------------------------------------------------------------
TILE_POINTS, TILE_SIDE, TILE_GRID, TILE_DIMS = 1_000_000, 1000.0, 10, 13


def synthetic_tile(seed=3000):
    """1 km^2 tile, 1M points: x, y in metres, z (HAG), 10 feature columns in [0, 1) (13-column rows as written by
    data_proc/2_preprocessing_filter_norm.py:76-104); cut into a 10 x 10 grid of 100 m windows (1_get_windows_split.py:57-80),
    every window filled with random duplicates to k * 2048 points (3_kmeans.py:54-69), coordinates normalised per window."""
    rng = np.random.default_rng(seed)
    pts = rng.random((TILE_POINTS, TILE_DIMS), dtype=np.float32)
    xy = pts[:, :2] * TILE_SIDE
    cell = (np.minimum((xy[:, 0] // (TILE_SIDE / TILE_GRID)).astype(np.int64), TILE_GRID - 1) * TILE_GRID +
            np.minimum((xy[:, 1] // (TILE_SIDE / TILE_GRID)).astype(np.int64), TILE_GRID - 1))
    order = np.argsort(cell, kind="stable")
    counts = np.bincount(cell, minlength=TILE_GRID * TILE_GRID)
    wins, ks, real = [], [], []
    o = 0
    for c in counts:
        w = pts[order[o:o + c]].copy()
        o += c
        side = TILE_SIDE / TILE_GRID
        w[:, 0] = (w[:, 0] * TILE_SIDE % side) / side
        w[:, 1] = (w[:, 1] * TILE_SIDE % side) / side
        w[:, 2] *= 0.3
        k = min(int(np.ceil(len(w) / NN_POINTS)), 9)
        need = k * NN_POINTS
        if len(w) < need:
            w = np.concatenate([w, w[rng.integers(0, len(w), need - len(w))]], 0)
        else:
            w = w[rng.permutation(len(w))[:need]]
        wins.append(w); ks.append(k); real.append(int(min(c, need)))
    return wins, ks, real


def bench_tile(dist, amp, steps, warmup, with_cpu, precision="fp32"):
    """One step = the whole tile: per rank, k-means block split of its windows (one launch for all windows), regroup,
    encoder over all blocks, attention + head per window (batched over windows of equal block count)."""
    dev = dist.device
    enc, seg = build_modules(amp, dev)
    enc.eval(); seg.eval()
    enc.precision = seg.precision = precision
    wins, ks, real = synthetic_tile()
    lo_w, hi_w = amp.shard_windows(ks, dist.world)[dist.rank]
    wins, ks, real = wins[lo_w:hi_w], ks[lo_w:hi_w], real[lo_w:hi_w]
    # the rank walks its windows in order of block count (windows are independent): the blocks of every block-count group are
    # then one contiguous range of the encoder output and the attention / head call reads them in place
    by = sorted(range(len(ks)), key=lambda i: ks[i])
    wins, ks, real = [wins[i] for i in by], [ks[i] for i in by], [real[i] for i in by]
    flush = _Flush(dev)
    out = {}
    if len(wins):
        host = torch.from_numpy(np.concatenate(wins, 0)).pin_memory()
        offsets = np.concatenate([[0], np.cumsum([len(w) for w in wins])]).astype(np.int64)
        pc = host.to(dev)
        by_k = {}
        for i, k in enumerate(ks):
            by_k.setdefault(k, []).append(i)
        # block / window index tensors of every block-count group: they depend on the window list only, not on the step
        blk0 = np.concatenate([[0], np.cumsum(ks)])
        rng_of = {k: (int(blk0[idx[0]]), int(blk0[idx[-1]] + k)) for k, idx in by_k.items()}      # block range of the group
        win_of = {k: (idx[0], idx[-1] + 1) for k, idx in by_k.items()}                               # window range of the group
    ev = {}

    def step(pc_dev):
        if not len(wins):
            return
        ev["a"] = torch.cuda.Event(enable_timing=True); ev["b"] = torch.cuda.Event(enable_timing=True)
        ev["a"].record()
        feats = amp.gather_feats(pc_dev, (0, 1, 9))                                   # 3_kmeans.py:81
        labels, _, _ = amp.kmeans_constrained_windows(feats, offsets, ks, NN_POINTS, NN_POINTS, check_range=False)   # columns in [-1, 1] / [0, 1]
        order, _, xy = amp.regroup_windows(labels, offsets, ks, pc_dev)
        ev["b"].record()
        grouped = pc_dev.index_select(0, order)                                       # rows sorted by (window, block)
        x9 = torch.cat((grouped[:, 0:3], grouped[:, 4:10]), 1)                        # datasets.py:342 keeps cols 0:3, 4:10
        x9[:, :2] = x9[:, :2] * 2 - 1                                                 # datasets.py:378-379
        blocks = x9.view(-1, NN_POINTS, NN_DIMS)
        enc_out, _ = enc(blocks)                                                      # encoder over all blocks of the rank
        preds = {}
        for k, idx in by_k.items():                                                   # windows of equal block count together
            b_lo, b_hi = rng_of[k]
            e = enc_out[b_lo:b_hi].view(len(idx), k, NN_POINTS, 320)                  # a view: no copy of the 320-wide rows
            gl = e[:, :, 0, :256].permute(1, 0, 2).contiguous()                       # [k, B, 256]
            lo = e.view(len(idx), k * NN_POINTS, 320)[:, :, 256:]                     # row-strided slice, read in place by seg
            cent = xy[win_of[k][0]:win_of[k][1], :k, :]
            logits, _ = seg(gl, lo, cent, [NN_POINTS] * k, None)
            preds[k] = logits.argmax(1)
        out["preds"] = preds

    ms = _timed(dist, lambda: step(pc if len(wins) else None), steps, warmup, flush)
    km_ms = ev["a"].elapsed_time(ev["b"]) if ev else 0.0

    def step_e2e():
        if not len(wins):
            return
        step(host.to(dev, non_blocking=True))
        for k, pr in out["preds"].items():
            out.setdefault("host", {})[k] = pr.to("cpu", non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e_ms = _timed(dist, step_e2e, max(2, steps // 2), 3, flush)
    es = max(2, steps // 2)
    n_pts = TILE_POINTS
    n_rows = int(sum(len(w) for w in wins))
    res = {
        "value": n_pts * steps / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms / steps, "scaling": "strong",
        "e2e": {"value": n_pts * es / (e_ms * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": n_rows * TILE_DIMS * 4, "d2h_bytes_per_step": n_rows * 8},
        "stages": {"kmeans_split_ms_rank0_last_step": km_ms, "blocks_this_rank": int(sum(ks)), "windows_this_rank": len(ks),
                   "rows_with_duplicate_fill_this_rank": n_rows},
        "config": {"workload": "configs[3]: 1M-point synthetic tile, 10 x 10 windows, constrained k-means into 2048-point blocks, "
                               "windows sharded over the GPUs, eval forward"},
        "notes": {"l2": "flushed between steps (256 MiB write)", "precision": precision, "n_gpus": dist.world},
        "dtype": "f32" if precision == "fp32" else "bf16",
    }
    pk, src = _peaks()
    fl = n_pts * FWD_FLOP_PER_POINT / (ms / steps * 1e-3) / 1e12
    res["roofline"] = {"bound": "tensor", "kernel": "whole step (k-means block split + encoder + head)", "achieved": fl, "peak": pk * dist.world,
                       "unit": "TFLOP/s", "frac": fl / (pk * dist.world), "traffic": None, "peak_source": src,
                       "model": "413 143 FLOP per real point x 1M points / step time (the k-means stage adds no FLOPs to the model)"}
    if with_cpu:
        from oracle import nn_bench as onb
        res["cpu_baseline"] = onb.cpu_tile(wins[:2], ks[:2])
    return res


def hooks():
    return {"fwd": bench_fwd, "fwd_bf16": bench_fwd_bf16, "train": bench_train, "train_w9": bench_train_w9, "tile": bench_tile}

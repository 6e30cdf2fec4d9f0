"""CPU baseline legs of bench.py for the PointNet-attention workloads -- TEST / BENCH INFRASTRUCTURE ONLY.

Times the oracle restatement (oracle/nn_oracle.py, pinned to the unmodified reference modules) of the
reference's CPU path on the box's host cores: the forward of pointNet/model/pointnetAtt.py:80-112,176-209 and
the training step of train_pointnet-attention.py:445-470 (autograd through the same ops + 2 x Adam), fp32,
all host threads. /root/reference does not exist on the GPU box, so the unmodified modules cannot be timed there.
"""
import os
import time

import numpy as np
import torch

from . import nn_oracle, nn_params

B, N = 32, 2048


def _inputs():
    xs, cent = nn_params.synthetic_blocks(B, N, 1, 2000)
    sd_e = nn_params.synthetic_state_dict(nn_params.encoder_shapes(), 0, trained_bn=False)
    sd_s = nn_params.synthetic_state_dict(nn_params.seg_shapes(), 1, trained_bn=False)
    return xs, cent, sd_e, sd_s


def cpu_forward(sample_steps=3):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    xs, cent, sd_e, sd_s = _inputs()
    with torch.no_grad():
        nn_oracle.forward_windows(sd_e, sd_s, xs, cent)
        ts = []
        for _ in range(sample_steps):
            t = time.perf_counter()
            nn_oracle.forward_windows(sd_e, sd_s, xs, cent)
            ts.append(time.perf_counter() - t)
    best = min(ts)
    return {"value": B * N / best, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": "%d forward passes of the full 32 x 2048 batch (oracle port of pointnetAtt.py on torch CPU fp32, %d threads), "
                      "best %.3f s" % (sample_steps, cores, best), "seconds": ts}


def cpu_train(sample_steps=2):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    xs, cent, sd_e, sd_s = _inputs()
    tg = torch.from_numpy(np.random.default_rng(0).integers(0, 5, (B, N)).astype(np.int64))
    leaves = []
    for sd in (sd_e, sd_s):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
                leaves.append(v)
    opt = torch.optim.Adam(leaves, lr=1e-3)
    ts = []
    for i in range(sample_steps + 1):
        t = time.perf_counter()
        opt.zero_grad()
        logits, ft = nn_oracle.forward_windows(sd_e, sd_s, xs, cent, training=True, stats_enc={}, stats_seg={})
        loss, _, _ = nn_oracle.train_step_loss(logits, tg, ft)
        loss.backward()
        opt.step()
        if i:
            ts.append(time.perf_counter() - t)
    best = min(ts)
    return {"value": B * N / best, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": "%d training steps of the full 32 x 2048 batch (oracle port on torch CPU fp32 autograd + Adam, %d threads), "
                      "best %.3f s" % (sample_steps, cores, best), "seconds": ts}


def cpu_train_w9(B=4, W=9):
    """CPU arm of train_w9 on a bounded sample: ONE training step of B samples x 9 windows x 2048 points (host augmentation as the
    reference loop does it, oracle forward / autograd backward, Adam), all host threads."""
    from . import train_loop_oracle as tlo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    pc = torch.from_numpy(rng.random((B, N, 9, W), dtype=np.float32))
    tg = torch.from_numpy(rng.integers(0, 5, (B, N, W)).astype(np.int64))
    sd_e = nn_params.synthetic_state_dict(nn_params.encoder_shapes(), 0, trained_bn=False)
    sd_s = nn_params.synthetic_state_dict(nn_params.seg_shapes(), 1, trained_bn=False)
    leaves = []
    for sd in (sd_e, sd_s):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True); leaves.append(v)
    opt = torch.optim.Adam(leaves, lr=1e-3)
    t0 = time.perf_counter()
    pcs, tgs = tlo.shuffle_clusters(pc, tg)
    ang = np.random.uniform() * 2 * np.pi
    xs, ts = [], []
    for w in range(W):
        a = pcs[:, :, :, w].numpy().copy()
        a[:, :, :3] = tlo.rotate_point_cloud_z(a[:, :, :3], rotation_angle=ang)
        a, t, _ = tlo.shuffle_data(a, tgs[:, :, w])
        xs.append(torch.Tensor(a)); ts.append(torch.LongTensor(t))
    cent = torch.stack([x[:, :, :2].mean(1) for x in xs], 1)
    opt.zero_grad()
    logits, ft = nn_oracle.forward_windows(sd_e, sd_s, xs, cent, training=True, stats_enc={}, stats_seg={})
    loss, _, _ = nn_oracle.train_step_loss(logits, torch.cat(ts, 1), ft)
    loss.backward()
    opt.step()
    dt = time.perf_counter() - t0
    return {"value": B * W * N / dt, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": "1 step of %d samples x %d windows x %d points (oracle port, %d threads), %.1f s" % (B, W, N, cores, dt)}


def cpu_tile(wins, ks):
    """CPU arm of the tile workload on a bounded sample: the oracle's constrained k-means (numpy restatement; the reference's
    third-party solver is not installable) + regroup + the oracle forward of the resulting blocks, for the given windows."""
    from . import kmeans_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd_e = nn_params.synthetic_state_dict(nn_params.encoder_shapes(), 0, trained_bn=False)
    sd_s = nn_params.synthetic_state_dict(nn_params.seg_shapes(), 1, trained_bn=False)
    t0 = time.perf_counter()
    n_pts = 0
    with torch.no_grad():
        for w, k in zip(wins, ks):
            lab, _, _ = kmeans_oracle.kmeans_constrained(np.ascontiguousarray(w[:, [0, 1, 9]]), int(k), 2048, 2048)
            order = np.argsort(lab, kind="stable")
            g = w[order]
            x9 = np.concatenate([g[:, 0:3], g[:, 4:10]], 1).reshape(int(k), 2048, 9).copy()
            x9[:, :, :2] = x9[:, :, :2] * 2 - 1
            xs = [torch.from_numpy(x9[i:i + 1]) for i in range(int(k))]
            cent = torch.stack([x[:, :, :2].mean(1) for x in xs], 1)
            nn_oracle.forward_windows(sd_e, sd_s, xs, cent)
            n_pts += len(w)
    dt = time.perf_counter() - t0
    return {"value": n_pts / dt, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": "%d windows (%d rows): oracle constrained k-means + forward of their blocks, %d threads, %.1f s" % (len(wins), n_pts, cores, dt)}


def reference_line(workload, args):
    if workload == "train_w9":
        r = cpu_train_w9()
        v = r["value"]
        return {"impl": "reference", "metric": "train pts/sec (9 windows per sample)", "value": v, "unit": "points/s", "n_gpus": args.gpus,
                "steps": 1, "warmup": 0, "ms_per_step": 1e3 * 32 * 9 * N / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "train_w9: training step on a collated batch, 32 samples x 9 windows x 2048 points per GPU"},
                "cpu_baseline": r, "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    fn = cpu_forward if workload == "fwd" else cpu_train
    for _ in range(min(args.warmup, 1)):
        fn(1)
    r = fn(max(1, args.steps))
    secs = r.pop("seconds")
    v = B * N * len(secs) / sum(secs)
    metric = "segmented points/sec (fwd)" if workload == "fwd" else "train pts/sec"
    cfg = ("configs[0]: segmentation forward, batch 32 x 2048 points, 9 channels, eval, random-init weights" if workload == "fwd"
           else "configs[2]: training step fwd+loss+bwd+2xAdam, batch 32 x 2048 points per GPU")
    r["value"] = v
    return {"impl": "reference", "metric": metric, "value": v, "unit": "points/s", "n_gpus": args.gpus, "steps": len(secs),
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": cfg}, "cpu_baseline": r,
            "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

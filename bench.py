"""bench.py -- throughput of the AMP-Net hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload fwd|train|fps|kmeans|tile|fwd_bf16] [--impl ours|reference]

One JSON line on stdout (rank 0). Under torchrun every rank processes its own shard (weak scaling,
no data-path collective for fps/fwd; NCCL gradient all-reduce for train; the tile is sharded: strong).

Workloads (BASELINE.json `configs`):
  fwd    configs[0]  PointNet-attention segmentation forward, 32 x 2048 points, fp32 parity path   unit: points/s
  train  configs[2]  fwd + loss + bwd + 2 x Adam, 32 x 2048 points                                 unit: points/s
  train_w9           the script's real step: 32 samples x 9 windows x 2048 points, device-side assembly  unit: points/s
  fps    configs[1]  FPS of 64 windows x 40 000 points -> 2048 per GPU                             unit: clouds/s
  kmeans             one k-means assignment pass over 16.8 M points, k = 9                         unit: points/s
  tile   configs[3]  1M-point tile: k-means block split + forward, windows sharded over the GPUs   unit: points/s
  fwd_bf16           configs[0] in bf16 (narrower than the reference: labelled, never the headline)
The default headline is `fwd` (the first metric BASELINE.json names, in the reference's fp32); every other workload is
measured in the same run and reported, compactly, under "workloads" of the same JSON line -- each with its own
`cpu_baseline` (oracle port on the box's host cores), `roofline` and `e2e`. The full records go to
gpurun_out/bench_detail.json when that directory exists.

Timing: CUDA events on the launching stream around every step, L2 flushed (256 MiB write) between
steps outside the events, barrier + synchronize on both sides of the region, max over ranks.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "3d-semantic-segmentation-amp-net_b200"

FPS_CLOUDS, FPS_POINTS, FPS_DIMS, FPS_SAMPLES = 64, 40000, 11, 2048
# fps_cluster_kernel as built: thread instructions per (candidate, pick) and DRAM bytes per 64-cloud launch, from ncu
FPS_WARP_INST_PER_UPDATE, FPS_DRAM_BYTES_PER_LAUNCH, FPS_NCU_SOURCE = 11.0, 119.6e6, "profiles/r02_fps_ncu_full.txt"
FPS_SMEM_SLOT_FRACTION = 28.0 / 40.0        # configs[1]: 40 slots per thread, 28 of them in shared memory (12 in registers)
NN_BATCH, NN_POINTS, NN_DIMS = 32, 2048, 9


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def clock_mhz():
    """Current SM clock (MHz) from nvidia-smi; the nominal maximum if it cannot be read."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True, timeout=10).stdout.split()
        return float(out[0])
    except Exception:
        return 1965.0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# distributed plumbing
# ------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self, gpus):
        import torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.torch = torch
        self.pg = False
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.pg = True
        else:
            torch.cuda.set_device(0)
        self.device = torch.device("cuda", self.local_rank if self.world > 1 else 0)

    def barrier(self):
        if self.pg:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if not self.pg:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        self.torch.distributed.all_reduce(t, op=self.torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.pg:
            self.torch.distributed.destroy_process_group()


class L2Flush:
    def __init__(self, torch, device, mib=256):
        self.buf = torch.empty(mib << 20, dtype=torch.uint8, device=device)

    def __call__(self):
        self.buf.fill_(1)


def timed_steps(dist, fn, steps, warmup, flush):
    """Returns (sum of per-step event times in ms (max over ranks), list of per-step ms on this rank)."""
    torch = dist.torch
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
    per = [a.elapsed_time(b) for a, b in evs]
    return dist.max_over_ranks(float(sum(per))), per


# ------------------------------------------------------------------------------------------------
# workload: FPS (configs[1])
# ------------------------------------------------------------------------------------------------
def fps_inputs(rank, n_clouds=FPS_CLOUDS):
    rng = np.random.default_rng(1000 + rank)
    return rng.random((n_clouds, FPS_POINTS, FPS_DIMS), dtype=np.float32)


def bench_fps(dist, amp, steps, warmup, with_cpu):
    torch = dist.torch
    host = torch.from_numpy(fps_inputs(dist.rank)).pin_memory()
    dev = host.to(dist.device)
    flush = L2Flush(torch, dist.device)
    out = {}

    def step_resident():
        idx = amp.fps_indices(dev, FPS_SAMPLES, check_finite=False)
        out["rows"] = amp.gather_rows(dev, idx)
        out["idx"] = idx

    n0 = amp._lib.launch_count()
    total_ms, per = timed_steps(dist, step_resident, steps, warmup, flush)
    launches = (amp._lib.launch_count() - n0) // (steps + warmup) * steps
    # the FPS kernel alone (dominant kernel), timed with events on the launching stream
    def step_kernel():
        out["idx"] = amp.fps_indices(dev, FPS_SAMPLES, check_finite=False)
    k_ms, _ = timed_steps(dist, step_kernel, steps, warmup, flush)
    # end to end through the reference-shaped host API: pinned host rows in, sampled rows out on the host
    rows_host = torch.empty((FPS_CLOUDS, FPS_SAMPLES, FPS_DIMS), dtype=torch.float32).pin_memory()

    def step_e2e():
        amp.fps_host_batch(host, FPS_SAMPLES, out=rows_host)

    serial_ms, _ = timed_steps(dist, step_e2e, steps, warmup, flush)
    # ... and as the sample_fps.py loop over many windows runs it: amp.FpsHostStream keeps 3 batches in flight, so the 2 ms
    # host-to-device copy of the next batch hides behind the 3.3 ms kernel of the current one. Every step still copies ITS
    # 30.7 MB in and ITS rows out inside the timed region; the host batches rotate over a pinned pool larger than L2
    # (5 x 30.7 MB), no flush kernel (it would sit on the run stream between the launches).
    pool = [host] + [torch.from_numpy(fps_inputs(100 + dist.rank * 8 + i)).pin_memory() for i in range(4)]
    stream = amp.FpsHostStream(host, FPS_SAMPLES, depth=3, device=dist.device)
    for _ in stream.run(pool[i % 5] for i in range(max(warmup, 3))):
        pass
    torch.cuda.synchronize(dist.device)
    dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream.pipe.s_in)
    for _ in stream.run(pool[(warmup + i) % 5] for i in range(steps)):
        pass
    ev1.record(stream.pipe.s_out)
    ev1.synchronize()
    dist.barrier()
    e_ms = dist.max_over_ranks(float(ev0.elapsed_time(ev1)))
    clouds = FPS_CLOUDS * dist.world * steps
    # Roofline of fps_cluster_kernel: the whole cloud lives on chip (registers + shared memory; ncu: dram bytes = 0.2 % of the
    # HBM model's), so HBM bounds nothing. The binding resource is the shared-memory read port: every pick reads the 12-byte
    # coordinates of every shared-memory-resident candidate once (ncu: the LSU shared pipe is 57 % busy over the active
    # cycles, the issue slots 50 %; the packed-FFMA2 pipe needs about as many cycles per pick). achieved = algorithmic shared-memory bytes / kernel time, peak = 128 B per clock per SM.
    sm_mhz = clock_mhz()
    smem_bytes = 12.0 * FPS_SMEM_SLOT_FRACTION * (FPS_SAMPLES - 1) * FPS_POINTS * FPS_CLOUDS
    ach = smem_bytes / (k_ms / steps * 1e-3) / 1e9
    peak = 128.0 * 148 * sm_mhz * 1e-3
    warp_inst = FPS_WARP_INST_PER_UPDATE * (FPS_SAMPLES - 1) * FPS_POINTS * FPS_CLOUDS / 32.0
    issue_frac = warp_inst / (k_ms / steps * 1e-3) / 1e9 / (4.0 * 148 * sm_mhz * 1e-3)
    res = {
        "value": clouds / (total_ms * 1e-3), "unit": "clouds/s", "ms_per_step": total_ms / steps,
        "gpu_launches": int(launches),
        "e2e": {"value": clouds / (e_ms * 1e-3), "unit": "clouds/s",
                "h2d_bytes_per_step": int(host.numel() * 4), "d2h_bytes_per_step": int(rows_host.numel() * 4)},
        "roofline": {"bound": "smem", "kernel": "fps_cluster_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak, "traffic": FPS_DRAM_BYTES_PER_LAUNCH,
                     "model": "12 B of coordinates per shared-memory-resident candidate (%.0f %% of the slots) per pick x (S-1) x P x "
                              "clouds / kernel time; peak = 128 B/clk x 148 SMs x %.0f MHz (the launch occupies 128 SMs: 64 clouds x "
                              "2 CTAs). Issue slots: %.1f thread instructions per candidate update (ncu, %s) = %.2f of 4 per SM per "
                              "clock; traffic = DRAM bytes per launch (ncu): HBM is idle, the cloud is on chip"
                              % (100 * FPS_SMEM_SLOT_FRACTION, sm_mhz, FPS_WARP_INST_PER_UPDATE, FPS_NCU_SOURCE, issue_frac),
                     "kernel_ms": k_ms / steps},
        "config": {"workload": "configs[1]: FPS %d windows x %d pts -> %d per GPU, float32 rows of %d columns"
                               % (FPS_CLOUDS, FPS_POINTS, FPS_SAMPLES, FPS_DIMS)},
        "notes": {"l2": "flushed between steps (256 MiB write)", "clouds_per_gpu": FPS_CLOUDS,
                  "e2e": "FpsHostStream, 3 batches in flight, per-step H2D + D2H timed, 154 MB pinned input pool (> L2), no flush; "
                         "serial fps_host_batch loop (L2 flushed): %.0f clouds/s" % (clouds / (serial_ms * 1e-3))},
        "dtype": "f32",
    }
    if with_cpu:
        res["cpu_baseline"] = cpu_fps(sample_clouds=min(max(4 * (os.cpu_count() or 1), 16), 128))   # ~10-20 s of CPU work
    return res


def cpu_fps(sample_clouds, repeat=1):
    """The oracle's C restatement of utils/utils.py:889-933, one cloud per host thread (ctypes drops the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import fps_oracle
    cores = os.cpu_count() or 1
    pcs = fps_inputs(0, sample_clouds)
    fps_oracle.fps_indices_c(pcs[0][:2000], 64)
    best = None
    for _ in range(repeat):
        t = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(lambda p: fps_oracle.fps_indices_c(p, FPS_SAMPLES), pcs))
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return {"value": sample_clouds / best, "unit": "clouds/s", "cores": cores, "kind": "port",
            "sample": "%d of the %d clouds (40 000 -> 2048), C port of the reference fps, %d threads, %.1f s"
                      % (sample_clouds, FPS_CLOUDS, cores, best)}


# ------------------------------------------------------------------------------------------------
# workload: k-means assignment step (the distance + argmin stream of 3_kmeans.py:82 / utils.py:505)
# ------------------------------------------------------------------------------------------------
KM_POINTS, KM_K = 1 << 24, 9


def bench_kmeans(dist, amp, steps, warmup, with_cpu):
    """One step = one Lloyd assignment pass over 16.8 M points (3 clustering features) against k = 9 centroids:
    12 B read + 4 B label written per point, centroids on chip. The 201 MB of features exceed the 126 MB L2."""
    torch = dist.torch
    g = torch.Generator(device="cpu").manual_seed(4000 + dist.rank)
    host = torch.rand((KM_POINTS, 3), generator=g, dtype=torch.float32).pin_memory()
    cent = torch.rand((KM_K, 3), generator=g, dtype=torch.float32).to(dist.device)
    dev = host.to(dist.device)
    flush = L2Flush(torch, dist.device)
    out = {}

    def step():
        out["labels"] = amp.kmeans_assign(dev, cent)

    n0 = amp._lib.launch_count()
    ms, _ = timed_steps(dist, step, steps, warmup, flush)
    launches = (amp._lib.launch_count() - n0) // (steps + warmup) * steps
    lab_host = torch.empty((KM_POINTS,), dtype=torch.int32).pin_memory()

    def step_e2e():
        lab_host.copy_(amp.kmeans_assign(host.to(dist.device, non_blocking=True), cent), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e_ms, _ = timed_steps(dist, step_e2e, max(2, steps // 2), 1, flush)
    es = max(2, steps // 2)
    pk = peaks()
    ach = KM_POINTS * 16.0 / (ms / steps * 1e-3) / 1e9
    res = {
        "value": KM_POINTS * dist.world * steps / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms / steps,
        "gpu_launches": int(launches),
        "e2e": {"value": KM_POINTS * dist.world * es / (e_ms * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": KM_POINTS * 12, "d2h_bytes_per_step": KM_POINTS * 4},
        "roofline": {"bound": "hbm", "kernel": "kmeans_assign_small_k_kernel<9>", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": ach / pk["hbm_gbs"],
                     "traffic": 250.1e6,   # dram__bytes_read + write of one launch, ncu --set full (profiles/r01_final_kmeans_ncu.txt)
                     "peak_source": pk["source"] + " (burst copy)",
                     "model": "16 B per point per assignment pass (12 B features read + 4 B int32 label written), k = %d" % KM_K},
        "config": {"workload": "k-means assignment pass, %d points x 3 features, k = %d" % (KM_POINTS, KM_K)},
        "notes": {"l2": "flushed between steps (256 MiB write); working set 268 MB > L2"},
        "dtype": "f32",
    }
    if with_cpu:
        from oracle import kmeans_oracle
        n = 1 << 21
        x = host[:n].numpy(); c = cent.cpu().numpy()
        t = time.perf_counter()
        kmeans_oracle.assign(x, c)
        dt = time.perf_counter() - t
        res["cpu_baseline"] = {"value": n / dt, "unit": "points/s", "cores": 1, "kind": "port",
                               "sample": "%d points, numpy restatement of the assignment step, %.2f s" % (n, dt)}
    return res


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path (oracle port; the reference is
# pure Python and /root/reference does not exist on the GPU box)
# ------------------------------------------------------------------------------------------------
def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    wl = args.workload
    if wl == "tile":
        wl = "fwd"                  # the tile's CPU arm is the forward of its blocks (the constrained solver has no CPU reference)
    if wl == "fps":
        cores = os.cpu_count() or 1
        sample = min(max(4 * cores, 16), 128)         # ~1 s of wall clock per step on the box's host cores
        t_all = []
        for _ in range(args.warmup and 1):
            cpu_fps(min(sample, cores))
        for _ in range(args.steps):
            r = cpu_fps(sample)
            t_all.append(sample / r["value"])
        v = sample * len(t_all) / sum(t_all)
        line = {"impl": "reference", "metric": "FPS clouds/sec", "value": v, "unit": "clouds/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sum(t_all) / len(t_all) * FPS_CLOUDS / sample,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1]: FPS %d windows x %d pts -> %d per GPU, float32 rows of %d columns"
                                       % (FPS_CLOUDS, FPS_POINTS, FPS_SAMPLES, FPS_DIMS)},
                "cpu_baseline": {"value": v, "unit": "clouds/s", "cores": cores, "kind": "port",
                                 "sample": "each step = %d clouds on %d threads (C port of utils/utils.py:889-933)"
                                           % (sample, cores)},
                "e2e": {"value": v, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    elif wl == "kmeans":
        from oracle import kmeans_oracle
        rng = np.random.default_rng(4000)
        n = 1 << 21
        x = rng.random((n, 3), dtype=np.float32); c = rng.random((KM_K, 3), dtype=np.float32)
        ts = []
        for i in range(args.warmup + max(1, args.steps)):
            t = time.perf_counter(); kmeans_oracle.assign(x, c); dt = time.perf_counter() - t
            if i >= args.warmup:
                ts.append(dt)
        v = n * len(ts) / sum(ts)
        line = {"impl": "reference", "metric": "k-means assigned points/sec", "value": v, "unit": "points/s", "n_gpus": args.gpus,
                "steps": len(ts), "warmup": args.warmup, "ms_per_step": 1e3 * sum(ts) / len(ts) * KM_POINTS / n, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "k-means assignment pass, %d points x 3 features, k = %d" % (KM_POINTS, KM_K)},
                "cpu_baseline": {"value": v, "unit": "points/s", "cores": 1, "kind": "port",
                                 "sample": "each step = %d points (numpy restatement of the assignment step; the reference's "
                                           "third-party solver is not installable here)" % n},
                "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    else:
        from oracle import nn_bench
        line = nn_bench.reference_line("fwd" if wl.startswith("fwd") else (wl if wl in ("train", "train_w9") else "fwd"), args)
    emit(line)


METRICS = {"fps": "FPS clouds/sec", "fwd": "segmented points/sec (fwd)", "fwd_bf16": "segmented points/sec (fwd), bf16",
           "kmeans": "k-means assigned points/sec", "train": "train pts/sec", "train_w9": "train pts/sec (9 windows per sample)",
           "tile": "segmented points/sec (1M-point tile)"}
ORDER = ["fwd", "train", "train_w9", "fps", "kmeans", "tile", "fwd_bf16"]


def _r(x, nd=4):
    """Round to nd significant digits (compact JSON)."""
    if not isinstance(x, float) or x == 0.0 or x != x:
        return x
    from math import floor, log10
    return round(x, nd - 1 - int(floor(log10(abs(x)))))


def compact(r):
    """The secondary-workload record: numbers only, short keys (the whole line must fit the driver's stdout tail)."""
    if "error" in r:
        return r
    o = {"value": _r(r["value"]), "unit": r["unit"], "ms": _r(r["ms_per_step"]), "dtype": r.get("dtype")}
    if "scaling" in r:
        o["scaling"] = r["scaling"]
    if "e2e" in r:
        o["e2e"] = _r(r["e2e"]["value"])
    rf = r.get("roofline")
    if rf:
        o["roofline"] = {"bound": rf["bound"], "achieved": _r(rf["achieved"]), "peak": _r(rf["peak"]), "unit": rf["unit"], "frac": _r(rf["frac"], 3)}
    cb = r.get("cpu_baseline")
    if cb:
        o["cpu"] = {"value": _r(cb["value"]), "cores": cb["cores"], "kind": cb["kind"]}
    if "stages" in r:
        o["stages"] = {k: _r(v) if isinstance(v, float) else v for k, v in r["stages"].items()}
    return o


def main():
    # stdout carries exactly ONE JSON line: anything a library prints there (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fwd", choices=ORDER)
    ap.add_argument("--only", action="store_true", help="measure only the headline workload")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args, emit)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    amp = importlib.import_module(PKG)
    amp._lib.lib()
    import bench_nn
    dist = Dist(args.gpus)
    with_cpu = (dist.rank == 0 and dist.world == 1 and not args.no_cpu)
    table = {"fps": bench_fps, "kmeans": bench_kmeans}
    table.update(bench_nn.hooks())
    with ClockSampler(dist.local_rank) as clk:
        head = table[args.workload](dist, amp, args.steps, args.warmup, with_cpu)
    detail = {args.workload: dict(head)}
    line = {"metric": METRICS[args.workload],
            "value": head.pop("value"), "unit": head.pop("unit"), "n_gpus": dist.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head.pop("ms_per_step"), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic"}
    line.update(head)          # a workload may override "scaling" (the tile workload shards one tile: strong)
    line["clocks"] = clk.summary()
    if not args.only:
        others = {}
        for name in ORDER:
            if name == args.workload:
                continue
            try:
                # every workload starts from an empty caching allocator: blocks cached by the previous one (hundreds of MB of
                # training workspace) otherwise make the next one's first allocations fall back to cudaFree + cudaMalloc
                # inside its timed steps
                import gc
                gc.collect()
                torch.cuda.synchronize()
                torch.cuda.empty_cache()
                r = table[name](dist, amp, max(3, args.steps // 2), 3, with_cpu)
            except Exception as e:  # a secondary workload must not take the headline down
                r = {"error": repr(e)[:160]}
            detail[name] = r
            others[name] = compact(r)
        line["workloads"] = others
    if dist.rank == 0:
        emit(line)
        try:
            if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
                with open(os.path.join(ROOT, "gpurun_out", "bench_detail_n%d.json" % dist.world), "w") as f:
                    json.dump({"line": line, "detail": detail}, f, indent=1, default=str)
        except OSError:
            pass
    dist.close()


if __name__ == "__main__":
    main()

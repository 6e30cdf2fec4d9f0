"""Farthest-point sampling: host-side mirror of the reference interface.

`fps(pc, n_samples)` keeps the signature and return value of the reference
utils/utils.py:889-933 (rows of `pc` in pick order) and is what data_proc/sample_fps.py:24,29
and data_proc/3_kmeans.py:111 call. The work is done by csrc/fps.cu through the C ABI
(amp_fps_f32 / amp_fps_f64 / amp_gather_rows); there is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib


def _device(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("ampnet_b200: CUDA device required (no CPU fallback)")
    return torch.device(device if device is not None else "cuda")


def fps_indices(pc, n_samples, start_idx=0, check_finite=True):
    """pc: CUDA tensor [B, P, D] or [P, D] (float32 / float64, contiguous, D >= 3).
    Returns int64 indices [B, S] (or [S]) in pick order, on the device, enqueued on the
    current stream."""
    squeeze = pc.dim() == 2
    if squeeze:
        pc = pc.unsqueeze(0)
    if pc.dim() != 3:
        raise ValueError("pc must be [B, P, D] or [P, D]")
    if pc.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("ampnet_b200: fps needs float32 or float64 coordinates, got %s" % pc.dtype)
    _lib.require_cuda(pc, "pc")
    B, P, D = pc.shape
    if D < 3:
        raise ValueError("pc needs at least 3 columns")
    if n_samples > P:
        # the reference raises ValueError here too (np.argmax of an empty array, utils.py:927)
        raise ValueError("n_samples (%d) > number of points (%d)" % (n_samples, P))
    out, status = _fps_launch(pc, n_samples, start_idx)
    if check_finite and bool(status.any().item()):
        raise ValueError("fps: non-finite coordinates in input")
    return out[0] if squeeze else out


def _fps_launch(pc, n_samples, start_idx=0):
    """Enqueue amp_fps_* for a validated [B, P, D] CUDA tensor; returns (indices [B, S], status [B]) without synchronising."""
    B, P, D = pc.shape
    lib = _lib.lib()
    elem = pc.element_size()
    out = torch.empty((B, n_samples), dtype=torch.int64, device=pc.device)
    status = torch.empty((B,), dtype=torch.int32, device=pc.device)
    ws_bytes = lib.amp_fps_workspace_bytes(B, P, elem)
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=pc.device)
    fn = lib.amp_fps_f32 if elem == 4 else lib.amp_fps_f64
    with torch.cuda.device(pc.device):
        _lib.check(fn(pc.data_ptr(), B, P, D, int(n_samples), int(start_idx), out.data_ptr(),
                      status.data_ptr(), ws.data_ptr() if ws_bytes else None, ws_bytes, _lib.stream_ptr()))
    return out, status


def gather_rows(pc, idx):
    """pc [B, P, D] CUDA, idx [B, S] int64 CUDA -> [B, S, D] (the `pc[sample_inds]` of utils.py:933)."""
    _lib.require_cuda(pc, "pc")
    _lib.require_cuda(idx, "idx", torch.int64)
    B, P, D = pc.shape
    S = idx.shape[1]
    out = torch.empty((B, S, D), dtype=pc.dtype, device=pc.device)
    with torch.cuda.device(pc.device):
        _lib.check(_lib.lib().amp_gather_rows(pc.data_ptr(), B, P, D, pc.element_size(), idx.data_ptr(), S,
                                              out.data_ptr(), _lib.stream_ptr()))
    return out


def fps_batch(pc, n_samples, start_idx=0):
    """Device-resident batch form: pc [B, P, D] CUDA -> (rows [B, S, D], idx [B, S])."""
    idx = fps_indices(pc, n_samples, start_idx)
    return gather_rows(pc, idx), idx


def fps_host_batch(pc_host, n_samples, out=None, device=None):
    """Host-buffer batch form of the sample_fps.py loop (data_proc/sample_fps.py:12-34): pc_host is a CPU
    tensor / ndarray [B, P, D] float32 (pinned memory makes the copies asynchronous); returns the sampled
    rows [B, S, D] on the host (written into `out` when given). One H2D copy, one FPS launch, one gather,
    one D2H copy, one stream synchronise."""
    t = torch.from_numpy(pc_host) if isinstance(pc_host, np.ndarray) else pc_host
    if t.dim() != 3 or t.dtype not in (torch.float32, torch.float64):
        raise ValueError("pc_host must be [B, P, D] float32/float64")
    dev = _device(device)
    d = t.to(dev, non_blocking=True)
    idx = fps_indices(d, n_samples, check_finite=False)
    rows = gather_rows(d, idx)
    if out is None:
        out = torch.empty(rows.shape, dtype=rows.dtype, pin_memory=True)
    out.copy_(rows, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return out


class FpsHostStream:
    """The sample_fps.py loop over MANY batches (data_proc/sample_fps.py:12-34 walks a directory of windows) with the copies
    hidden: `run(batches)` takes an iterable of equally shaped host tensors / arrays [B, P, D] (pinned memory, or the copies
    serialise) and yields the sampled rows [B, S, D] of each batch as pinned host tensors, in order. `depth` batches are in
    flight (StreamedForward: copy-in / run / copy-out streams, one captured launch sequence per slot), so the host-to-device
    copy of the next batch -- 2 ms for 64 x 40 000 x 3 floats, against 3.3 ms of sampling -- hides behind the kernel of the
    current one. A yielded tensor is valid until `depth` more were requested. Non-finite coordinates raise ValueError when
    the offending batch's result arrives."""

    def __init__(self, example, n_samples, depth=3, device=None):
        from .streamed import StreamedForward
        self.n_samples = int(n_samples)
        example = self._as_tensor(example)

        def run(pc):
            idx, status = _fps_launch(pc, self.n_samples, 0)
            return gather_rows(pc, idx), status

        self.pipe = StreamedForward(run, (example,), device=_device(device), depth=depth, warmup=1)

    def _as_tensor(self, b):
        t = torch.from_numpy(b) if isinstance(b, np.ndarray) else b
        if t.dim() != 3 or t.dtype not in (torch.float32, torch.float64) or t.shape[2] < 3:
            raise ValueError("batches must be [B, P, D >= 3] float32/float64")
        if self.n_samples > t.shape[1]:
            raise ValueError("n_samples (%d) > number of points (%d)" % (self.n_samples, t.shape[1]))
        return t

    def run(self, batches):
        for rows, status in self.pipe.run((self._as_tensor(b),) for b in batches):
            if bool(status.any()):
                raise ValueError("fps: non-finite coordinates in input")
            yield rows


def fps_host_stream(batches, n_samples, depth=3, device=None):
    """FpsHostStream for a one-off iterable: builds the pipeline from the first batch and runs all of them through it."""
    import itertools
    it = iter(batches)
    first = next(it, None)
    if first is None:
        return
    yield from FpsHostStream(first, n_samples, depth=depth, device=device).run(itertools.chain([first], it))


def fps(pc, n_samples, device=None):
    """Drop-in for the reference `fps(pc, n_samples)` (utils/utils.py:889).

    pc: array-like [N, D] (NumPy array as in the reference, or a torch tensor).
    Returns the sampled ROWS `pc[sample_inds]` in pick order, same type/dtype as the input."""
    if isinstance(pc, torch.Tensor):
        t = pc
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        if not t.is_cuda:
            t = t.to(_device(device))
        rows, _ = fps_batch(t.contiguous().unsqueeze(0), n_samples)
        rows = rows[0]
        return rows.to(pc.device).to(pc.dtype) if rows.device != pc.device or rows.dtype != pc.dtype else rows
    arr = np.asarray(pc)
    work = arr if arr.dtype in (np.float32, np.float64) else arr.astype(np.float64)
    t = torch.from_numpy(np.ascontiguousarray(work)).to(_device(device))
    idx = fps_indices(t, n_samples)
    return arr[idx.cpu().numpy()]

"""Data preparation on the device: host-side mirror of the reference's tile -> windows -> normalised rows steps.

  split_windows()             <- data_proc/1_get_windows_split.py:53-80   (the numerical core; LAS I/O stays on the host)
  filter_normalize_windows()  <- data_proc/2_preprocessing_filter_norm.py:40-123

Everything is float64 like the reference's numpy code and bit-exact to it (oracle/dataprep_oracle.py, pinned to the
unmodified functions). CUDA only: CPU tensors raise.
"""
import numpy as np
import torch

from . import _lib


def _bytes(n, device):
    return torch.empty((max(int(n), 1),), dtype=torch.uint8, device=device)


def split_windows(x, y, w_size=(40, 40)):
    """x, y: CUDA float64 [P] (may be strided views of one [P, C] tensor with equal strides).
    Returns dict(ids int32 [P] (-1 = no window), order int64 [P], offsets int64 [nx*ny + 1] (device), nx, ny, x0, y0):
    window w (y-major: w = iy * nx + ix) holds the points order[offsets[w]:offsets[w+1]] in their original order."""
    if not (x.is_cuda and y.is_cuda) or x.dtype != torch.float64 or y.dtype != torch.float64 or x.shape != y.shape or x.dim() != 1:
        raise RuntimeError("ampnet_b200: x and y must be CUDA float64 vectors of equal length (no CPU fallback)")
    if x.stride(0) != y.stride(0):
        x, y = x.contiguous(), y.contiguous()
    P, stride, dev = x.shape[0], x.stride(0), x.device
    lib = _lib.lib()
    mm = torch.empty(4, dtype=torch.float64, device=dev)
    ws = _bytes(64, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.amp_minmax_f64(x.data_ptr(), y.data_ptr(), P, stride, mm.data_ptr(), ws.data_ptr(), 64, _lib.stream_ptr()))
    xmin, xmax, ymin, ymax = (float(v) for v in mm.cpu())
    x0, y0 = round(xmin), round(ymin)                                   # Python round: half to even, like the reference (:57, :60)
    nx, ny = len(range(x0, round(xmax), int(w_size[0]))), len(range(y0, round(ymax), int(w_size[1])))
    ids = torch.empty(P, dtype=torch.int32, device=dev)
    order = torch.empty(P, dtype=torch.int64, device=dev)
    n_bins = max(nx * ny, 1)
    offsets = torch.empty(n_bins + 1, dtype=torch.int64, device=dev)
    ws_bytes = lib.amp_window_partition_workspace_bytes(P)
    ws = _bytes(ws_bytes, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.amp_window_ids_f64(x.data_ptr(), y.data_ptr(), P, stride, float(x0), float(y0), int(w_size[0]), int(w_size[1]),
                                          nx, ny, ids.data_ptr(), _lib.stream_ptr()))
        _lib.check(lib.amp_window_partition(ids.data_ptr(), P, n_bins, order.data_ptr(), offsets.data_ptr(), ws.data_ptr(), ws_bytes,
                                            _lib.stream_ptr()))
    return {"ids": ids, "order": order, "offsets": offsets, "nx": nx, "ny": ny, "x0": x0, "y0": y0}


def filter_normalize_windows(cols, order, offsets, max_z=100.0, max_intensity=5000, n_points=1024):
    """cols: CUDA float64 [P, 10] = (x, y, z, HeightAboveGround, class, intensity, red, green, blue, nir); order / offsets as
    returned by split_windows. Returns (rows float64 [n, 13] on the device, out_offsets int64 [W + 1] on the host,
    stored bool [W]): window w's normalised rows are rows[out_offsets[w]:out_offsets[w+1]]; stored[w] is False where the
    reference writes no file (nothing kept, zero extent, or fewer than n_points rows: 2_preprocessing_filter_norm.py:56,92,107)."""
    _lib.require_cuda(cols, "cols", torch.float64)
    if cols.dim() != 2 or cols.shape[1] != 10:
        raise ValueError("cols must be [P, 10]")
    dev = cols.device
    W = offsets.shape[0] - 1
    lib = _lib.lib()
    out = torch.empty((cols.shape[0], 13), dtype=torch.float64, device=dev)
    out_off = torch.empty(W + 1, dtype=torch.int64, device=dev)
    ws_bytes = lib.amp_filter_normalize_workspace_bytes(W)
    ws = _bytes(ws_bytes, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.amp_filter_normalize_f64(cols.data_ptr(), order.data_ptr(), offsets.data_ptr(), W, float(max_z), float(max_intensity),
                                                out.data_ptr(), out_off.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr()))
    oo = out_off.cpu().numpy()
    stored = np.diff(oo) >= max(int(n_points), 1)
    return out[:int(oo[-1])], oo, stored

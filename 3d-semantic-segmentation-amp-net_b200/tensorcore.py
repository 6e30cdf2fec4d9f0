"""Point-wise linear layer on the tcgen05 tensor cores (amp_tc_linear_bf16 of include/ampnet_b200.h).

The unit the fused bf16 chains of the encoder / segmentation head are built from; exposed for tests and for callers that
want a single Conv1d(kernel 1) / Linear of pointNet/model/pointnetAtt.py on the tensor pipe. CUDA only."""
import torch

from . import _lib


def tc_linear(x, weight, bias=None, relu=False, pool=False):
    """x [clouds, rows, K] f32, weight [N, K] f32 -> y [clouds, rows, N] f32 (bf16 operands, fp32 accumulate).
    pool=True returns max over rows of relu(y) per cloud [clouds, N] instead (needs relu, N % 128 == 0)."""
    lib = _lib.lib()
    x = _lib.require_cuda(x, "x", torch.float32)
    w = _lib.require_cuda(weight, "weight", torch.float32)
    if x.dim() != 3 or w.dim() != 2 or x.shape[2] != w.shape[1]:
        raise ValueError("x must be [clouds, rows, K] and weight [N, K]")
    b = _lib.require_cuda(bias, "bias", torch.float32) if bias is not None else None
    C, R, K = x.shape
    N = w.shape[0]
    dev = x.device
    ws_bytes = lib.amp_tc_linear_workspace_bytes(K, N)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    y = None if pool else torch.empty((C, R, N), dtype=torch.float32, device=dev)
    pm = torch.zeros((C, N), dtype=torch.int32, device=dev) if pool else None
    with torch.cuda.device(dev):
        _lib.check(lib.amp_tc_linear_bf16(x.data_ptr(), C, R, K, w.data_ptr(), b.data_ptr() if b is not None else None, N,
                                          1 if relu else 0, y.data_ptr() if y is not None else None,
                                          pm.data_ptr() if pm is not None else None, ws.data_ptr(), ws_bytes, _lib.stream_ptr()))
    return pm.view(torch.float32) if pool else y


def linear_wgrad(dy, a, bias=False):
    """dy [clouds, rows, N] f32, a [clouds, rows, K] f32 -> dw [N, K] (and db [N] when bias=True): the weight gradient of a
    point-wise linear layer summed over all rows (amp_wgrad_f32; tensor cores when the shape allows)."""
    lib = _lib.lib()
    dy = _lib.require_cuda(dy, "dy", torch.float32)
    a = _lib.require_cuda(a, "a", torch.float32)
    if dy.dim() != 3 or a.dim() != 3 or dy.shape[:2] != a.shape[:2]:
        raise ValueError("dy must be [clouds, rows, N] and a [clouds, rows, K]")
    C, R, N = dy.shape
    K = a.shape[2]
    dev = dy.device
    dw = torch.empty((N, K), dtype=torch.float32, device=dev)
    db = torch.empty((N,), dtype=torch.float32, device=dev) if bias else None
    ws_bytes = lib.amp_wgrad_workspace_bytes(C, R, N, K)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.amp_wgrad_f32(dy.data_ptr(), a.data_ptr(), C, R, N, K, dw.data_ptr(), db.data_ptr() if bias else None,
                                     ws.data_ptr(), ws_bytes, _lib.stream_ptr()))
    return (dw, db) if bias else dw

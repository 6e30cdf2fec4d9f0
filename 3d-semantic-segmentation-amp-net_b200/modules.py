"""Drop-in torch.nn.Modules for the reference model classes of pointNet/model/pointnetAtt.py.

    TransformationNet            <- pointnetAtt.py:7-47
    BasePointNet                 <- pointnetAtt.py:50-112
    SegmentationWithAttention    <- pointnetAtt.py:154-209

Constructor signatures, parameter / buffer names, shapes and construction order (hence default random
init under a given torch.manual_seed, and checkpoint compatibility by key: utils/utils.py:424-427,
train_pointnet-attention.py:155-156, test_pointnet_att_segmen.py:87-88) are the reference's. The nn.Conv1d /
nn.BatchNorm1d / nn.Linear / nn.MultiheadAttention submodules are PARAMETER CONTAINERS only: forward and
backward run hand-written CUDA through the C ABI (amp_encoder_fwd/bwd, amp_seg_fwd/bwd in
include/ampnet_b200.h). There is no CPU or PyTorch fallback: CPU tensors raise.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib


PRECISIONS = {"fp32": 0, "bf16": 1, "fp32_strict": 2}          # AMP_PREC_* of include/ampnet_b200.h
_default_precision = "fp32"


def set_default_precision(name):
    """Arithmetic of the forward for modules that do not set `.precision` themselves:
    "fp32" (the parity path: split-bf16 tensor-core GEMMs at fp32-class accuracy + fp32 CUDA-core layers), "bf16"
    (eval only: fused tcgen05 chains, bf16 operands; training then runs "fp32") or "fp32_strict" (train + eval: plain fp32
    FMA kernels only -- slower, for gradient comparisons below the 1e-3 level)."""
    global _default_precision
    if name not in PRECISIONS:
        raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
    _default_precision = name


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def _bytes(n, device):
    return torch.empty((max(int(n), 1),), dtype=torch.uint8, device=device)


class _NativeModule(nn.Module):
    """Caches the state_dict-ordered tensor list the C ABI takes and checks it against the library's names."""
    _abi_count = None
    _abi_name = None
    precision = None        # None: follow set_default_precision(); "fp32" | "bf16" (eval only; training is fp32)

    def _precision(self):
        name = self.precision or _default_precision
        if name not in PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
        if self.training and name == "bf16":
            return 0
        return PRECISIONS[name]

    def _live_ids(self):
        """ids of every parameter / buffer object currently registered below this module (cheap: ~25 dict walks)."""
        subs = self.__dict__.get("_amp_submodules")
        if subs is None:
            subs = list(self.modules())
            self.__dict__["_amp_submodules"] = subs
        return tuple(id(t) for m in subs for d in (m._parameters, m._buffers) for t in d.values())

    def _pack_cache(self, nbytes, device):
        """(buffer, valid) of the packed-weight cache of the fused eval path (amp_*_fwd `pack_cache`): rebuilt by the library
        when a parameter or BatchNorm buffer was replaced or written in place since the last eval call (`_version` counters)."""
        ts = self._native_tensors()
        key = (tuple((t.data_ptr(), t._version) for t in ts), str(device))
        st = self.__dict__.get("_amp_pack")
        if st is None or st[0].numel() < nbytes or st[0].device != device:
            st = [_bytes(nbytes, device), None]
            self.__dict__["_amp_pack"] = st
        valid = st[1] == key
        st[1] = key
        return st[0], 1 if valid else 0

    def _native_tensors(self):
        ts = self.__dict__.get("_amp_tensors")
        if ts is not None and self.__dict__.get("_amp_tensor_ids") != self._live_ids():
            ts = None           # a parameter / buffer object was replaced (load_state_dict(assign=True) on a parent, p = nn.Parameter(..))
        if ts is None:
            sd = self.state_dict(keep_vars=True)
            lib = _lib.lib()
            n = getattr(lib, self._abi_count)()
            names = [getattr(lib, self._abi_name)(i).decode() for i in range(n)]
            if list(sd.keys()) != names:
                raise RuntimeError("ampnet_b200: state_dict order of %s does not match the library's parameter table"
                                   % type(self).__name__)
            ts = list(sd.values())
            self.__dict__["_amp_tensors"] = ts
            self.__dict__["_amp_tensor_ids"] = self._live_ids()
        dev = ts[0].device
        if dev.type != "cuda":
            raise RuntimeError("ampnet_b200: %s must live on a CUDA device (no CPU fallback); call .to('cuda')"
                               % type(self).__name__)
        for t in ts:
            if t.device != dev or not t.is_contiguous() or (t.is_floating_point() and t.dtype != torch.float32):
                raise RuntimeError("ampnet_b200: parameters must be contiguous float32 tensors on one CUDA device")
        return ts

    def _apply(self, fn, *a, **k):  # .to() / .cuda() replace buffer objects: drop the cache
        self.__dict__.pop("_amp_tensors", None)
        for m in self.children():
            m.__dict__.pop("_amp_tensors", None)
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.__dict__.pop("_amp_tensors", None)
        return super().load_state_dict(*a, **k)


def _input(t, name, device):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("ampnet_b200: `%s` must be a CUDA tensor (no CPU fallback)" % name)
    if t.device != device:
        raise RuntimeError("ampnet_b200: `%s` is on %s but the module is on %s" % (name, t.device, device))
    if t.dtype != torch.float32:
        raise RuntimeError("ampnet_b200: `%s` must be float32, got %s" % (name, t.dtype))
    return t.contiguous()


def _grad_targets(tensors):
    """Where the library writes each parameter gradient, and what the autograd Function returns for it.

    Default: a fresh tensor per trainable parameter, handed to autograd (which stores or accumulates it in `.grad`).
    A parameter with a gradient sink (`parallel.GradAllReduce(..., zero_copy=True)` makes `p.grad` a view of its flat
    all-reduce buffer and sets `p._amp_grad_sink`) is written in place by the FIRST backward of a step and returned as
    None: no pack / unpack copies around the collective. Further backward calls of the same step (the scripts run the encoder
    once per window, train_pointnet-attention.py:396-435) return ordinary gradients, which autograd adds to `.grad` --
    the same view -- in place. `GradAllReduce.all_reduce()` ends the step. Frozen parameters get a scratch target
    because the library writes every gradient."""
    grads, targets = [], []
    for t in tensors:
        sink = getattr(t, "_amp_grad_sink", None)
        if sink is not None and t.requires_grad and not getattr(t, "_amp_sink_written", False):
            t._amp_sink_written = True
            grads.append(None); targets.append(sink)
        elif t.is_floating_point() and t.requires_grad:
            g = torch.empty_like(t)
            grads.append(g); targets.append(g)
        else:
            grads.append(None); targets.append(torch.empty_like(t) if isinstance(t, nn.Parameter) else None)
    return grads, targets


def _row_strided(t, name, device, align=1):
    """[A, B, C] float32 input read in place when its rows are dense and evenly spaced (a slice of a wider tensor, as the
    scripts pass: train_pointnet-attention.py:427-433); anything else is copied. Returns (tensor, floats between rows)."""
    if (isinstance(t, torch.Tensor) and t.is_cuda and t.device == device and t.dtype == torch.float32 and t.dim() == 3
            and t.stride(2) == 1 and t.stride(1) >= t.shape[2] and t.stride(1) % align == 0
            and (t.shape[0] == 1 or t.stride(0) == t.shape[1] * t.stride(1))
            and t.data_ptr() % (4 * align) == 0):
        return t, t.stride(1)
    t = _input(t, name, device)
    return t, (t.shape[-1] if t.dim() == 3 else 0)


class TransformationNet(nn.Module):
    """Parameter container with the reference's layout (pointnetAtt.py:10-26). It runs inside BasePointNet's
    fused forward; the reference scripts never call it on its own."""

    def __init__(self, input_dim, output_dim, device=None):
        super().__init__()
        self.device = device
        self.output_dim = output_dim
        self.conv_1 = nn.Conv1d(input_dim, 64, 1, bias=False)
        self.conv_2 = nn.Conv1d(64, 128, 1, bias=False)
        self.conv_3 = nn.Conv1d(128, 256, 1, bias=False)
        self.bn_1 = nn.BatchNorm1d(64)
        self.bn_2 = nn.BatchNorm1d(128)
        self.bn_3 = nn.BatchNorm1d(256)
        self.bn_4 = nn.BatchNorm1d(256)
        self.bn_5 = nn.BatchNorm1d(128)
        self.fc_1 = nn.Linear(256, 256, bias=False)
        self.fc_2 = nn.Linear(256, 128, bias=False)
        self.fc_3 = nn.Linear(128, self.output_dim * self.output_dim)

    def forward(self, x):
        raise RuntimeError("ampnet_b200: TransformationNet runs fused inside BasePointNet.forward")


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, training, precision, pack, *tensors):
        lib = _lib.lib()
        B, N, _ = x.shape
        pack_buf, pack_valid = pack if pack is not None else (None, 0)
        dev = x.device
        out = torch.empty((B, N, 320), dtype=torch.float32, device=dev)
        ft = torch.empty((B, 64, 64), dtype=torch.float32, device=dev)
        tr = 1 if training else 0
        saved_bytes = lib.amp_encoder_saved_bytes(B, N, tr)
        saved = _bytes(saved_bytes, dev) if training else None
        ws_bytes = lib.amp_encoder_workspace_bytes(B, N, 0)
        ws = _bytes(ws_bytes, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.amp_encoder_fwd(_ptr_array(tensors), x.data_ptr(), B, N, tr, precision, out.data_ptr(), ft.data_ptr(),
                                           saved.data_ptr() if training else None, saved_bytes, ws.data_ptr(), ws_bytes,
                                           pack_buf.data_ptr() if pack_buf is not None else None,
                                           pack_buf.numel() if pack_buf is not None else 0, pack_valid, _lib.stream_ptr()))
        if training:
            ctx.tensors = tensors
            ctx.saved = saved
            ctx.save_for_backward(x, out, ft)
        else:
            ctx.mark_non_differentiable(out, ft)
        return out, ft

    @staticmethod
    def backward(ctx, d_out, d_ft):
        lib = _lib.lib()
        x, out, ft = ctx.saved_tensors
        B, N, _ = x.shape
        dev = x.device
        tensors = ctx.tensors
        d_out = torch.zeros_like(out) if d_out is None else d_out.contiguous()
        d_ft = None if d_ft is None else d_ft.contiguous()
        grads, targets = _grad_targets(tensors)
        ws_bytes = lib.amp_encoder_workspace_bytes(B, N, 1)
        ws = _bytes(ws_bytes, dev)
        saved = ctx.saved
        with torch.cuda.device(dev):
            _lib.check(lib.amp_encoder_bwd(_ptr_array(tensors), _ptr_array(targets), x.data_ptr(), out.data_ptr(),
                                           ft.data_ptr(), d_out.data_ptr(), d_ft.data_ptr() if d_ft is not None else None,
                                           B, N, saved.data_ptr(), saved.numel(), ws.data_ptr(), ws_bytes,
                                           _lib.stream_ptr()))
        ctx.saved = None
        return (None, None, None, None) + tuple(grads)


class BasePointNet(_NativeModule):
    """Drop-in for the reference BasePointNet (pointnetAtt.py:50-112).

    forward(x [B, N, 9]) -> (out [B, N, 320] = [global 256 repeated | local 64], feature_transform [B, 64, 64]).
    In .train() mode BatchNorm uses batch statistics and updates its running buffers in place, and the outputs
    are differentiable w.r.t. the parameters; in .eval() mode the outputs carry no autograd graph."""
    _abi_count = "amp_encoder_param_count"
    _abi_name = "amp_encoder_param_name"

    def __init__(self, point_dimension=2, return_local_features=False, global_feat_dim=256, device="cuda"):
        super().__init__()
        self.global_feat_dim = global_feat_dim
        self.point_dimension = point_dimension
        self.return_local_features = return_local_features
        self.input_transform = TransformationNet(input_dim=point_dimension, output_dim=point_dimension, device=device)
        self.feature_transform = TransformationNet(input_dim=64, output_dim=64, device=device)
        self.conv_1 = nn.Conv1d(9 + point_dimension, 64, 1, bias=False)
        self.conv_2 = nn.Conv1d(64, 64, 1, bias=False)
        self.conv_3 = nn.Conv1d(64, 64, 1, bias=False)
        self.conv_4 = nn.Conv1d(64, 128, 1, bias=False)
        self.conv_5 = nn.Conv1d(128, 128, 1, bias=False)
        self.conv_6 = nn.Conv1d(128, self.global_feat_dim, 1, bias=False)
        self.bn_1 = nn.BatchNorm1d(64)
        self.bn_2 = nn.BatchNorm1d(64)
        self.bn_3 = nn.BatchNorm1d(64)
        self.bn_4 = nn.BatchNorm1d(128)
        self.bn_5 = nn.BatchNorm1d(128)
        self.bn_6 = nn.BatchNorm1d(self.global_feat_dim)

    def forward(self, x):
        if self.point_dimension != 3 or self.global_feat_dim != 256:
            raise RuntimeError("ampnet_b200: BasePointNet kernels are built for point_dimension=3, global_feat_dim=256 "
                               "(the configuration of train_pointnet-attention.py:110-113)")
        tensors = self._native_tensors()
        x = _input(x, "x", tensors[0].device)
        if x.dim() != 3 or x.shape[2] != 9:
            raise ValueError("x must be [B, N, 9], got %s" % (tuple(x.shape),))
        if x.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("ampnet_b200: BasePointNet has no gradient with respect to its input points (the scripts never "
                               "ask for one: train_pointnet-attention.py:407-410 builds them from data); detach `x`")
        prec = self._precision()
        pack = None
        if not self.training and prec == 0:
            pack = self._pack_cache(_lib.lib().amp_encoder_pack_bytes(), x.device)
        out, ft = _EncoderFn.apply(x, self.training, prec, pack, *tensors)
        if self.return_local_features:
            return out, ft
        return out[:, 0, :256], ft


class _SegFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gl, lo, cent, training, meta, *tensors):
        lib = _lib.lib()
        npc, group_rows, mask, E, heads, C, p, seed, precision, gl_ld, lo_ld, pack = meta
        pack_buf, pack_valid = pack if pack is not None else (None, 0)
        W, B, _ = gl.shape
        R = lo.shape[1]
        dev = lo.device
        logits = torch.empty((B, C, R), dtype=torch.float32, device=dev)
        saved_bytes = lib.amp_seg_saved_bytes(B, W, R, E, heads)
        saved = _bytes(saved_bytes, dev)
        ws_bytes = lib.amp_seg_workspace_bytes(B, W, R, E, 0)
        ws = _bytes(ws_bytes, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.amp_seg_fwd(_ptr_array(tensors), gl.data_ptr(), gl_ld, lo.data_ptr(), lo_ld, cent.data_ptr(), npc,
                                       group_rows.data_ptr(), mask.data_ptr() if mask is not None else None, B, W, R, E,
                                       heads, C, 1 if training else 0, precision, p, seed, logits.data_ptr(), saved.data_ptr(),
                                       saved_bytes, ws.data_ptr(), ws_bytes,
                                       pack_buf.data_ptr() if pack_buf is not None else None,
                                       pack_buf.numel() if pack_buf is not None else 0, pack_valid, _lib.stream_ptr()))
        if training:
            ctx.tensors = tensors
            ctx.saved = saved
            ctx.meta = meta
            ctx.save_for_backward(gl, lo, cent)
        else:
            ctx.mark_non_differentiable(logits)
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        lib = _lib.lib()
        gl, lo, cent = ctx.saved_tensors
        npc, group_rows, mask, E, heads, C, p, seed, _, _, lo_ld, _ = ctx.meta
        tensors = ctx.tensors
        W, B, _ = gl.shape
        R = lo.shape[1]
        dev = lo.device
        d_logits = d_logits.contiguous()
        grads, targets = _grad_targets(tensors)
        d_gl = torch.empty(gl.shape, dtype=torch.float32, device=dev)       # dense, whatever the strides of the inputs
        d_lo = torch.empty(lo.shape, dtype=torch.float32, device=dev)
        ws_bytes = lib.amp_seg_workspace_bytes(B, W, R, E, 1)
        ws = _bytes(ws_bytes, dev)
        saved = ctx.saved
        with torch.cuda.device(dev):
            _lib.check(lib.amp_seg_bwd(_ptr_array(tensors), _ptr_array(targets), lo.data_ptr(), lo_ld, cent.data_ptr(), npc,
                                       group_rows.data_ptr(), d_logits.data_ptr(), B, W, R, E, heads, C, p, seed,
                                       d_gl.data_ptr(), d_lo.data_ptr(), saved.data_ptr(), saved.numel(), ws.data_ptr(),
                                       ws_bytes, _lib.stream_ptr()))
        ctx.saved = None
        return (d_gl, d_lo, None, None, None) + tuple(grads)


class SegmentationWithAttention(_NativeModule):
    """Drop-in for the reference SegmentationWithAttention (pointnetAtt.py:154-209).

    forward(gl_feats [W, B, E], lo_feats [B, sumN, 64], centroids [B, W, 2], np_cluster list[int],
            attn_mask [B, W] bool | None) -> (logits [B, num_classes, sumN], 0)."""
    _abi_count = "amp_seg_param_count"
    _abi_name = "amp_seg_param_name"

    def __init__(self, embed_dim, num_heads, num_classes=2, local_dim=128, dropout=0.3, device="cuda"):
        super().__init__()
        self.embed_dim = embed_dim
        self.device = device
        self.num_heads = num_heads
        self.num_classes = num_classes
        self.local_dim = local_dim
        self.dropout_p = float(dropout)
        self.fc1 = nn.Linear(2, 16)
        self.fc2 = nn.Linear(16, embed_dim)
        self.attention = nn.MultiheadAttention(embed_dim, num_heads=num_heads, dropout=dropout)
        self.conv_2 = nn.Conv1d(local_dim + embed_dim, int(embed_dim / 2), 1)
        self.conv_3 = nn.Conv1d(int(embed_dim / 2), 64, 1)
        self.conv_4 = nn.Conv1d(64, num_classes, 1)
        self.dropout = nn.Dropout(dropout)
        self.bn_2 = nn.BatchNorm1d(int(embed_dim / 2))
        self.bn_3 = nn.BatchNorm1d(64)
        self._group_cache = {}

    def _groups(self, np_cluster, device):
        key = (tuple(int(n) for n in np_cluster), str(device))
        hit = self._group_cache.get(key)
        if hit is None:
            if len(self._group_cache) > 256:
                self._group_cache.clear()
            npc = (ctypes.c_int32 * len(key[0]))(*key[0])
            starts, acc = [], 0
            for n in key[0]:
                starts.append(acc)
                acc += n
            hit = (npc, torch.tensor(starts, dtype=torch.int32, device=device), acc)
            self._group_cache[key] = hit
        return hit

    def forward(self, gl_feats, lo_feats, centroids, np_cluster, attn_mask=None):
        if self.embed_dim != 256 or self.local_dim != 64:
            raise RuntimeError("ampnet_b200: SegmentationWithAttention kernels are built for embed_dim=256, local_dim=64 "
                               "(the configuration of train_pointnet-attention.py:118)")
        tensors = self._native_tensors()
        dev = tensors[0].device
        gl, gl_ld = _row_strided(gl_feats, "gl_feats", dev)
        lo, lo_ld = _row_strided(lo_feats, "lo_feats", dev, align=4)
        cent = centroids.to(device=dev, dtype=torch.float32).contiguous()
        W, B, E = gl.shape
        npc, group_rows, total = self._groups(np_cluster, dev)
        if len(np_cluster) != W:
            raise ValueError("np_cluster has %d entries but gl_feats has %d blocks" % (len(np_cluster), W))
        if lo.dim() != 3 or lo.shape[0] != B or lo.shape[1] != total or lo.shape[2] != 64:
            raise ValueError("lo_feats must be [B, sum(np_cluster), 64], got %s" % (tuple(lo.shape),))
        if tuple(cent.shape) != (B, W, 2):
            raise ValueError("centroids must be [B, W, 2], got %s" % (tuple(cent.shape),))
        mask = None
        if attn_mask is not None:
            mask = attn_mask.to(device=dev).to(torch.uint8).contiguous()
            if tuple(mask.shape) != (B, W):
                raise ValueError("attn_mask must be [B, W]")
        p = self.dropout_p if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0.0 else 0
        prec = self._precision()
        pack = None
        if not self.training and prec == 0:
            pack = self._pack_cache(_lib.lib().amp_seg_pack_bytes(self.num_classes), dev)
        meta = (npc, group_rows, mask, E, self.num_heads, self.num_classes, p, seed, prec, gl_ld, lo_ld, pack)
        logits = _SegFn.apply(gl, lo, cent, self.training, meta, *tensors)
        return logits, 0

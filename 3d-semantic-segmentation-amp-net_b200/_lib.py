"""ctypes binding of csrc/libampnet_b200.so (the C ABI of include/ampnet_b200.h).

There is no CPU fallback: if the library is missing the import of any op fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libampnet_b200.so")

_c = ctypes
_vp, _i64, _i32, _sz, _dbl = _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_size_t, _c.c_double
_f32, _u64 = _c.c_float, _c.c_uint64

# name -> (restype, argtypes); must list every symbol declared in include/ampnet_b200.h
SIGNATURES = {
    "amp_last_error": (_c.c_char_p, []),
    "amp_abi_version": (_c.c_int, []),
    "amp_launch_count": (_i64, []),
    "amp_path_count": (_i64, [_c.c_char_p]),
    "amp_set_dropout_offset": (_c.c_int, [_c.c_void_p]),
    "amp_adam_chunk_elems": (_i32, []),
    "amp_adam_step": (_c.c_int, [_vp, _i64, _vp, _f32, _f32, _f32, _f32, _vp]),
    "amp_debug_set_disabled": (_c.c_int, [_c.c_char_p]),
    "amp_fps_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "amp_fps_f32": (_c.c_int, [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "amp_fps_f64": (_c.c_int, [_vp, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "amp_gather_rows": (_c.c_int, [_vp, _i64, _i64, _i64, _i32, _vp, _i64, _vp, _vp]),
    "amp_kmeans_assign_f32": (_c.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "amp_kmeans_gather_feats_f32": (_c.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp]),
    "amp_kmeans_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "amp_kmeans_constrained_f32": (_c.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32,
                                              _dbl, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "amp_kmeans_regroup": (_c.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "amp_minmax_f64": (_c.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "amp_window_ids_f64": (_c.c_int, [_vp, _vp, _i64, _i64, _dbl, _dbl, _i32, _i32, _i32, _i32, _vp, _vp]),
    "amp_window_partition_workspace_bytes": (_sz, [_i64]),
    "amp_window_partition": (_c.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "amp_filter_normalize_workspace_bytes": (_sz, [_i64]),
    "amp_filter_normalize_f64": (_c.c_int, [_vp, _vp, _vp, _i64, _dbl, _dbl, _vp, _vp, _vp, _sz, _vp]),
    "amp_assemble_windows_f32": (_c.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _dbl, _dbl, _vp, _vp, _vp]),
    "amp_encoder_param_count": (_c.c_int, []),
    "amp_encoder_param_name": (_c.c_char_p, [_c.c_int]),
    "amp_encoder_saved_bytes": (_sz, [_i64, _i64, _i32]),
    "amp_encoder_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "amp_encoder_pack_bytes": (_sz, []),
    "amp_encoder_fwd": (_c.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _sz, _i32, _vp]),
    "amp_encoder_bwd": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _sz, _vp, _sz, _vp]),
    "amp_seg_param_count": (_c.c_int, []),
    "amp_seg_param_name": (_c.c_char_p, [_c.c_int]),
    "amp_seg_saved_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32]),
    "amp_seg_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32]),
    "amp_seg_pack_bytes": (_sz, [_i32]),
    "amp_seg_fwd": (_c.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _u64,
                               _vp, _vp, _sz, _vp, _sz, _vp, _sz, _i32, _vp]),
    "amp_seg_bwd": (_c.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _f32, _u64,
                               _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "amp_tc_linear_workspace_bytes": (_sz, [_i32, _i32]),
    "amp_tc_linear_bf16": (_c.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "amp_wgrad_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "amp_wgrad_f32": (_c.c_int, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "ampnet_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'`; there is no CPU or PyTorch fallback for this path" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("ampnet_b200: %s (code %d)" % (lib().amp_last_error().decode(), rc))


def launch_count():
    return int(lib().amp_launch_count())


def path_count(name):
    """Launches served so far by the kernel family `name` (amp_path_count of include/ampnet_b200.h)."""
    return int(lib().amp_path_count(name.encode()))


def set_disabled(names):
    """Debug switch: turn optional fast paths off by name (None restores the AMP_DISABLE environment list)."""
    check(lib().amp_debug_set_disabled(None if names is None else ",".join(names).encode()))


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name, dtype=None):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("ampnet_b200: `%s` must be a CUDA tensor (no CPU fallback)" % name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("ampnet_b200: `%s` must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("ampnet_b200: `%s` must be contiguous" % name)
    return t

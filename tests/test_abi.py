"""The C-ABI library loads and exports every symbol include/ampnet_b200.h declares (no compute)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    with open(os.path.join(ROOT, "include", "ampnet_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    assert "amp_fps_f32" in names and "amp_kmeans_assign_f32" in names and "amp_last_error" in names


def test_library_exports_every_declared_symbol(amp):
    handle = ctypes.CDLL(amp._lib.LIB_PATH)
    for name in _declared():
        assert hasattr(handle, name), "missing export: " + name


def test_binding_table_matches_header(amp):
    assert sorted(amp._lib.SIGNATURES) == _declared()


def test_version_and_error_string(amp):
    lib = amp._lib.lib()
    assert lib.amp_abi_version() >= 1000
    assert isinstance(lib.amp_last_error(), bytes)


def test_bad_arguments_fail_without_gpu(amp):
    lib = amp._lib.lib()
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.amp_fps_f32(None, 1, 10, 3, 4, 0, None, None, None, 0, None) == -1
    assert b"null" in lib.amp_last_error()
    assert lib.amp_kmeans_assign_f32(None, None, 10, 3, None, None, None) == -1
    assert lib.amp_fps_workspace_bytes(64, 40000, 4) == 0
    assert lib.amp_fps_workspace_bytes(1, 1 << 20, 4) > 0


def test_ops_refuse_cpu_tensors(amp):
    import pytest
    import torch
    with pytest.raises(RuntimeError, match="CUDA"):
        amp.fps_indices(torch.zeros(100, 3), 10)
    with pytest.raises(RuntimeError, match="CUDA"):
        amp.kmeans_assign(torch.zeros(100, 3), torch.zeros(2, 3))


def _header_prototypes():
    """name -> list of C parameter types, parsed from the declarations of include/ampnet_b200.h."""
    with open(os.path.join(ROOT, "include", "ampnet_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(amp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if args == ["void"] else args
    return protos


def _ctype_class(c_arg):
    """The ctypes class a C parameter declaration maps to in _lib.SIGNATURES."""
    if "*" in c_arg:
        return ctypes.c_void_p
    base = re.sub(r"\b(const|unsigned)\b", "", c_arg).split()
    table = {"int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "int": ctypes.c_int, "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t,
             "float": ctypes.c_float, "double": ctypes.c_double}
    return table[base[0]]


def test_streamed_forward_needs_a_gpu(amp):
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="CUDA device"):
        amp.StreamedForward(lambda x: x, (torch.zeros(2, 3),))


def test_binding_argument_types_match_header(amp):
    """Every ctypes argtypes list has the arity and the scalar / pointer kinds of the header's prototype (an ABI change that
    reaches only one of the two sides -- e.g. the gl_ld / lo_ld strides of amp_seg_fwd -- would otherwise show up on the GPU only)."""
    protos = _header_prototypes()
    for name, (restype, argtypes) in amp._lib.SIGNATURES.items():
        assert name in protos, name
        want = [_ctype_class(a) for a in protos[name]]
        got = [ctypes.c_void_p if t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "contents") else t for t in argtypes]
        assert len(got) == len(want), "%s: %d ctypes arguments, header has %d" % (name, len(got), len(want))
        for i, (g, w) in enumerate(zip(got, want)):
            assert ctypes.sizeof(g) == ctypes.sizeof(w) and (g is ctypes.c_void_p) == (w is ctypes.c_void_p) \
                and (g in (ctypes.c_float, ctypes.c_double)) == (w in (ctypes.c_float, ctypes.c_double)), \
                "%s argument %d: ctypes %s vs header '%s'" % (name, i, g.__name__, protos[name][i])

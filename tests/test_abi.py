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

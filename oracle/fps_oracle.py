"""CPU restatement of the reference farthest-point sampler -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/utils/utils.py:889-933 (`fps`):
  * distances only on columns 0:3                                   (:894)
  * first pick is row 0                                             (:907-908)
  * picked rows leave the candidate set (np.delete)                 (:911, :931)
  * distance to the LAST pick, computed in the input dtype as
    ((a-b)**2).sum(-1) == (dx*dx + dy*dy) + dz*dz, no FMA           (:919-920)
  * running minimum kept per candidate                              (:923)
  * next pick = first maximum over the remaining candidates, which
    are in ascending original order -> lowest original index wins   (:927-928)
  * returns rows pc[sample_inds] in pick order                      (:933)
The restatement keeps candidates in place and masks picked ones with -1 (all real
distances are >= 0), which gives the same argmax as deleting them.
Pinned against the unmodified reference by tests/test_oracle_pinned.py and
tests/golden/fps_*.npz (oracle/make_golden.py).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def fps_indices(pc, n_samples, start_idx=0):
    """Indices (int64 [n_samples]) that reference `fps` would pick. NumPy, one cloud."""
    pts = np.asarray(pc)[:, :3]
    if pts.dtype not in (np.float32, np.float64):
        pts = pts.astype(np.float64)
    n = pts.shape[0]
    if n_samples > n:
        # utils.py:927 np.argmax on an empty candidate array raises ValueError
        raise ValueError("n_samples (%d) > number of points (%d)" % (n_samples, n))
    if not np.isfinite(pts).all():
        raise ValueError("non-finite coordinates")
    x, y, z = (np.ascontiguousarray(pts[:, i]) for i in range(3))
    dists = np.full(n, np.inf, dtype=pts.dtype)
    out = np.zeros(n_samples, dtype=np.int64)
    last = int(start_idx)
    out[0] = last
    dists[last] = -1.0
    for i in range(1, n_samples):
        dx = x[last] - x
        dy = y[last] - y
        dz = z[last] - z
        d = (dx * dx + dy * dy) + dz * dz          # same order as numpy's 3-term sum(-1)
        np.minimum(dists, d, out=dists, where=dists >= 0)
        last = int(np.argmax(dists))               # first maximum = lowest index
        out[i] = last
        dists[last] = -1.0
    return out


def fps(pc, n_samples):
    """Same return value as the reference: the picked ROWS (utils.py:933)."""
    pc = np.asarray(pc)
    return pc[fps_indices(pc, n_samples)]


# ---- C restatement (oracle/fps_oracle.c), used for the larger parity sizes and the CPU baseline ----
_clib = None


def _load_c():
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle C library not built: run `make -C oracle` "
                               "(or __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        lib.oracle_fps_f32.restype = ctypes.c_int
        lib.oracle_fps_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        lib.oracle_fps_f64.restype = ctypes.c_int
        lib.oracle_fps_f64.argtypes = lib.oracle_fps_f32.argtypes
        lib.oracle_kmeans_assign_f32.restype = ctypes.c_int
        lib.oracle_kmeans_assign_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                 ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        _clib = lib
    return _clib


def fps_indices_c(pc, n_samples, start_idx=0):
    """Same as fps_indices, through the C restatement (about 40x faster than NumPy)."""
    pc = np.asarray(pc)
    if pc.dtype not in (np.float32, np.float64):
        pc = pc.astype(np.float64)
    pc = np.ascontiguousarray(pc)
    n, d = pc.shape
    out = np.empty(n_samples, dtype=np.int64)
    fn = _load_c().oracle_fps_f32 if pc.dtype == np.float32 else _load_c().oracle_fps_f64
    rc = fn(pc.ctypes.data, n, d, n_samples, start_idx, out.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_fps failed with code %d" % rc)
    return out

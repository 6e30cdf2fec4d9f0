"""CPU restatement (numpy, float64) of the data-preparation arithmetic in front of the hot path -- TEST INFRASTRUCTURE ONLY.

  window_split()       data_proc/1_get_windows_split.py:52-80   LAS tile -> W x W metre windows
  filter_normalize()   data_proc/2_preprocessing_filter_norm.py:40-104   drop ground / noise classes and outliers, build the
                       13-column row, normalise x, y, HAG, clip intensity / NIR / NDVI

File I/O (laspy, pickles, the md5-keyed NIR join of 1_...:141-148 / 2_...:61-67) is not restated: the functions take and
return arrays. PINNED: tests/test_dataprep.py runs the UNMODIFIED reference functions on synthetic tiles through a fake
`laspy` module (oracle/fake_laspy.py) and requires bit-identical windows / rows.
"""
import numpy as np

DROP_CLASSES = (2, 7, 8, 13, 24, 30)     # ground classes and sensor noise (2_preprocessing_filter_norm.py:41-48)


def window_grid(x, y, w_size=(40, 40)):
    """(x0, y0, nx, ny) of the window grid: range(round(min), round(max), W) on both axes (1_get_windows_split.py:53-61;
    Python's round(): half to even)."""
    x0, y0 = round(float(np.min(x))), round(float(np.min(y)))
    nx = len(range(x0, round(float(np.max(x))), int(w_size[0])))
    ny = len(range(y0, round(float(np.max(y))), int(w_size[1])))
    return x0, y0, nx, ny


def window_split(x, y, w_size=(40, 40)):
    """Window id of every point, y-major (id = iy * nx + ix), -1 for points no window takes: the reference's masks are
    strict on both sides (`pc < x + W` and `pc > x`, :58-62), so points on a grid line, left of round(min) or right of the
    last window are dropped. Returns (ids int32 [P], nx, ny). Window w holds the points with id w in their original order."""
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    x0, y0, nx, ny = window_grid(x, y, w_size)
    ids = np.full(len(x), -1, dtype=np.int32)
    for iy in range(ny):
        ylo = y0 + iy * int(w_size[1])
        by = np.logical_and(y < (ylo + w_size[1]), y > ylo)
        for ix in range(nx):
            xlo = x0 + ix * int(w_size[0])
            b = np.logical_and(np.logical_and(x < (xlo + w_size[0]), x > xlo), by)
            ids[b] = iy * nx + ix
    return ids, nx, ny


def filter_normalize(x, y, z, hag, cls, intensity, red, green, blue, nir, max_z=100.0, max_intensity=5000, n_points=1024):
    """One window: returns the float64 [n, 13] array the reference pickles, or None when it stores nothing
    (empty after filtering, zero x / y extent, or fewer than n_points rows)   (2_preprocessing_filter_norm.py:40-123).
    `nir` is the per-point NIR value the reference looks up in its md5-keyed dictionary (:61-67)."""
    cls = np.asarray(cls)
    nir = np.asarray(nir).astype(np.int64)       # the reference rebuilds NIR from a dict of Python ints (:61-69): an int64 array,
    keep = np.ones(len(cls), dtype=bool)         # so `nir - red` against the uint16 colour does not wrap
    for c in DROP_CLASSES:
        keep &= cls != c
    hag = np.asarray(hag, dtype=np.float64)
    keep &= hag <= max_z                                                   # :51
    keep &= hag >= 0                                                       # :53
    if not keep.any():                                                     # :56
        return None
    f = lambda a: np.asarray(a)[keep]
    x, y, z, hag, cls, intensity, red, green, blue, nir = (f(a) for a in (x, y, z, hag, cls, intensity, red, green, blue, nir))
    with np.errstate(divide="ignore", invalid="ignore"):
        ndvi = (nir - red) / (nir + red)                                   # :70
    pc = np.vstack((x, y, hag, cls, intensity / max_intensity, red / 65536.0, green / 65536.0, blue / 65536.0,
                    nir / 65535.0, ndvi, x, y, z)).transpose()             # :75-90
    if pc[:, 0].max() - pc[:, 0].min() == 0 or pc[:, 1].max() - pc[:, 1].min() == 0:       # :92
        return None
    pc[:, 0] = 2 * ((pc[:, 0] - pc[:, 0].min()) / (pc[:, 0].max() - pc[:, 0].min())) - 1
    pc[:, 1] = 2 * ((pc[:, 1] - pc[:, 1].min()) / (pc[:, 1].max() - pc[:, 1].min())) - 1
    pc[:, 2] = pc[:, 2] / max_z
    pc = pc[pc[:, 2] >= 0]
    pc[:, 4] = np.clip(pc[:, 4], 0.0, 1.0)
    pc[:, 8] = np.clip(pc[:, 8], 0.0, 1.0)
    pc[:, 9] = (pc[:, 9] + 1) / 2
    pc[:, 9] = np.clip(pc[:, 9], 0.0, 1.0)
    if pc.shape[0] < n_points:                                             # :107
        return None
    return pc


def synthetic_tile(n_points, seed, extent=(400.0, 300.0), origin=(431000.25, 4582000.5)):
    """A LAS-like tile: float64 UTM x / y / z, HeightAboveGround, ASPRS classes, 16-bit colours and NIR."""
    rng = np.random.default_rng(seed)
    x = origin[0] + rng.random(n_points) * extent[0]
    y = origin[1] + rng.random(n_points) * extent[1]
    x[::97] = np.round(x[::97])                       # some points exactly on integer coordinates (grid lines drop them)
    hag = rng.random(n_points) * 130.0 - 5.0          # some below 0 and above max_z
    z = 200.0 + hag + rng.random(n_points)
    cls = rng.choice(np.array([1, 2, 3, 4, 5, 6, 7, 8, 13, 14, 15, 24, 30]), n_points).astype(np.uint8)
    intensity = rng.integers(0, 7000, n_points).astype(np.uint16)
    red, green, blue, nir = (rng.integers(0, 65536, n_points).astype(np.uint16) for _ in range(4))
    return {"x": x, "y": y, "z": z, "HeightAboveGround": hag, "classification": cls, "intensity": intensity,
            "red": red, "green": green, "blue": blue, "nir": nir}

"""Pin the oracle: against the committed golden vectors (always) and against the unmodified
reference executed in-container (when /root/reference is present)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import fps_oracle, kmeans_oracle


def _fps_golden():
    z = np.load(os.path.join(GOLDEN, "fps_reference.npz"))
    names = sorted(k[:-4] for k in z.files if k.endswith("__pc"))
    return [(n, z[n + "__pc"], z[n + "__idx"]) for n in names]


@pytest.mark.parametrize("name,pc,idx", _fps_golden(), ids=[c[0] for c in _fps_golden()])
def test_fps_oracle_matches_golden(name, pc, idx):
    assert (fps_oracle.fps_indices(pc, len(idx)) == idx).all()
    assert (fps_oracle.fps_indices_c(pc, len(idx)) == idx).all()


@pytest.mark.needs_reference
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fps_oracle_matches_reference_live(reference, seed):
    _, uu, _ = reference
    rng = np.random.default_rng(seed)
    pc = rng.random((2000 + 333 * seed, 11), dtype=np.float32)
    if seed == 2:
        pc[::3] = pc[1::3][: len(pc[::3])]         # many exact duplicates -> ties
    S = 300
    aug = np.concatenate([pc, np.arange(len(pc), dtype=np.float32)[:, None]], 1)
    ref_rows = uu.fps(aug, S)
    ref_idx = ref_rows[:, -1].astype(np.int64)
    assert (fps_oracle.fps_indices(pc, S) == ref_idx).all()
    assert (fps_oracle.fps_indices_c(pc, S) == ref_idx).all()
    assert (fps_oracle.fps(pc, S) == ref_rows[:, :-1]).all()


def test_fps_oracle_edges():
    pc = np.random.default_rng(0).random((10, 3), dtype=np.float32)
    assert sorted(fps_oracle.fps_indices(pc, 10)) == list(range(10))
    with pytest.raises(ValueError):
        fps_oracle.fps_indices(pc, 11)          # reference: ValueError from np.argmax of empty
    bad = pc.copy(); bad[3, 1] = np.nan
    with pytest.raises(ValueError):
        fps_oracle.fps_indices(bad, 4)
    assert fps_oracle.fps_indices(pc, 1).tolist() == [0]


def test_kmeans_assign_oracle_c_matches_numpy():
    import ctypes
    rng = np.random.default_rng(3)
    x = rng.random((5001, 3), dtype=np.float32)
    c = rng.random((9, 3), dtype=np.float32)
    c[4] = c[2]                                     # exact tie between two centroids
    lab, md = kmeans_oracle.assign(x, c)
    lib = fps_oracle._load_c()
    lab_c = np.empty(len(x), np.int32); md_c = np.empty(len(x), np.float32)
    assert lib.oracle_kmeans_assign_f32(x.ctypes.data, c.ctypes.data, len(x), 9, lab_c.ctypes.data, md_c.ctypes.data) == 0
    assert (lab == lab_c).all() and (md == md_c).all()
    assert not (lab == 4).any()                     # first minimum wins


def test_kmeans_oracle_invariants():
    rng = np.random.default_rng(5)
    n, k = 2048 * 4, 4
    x = rng.random((n, 3), dtype=np.float32)
    lab, c, it = kmeans_oracle.kmeans_constrained(x, k, 2048, 2048)
    assert (np.bincount(lab, minlength=k) == 2048).all() and 1 <= it <= 10
    un, _ = kmeans_oracle.assign(x, c)
    d = kmeans_oracle.sqdist(x, c)
    inertia_c = d[np.arange(n), lab].sum(); inertia_u = d[np.arange(n), un].sum()
    assert inertia_c <= 1.35 * inertia_u            # balancing costs a bounded amount of inertia
    # min-only variant
    x2 = rng.random((2048 * 3 + 700, 3), dtype=np.float32)
    lab2, _, _ = kmeans_oracle.kmeans_constrained(x2, 3, 2048, None)
    assert (np.bincount(lab2, minlength=3) >= 2048).all()
    groups = kmeans_oracle.regroup(np.arange(len(x2)), lab2, 3)
    assert sorted(np.concatenate(groups).tolist()) == list(range(len(x2)))
    assert all((np.diff(g) > 0).all() for g in groups)

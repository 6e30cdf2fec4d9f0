"""Data preparation in front of the hot path: window tiling of a LAS tile and the filter / normalise step.

 * CPU, needs_reference: the oracle (oracle/dataprep_oracle.py) against the UNMODIFIED reference functions
   data_proc/1_get_windows_split.py::split_dataset_windows and 2_preprocessing_filter_norm.py::remove_ground_and_outliers,
   executed in-container on synthetic tiles through a stand-in `laspy` (oracle/fake_laspy.py): bit-identical windows / rows.
 * GPU: the CUDA kernels (amp_window_ids_f64, amp_window_partition, amp_filter_normalize_f64) against the oracle, bit-exact.
"""
import glob
import hashlib
import importlib.util
import os
import pickle
import sys
import types

import numpy as np
import pytest

from oracle import dataprep_oracle as dpo, fake_laspy


def _load(reference_root, fname, modname):
    sys.modules["laspy"].read = fake_laspy.read                 # the stub module oracle/ref_import.py registered
    if "alive_progress" not in sys.modules:
        m = types.ModuleType("alive_progress")

        class _Bar:
            def __init__(self, *a, **k): pass
            def __enter__(self): return lambda *a, **k: None
            def __exit__(self, *a): return False
        m.alive_bar = _Bar
        sys.modules["alive_progress"] = m
    spec = importlib.util.spec_from_file_location(modname, os.path.join(reference_root, "data_proc", fname))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.needs_reference
@pytest.mark.parametrize("seed,w", [(1, 40), (2, 100), (3, 25)])
def test_window_split_oracle_matches_reference(reference, tmp_path, seed, w):
    from oracle import ref_import
    mod = _load(ref_import.REFERENCE_ROOT, "1_get_windows_split.py", "ref_get_windows_split")
    tile = dpo.synthetic_tile(30000, seed)
    las_dir = tmp_path / "las"; las_dir.mkdir()
    path = str(las_dir / "blockA.las"); open(path, "w").close()
    fake_laspy.register(path, tile)
    got = []
    mod.store_las_file_from_pc = lambda pc, file, save_path, dataset: got.append(np.array(pc, copy=True))
    mod.save_path = str(tmp_path / "out")
    mod.split_dataset_windows("RIBERA", str(las_dir), [w, w])
    ids, nx, ny = dpo.window_split(tile["x"], tile["y"], (w, w))
    want = [i for i in range(nx * ny) if (ids == i).any()]
    assert len(got) == len(want) > 4
    full = np.vstack((tile["x"], tile["y"], tile["z"], tile["classification"], tile["intensity"], tile["red"], tile["green"],
                      tile["blue"], tile["nir"]))
    for pc, i in zip(got, want):
        assert np.array_equal(pc, full[:, ids == i])
    assert (ids == -1).sum() > 0                                  # points on grid lines / outside the grid are dropped


@pytest.mark.needs_reference
@pytest.mark.parametrize("seed", [5, 6])
def test_filter_normalize_oracle_matches_reference(reference, tmp_path, seed):
    from oracle import ref_import
    mod = _load(ref_import.REFERENCE_ROOT, "2_preprocessing_filter_norm.py", "ref_preprocessing_filter_norm")
    tile = dpo.synthetic_tile(6000, seed, extent=(40.0, 40.0))
    path = str(tmp_path / "pc_RIBERA_blockA_w3.las"); open(path, "w").close()
    fake_laspy.register(path, tile)
    # the md5-keyed NIR dictionary the reference joins on (1_get_windows_split.py:141-148, 2_preprocessing...:61-67)
    nir_dict = {}
    for x, y, z, n in zip(tile["x"], tile["y"], tile["z"], tile["nir"]):
        nir_dict[hashlib.md5((str(int(x)) + "_" + str(int(y)) + "_" + str(int(z))).encode()).hexdigest()] = int(n)
    with open(path.replace(".las", "") + "_NIR.pkl", "wb") as f:
        pickle.dump(nir_dict, f)
    nir_joined = np.array([nir_dict[hashlib.md5((str(int(x)) + "_" + str(int(y)) + "_" + str(int(z))).encode()).hexdigest()]
                           for x, y, z in zip(tile["x"], tile["y"], tile["z"])])
    mod.out_path = str(tmp_path / "out"); os.makedirs(mod.out_path)
    mod.remove_ground_and_outliers(path, max_z=100.0, max_intensity=5000, n_points=1024, dataset="RIBERA")
    files = glob.glob(os.path.join(mod.out_path, "*.pkl"))
    assert len(files) == 1
    with open(files[0], "rb") as f:
        ref = pickle.load(f)
    ours = dpo.filter_normalize(tile["x"], tile["y"], tile["z"], tile["HeightAboveGround"], tile["classification"], tile["intensity"],
                                tile["red"], tile["green"], tile["blue"], nir_joined, 100.0, 5000, 1024)
    assert ours is not None and ours.dtype == ref.dtype == np.float64 and ours.shape == ref.shape and ours.shape[1] == 13
    assert np.array_equal(ours, ref, equal_nan=True)
    assert ours[:, 0].min() == -1 and ours[:, 0].max() == 1 and ours[:, 2].max() <= 1 and (ours[:, 9] >= 0).all()


def _tile_cols(tile):
    return np.stack([tile["x"], tile["y"], tile["z"], tile["HeightAboveGround"], tile["classification"].astype(np.float64),
                     tile["intensity"].astype(np.float64), tile["red"].astype(np.float64), tile["green"].astype(np.float64),
                     tile["blue"].astype(np.float64), tile["nir"].astype(np.float64)], axis=1)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,w,extent", [(200000, 11, 40, (400.0, 300.0)), (50000, 12, 25, (900.0, 800.0)), (3000, 13, 100, (40.0, 40.0))])
def test_window_split_and_filter_normalize_bit_exact(amp, cuda, n, seed, w, extent):
    import torch
    tile = dpo.synthetic_tile(n, seed, extent=extent)
    cols = torch.from_numpy(_tile_cols(tile)).to(cuda)
    n0 = amp._lib.launch_count()
    sp = amp.split_windows(cols[:, 0], cols[:, 1], (w, w))               # strided views of the [P, 10] tensor, read in place
    ids, nx, ny = dpo.window_split(tile["x"], tile["y"], (w, w))
    assert (sp["nx"], sp["ny"]) == (nx, ny)
    assert np.array_equal(sp["ids"].cpu().numpy(), ids)
    order, offsets = sp["order"].cpu().numpy(), sp["offsets"].cpu().numpy()
    assert offsets[0] == 0 and offsets[-1] == (ids >= 0).sum() and len(offsets) == nx * ny + 1
    for wi in range(nx * ny):                                            # stable: original order inside every window
        assert np.array_equal(order[offsets[wi]:offsets[wi + 1]], np.flatnonzero(ids == wi))
    assert sorted(order.tolist()) == list(range(n))                      # a permutation: dropped points sit behind offsets[-1]
    rows, oo, stored = amp.filter_normalize_windows(cols, sp["order"], sp["offsets"], 100.0, 5000, 8)
    assert amp._lib.launch_count() - n0 >= 10
    rows = rows.cpu().numpy()
    n_stored = 0
    for wi in range(nx * ny):
        m = ids == wi
        ref = dpo.filter_normalize(tile["x"][m], tile["y"][m], tile["z"][m], tile["HeightAboveGround"][m], tile["classification"][m],
                                   tile["intensity"][m], tile["red"][m], tile["green"][m], tile["blue"][m], tile["nir"][m], 100.0, 5000, 8) if m.any() else None
        if ref is None:
            assert not stored[wi]
        else:
            assert stored[wi]
            assert np.array_equal(rows[oo[wi]:oo[wi + 1]], ref, equal_nan=True)
            n_stored += 1
    assert n_stored >= 1


@pytest.mark.gpu
def test_dataprep_edge_cases(amp, cuda):
    import torch
    # every point on a grid line or outside the grid: no window takes any
    x = torch.tensor([10.0, 50.0, 90.0, 9.4], dtype=torch.float64, device=cuda)
    y = torch.tensor([20.0, 20.0, 20.0, 60.0], dtype=torch.float64, device=cuda)
    sp = amp.split_windows(x, y, (40, 40))
    ids, nx, ny = dpo.window_split(x.cpu().numpy(), y.cpu().numpy(), (40, 40))
    assert np.array_equal(sp["ids"].cpu().numpy(), ids) and int(sp["offsets"][-1]) == (ids >= 0).sum()
    # a window whose kept rows share one x: zero extent -> nothing stored (2_preprocessing_filter_norm.py:92)
    cols = torch.zeros((8, 10), dtype=torch.float64, device=cuda)
    cols[:, 0] = 5.0; cols[:, 1] = torch.arange(8, device=cuda); cols[:, 3] = 1.0; cols[:, 4] = 5.0
    order = torch.arange(8, dtype=torch.int64, device=cuda); offsets = torch.tensor([0, 8], dtype=torch.int64, device=cuda)
    rows, oo, stored = amp.filter_normalize_windows(cols, order, offsets, 100.0, 5000, 1)
    assert rows.shape[0] == 0 and not stored[0]
    with pytest.raises(RuntimeError, match="CUDA"):
        amp.split_windows(torch.zeros(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64))

"""The single-file block store against the reference's per-window formats: the `.pt` tensors of
data_proc/3_kmeans.py:116 (read by pointNet/datasets.py:335) and the pickle pairs of utils/utils.py:526-533 (read by
test_pointnet_att_segmen.py:140-143) must come back bit-identical."""
import os
import pickle

import numpy as np
import pytest
import torch


def _windows(seed, n=7):
    rng = np.random.default_rng(seed)
    return [torch.from_numpy(rng.random((2048, 13, int(rng.integers(1, 10))), dtype=np.float32)) for _ in range(n)]


def test_round_trip_of_3_kmeans_pt_files(amp, tmp_path):
    wins = _windows(0)
    paths = []
    for i, w in enumerate(wins):                                   # what 3_kmeans.py writes: torch.save(pc_w, ...)
        p = str(tmp_path / ("tower_kmeans_w%d.pt" % i)); torch.save(w, p); paths.append(p)
    store = amp.convert_kmeans_pt_files(paths, str(tmp_path / "blocks.amp"))
    assert len(store) == len(wins) and store.dims == 13
    for i, p in enumerate(paths):
        ref = torch.load(p, map_location=torch.device("cpu"))      # datasets.py:335
        got = store.as_kmeans_pt(i)
        assert got.dtype == ref.dtype and torch.equal(got, ref)
        assert store.by_name["tower_kmeans_w%d" % i] == i
        assert store.windows[i]["offset"] % 4096 == 0
        for b, blk in enumerate(store.blocks(i)):                  # a block is a contiguous [2048, D] encoder input
            assert blk.flags["C_CONTIGUOUS"] and np.array_equal(blk, ref[:, :, b].numpy())
    assert os.path.getsize(str(tmp_path / "blocks.amp")) < sum(os.path.getsize(p) for p in paths) + 4096 * (len(paths) + 2)


def test_round_trip_of_cluster_pickles_with_ragged_blocks(amp, tmp_path):
    rng = np.random.default_rng(1)
    with amp.BlockStoreWriter(str(tmp_path / "test.amp"), 10) as w:
        refs = []
        for i in range(4):
            k = int(rng.integers(2, 6))
            clusters = [torch.from_numpy(rng.random((2048 + int(rng.integers(0, 500)), 10), dtype=np.float32)) for _ in range(k)]
            cent = torch.from_numpy(rng.random((k, 2), dtype=np.float32))
            # what kmeans_clustering(out_path=...) writes (utils.py:526-533)
            with open(str(tmp_path / ("w%d_clusters_list.pkl" % i)), "wb") as f:
                pickle.dump(clusters, f)
            with open(str(tmp_path / ("w%d_centroids.pkl" % i)), "wb") as f:
                pickle.dump(cent, f)
            w.add_blocks("w%d" % i, clusters, cent)
            refs.append(i)
    store = amp.BlockStore(str(tmp_path / "test.amp"))
    for i in refs:
        with open(str(tmp_path / ("w%d_clusters_list.pkl" % i)), "rb") as f:
            clusters = pickle.load(f)                              # test_pointnet_att_segmen.py:140-143
        with open(str(tmp_path / ("w%d_centroids.pkl" % i)), "rb") as f:
            cent = pickle.load(f)
        got, gc = store.as_cluster_pickles(i)
        assert len(got) == len(clusters) and all(torch.equal(a, b) for a, b in zip(got, clusters)) and torch.equal(gc, cent)
        assert store.rows(i) == [c.shape[0] for c in clusters]
    with pytest.raises(ValueError):
        store.as_kmeans_pt(0) if len(set(store.rows(0))) != 1 else (_ for _ in ()).throw(ValueError())


def test_rejects_foreign_files_and_wrong_shapes(amp, tmp_path):
    p = str(tmp_path / "junk.amp")
    with open(p, "wb") as f:
        f.write(b"x" * 200)
    with pytest.raises(ValueError):
        amp.BlockStore(p)
    with amp.BlockStoreWriter(str(tmp_path / "a.amp"), 9) as w:
        with pytest.raises(ValueError):
            w.add_blocks("bad", [np.zeros((5, 8), np.float32)])


@pytest.mark.gpu
def test_store_feeds_the_encoder(amp, cuda, tmp_path):
    """A window read from the store goes to the device in one copy and its blocks are ready encoder inputs."""
    from oracle import nn_params
    wins = _windows(3, 3)
    with amp.BlockStoreWriter(str(tmp_path / "s.amp"), 13) as w:
        for i, t in enumerate(wins):
            w.add_kmeans_pt("w%d" % i, t)
    store = amp.BlockStore(str(tmp_path / "s.amp"))
    enc = amp.BasePointNet(point_dimension=3, return_local_features=True, global_feat_dim=256, device=cuda)
    enc.load_state_dict(nn_params.synthetic_state_dict(nn_params.encoder_shapes(), 3)); enc.to(cuda).eval()
    for i, t in enumerate(wins):
        dev, rows = store.to_device(i, cuda, non_blocking=False)
        k = t.shape[2]
        assert rows == [2048] * k and torch.equal(dev.cpu().view(k, 2048, 13), t.permute(2, 0, 1))
        x9 = torch.cat((dev[:, 0:3], dev[:, 4:10]), 1).view(k, 2048, 9).contiguous()      # datasets.py:358 keeps columns 0:3, 4:10
        out, ft = enc(x9)
        assert tuple(out.shape) == (k, 2048, 320) and torch.isfinite(out).all()

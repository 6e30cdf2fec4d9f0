"""A single memory-mappable block store for the k-means blocks.

The reference keeps one file per window: `torch.save(pc_w [n_points, D, k])` from data_proc/3_kmeans.py:116 (loaded by
LidarKmeansDataset.__getitem__, pointNet/datasets.py:335) and a pair of pickles per window from
utils/utils.py:526-533 kmeans_clustering (`<name>_clusters_list.pkl` = list of [n_i, D] tensors, `<name>_centroids.pkl` =
[k, 2]; loaded by test_pointnet_att_segmen.py:140-143). Opening and unpickling tens of thousands of small files is the
data-loader's floor. Here every window of a dataset lives in ONE file:

    [ 64-byte header | window data, each window 4096-byte aligned | index (JSON) ]

  window data = its blocks back to back, block b = float32 [rows_b, D] row-major: a block is a ready [N, D] encoder input,
  the whole window one contiguous range (one pread / one H2D copy), `np.memmap` gives zero-copy CPU views.

The store converts to and from both reference formats exactly (same tensors, same dtypes), so it is a drop-in source for
the untouched datasets: `as_kmeans_pt(i)` == torch.load(<3_kmeans .pt>), `as_cluster_pickles(i)` == the two pickles.
"""
import json
import os
import struct

import numpy as np
import torch

MAGIC = b"AMPBLK01"
ALIGN = 4096
HEADER = struct.Struct("<8sQQQQ24x")            # magic, version, n_windows, index_offset, index_bytes


class BlockStoreWriter:
    """Append windows, then close(): the index is written behind the data and the header patched last (a partly written
    file never looks valid)."""

    def __init__(self, path, dims):
        self.path, self.dims = path, int(dims)
        self.f = open(path, "wb")
        self.f.write(b"\0" * HEADER.size)
        self.index = []

    def _pad(self):
        pos = self.f.tell()
        pad = (-pos) % ALIGN
        if pad:
            self.f.write(b"\0" * pad)
        return pos + pad

    def add_blocks(self, name, blocks, centroids=None):
        """blocks: list of [rows_b, D] float arrays / tensors (kmeans_clustering's cluster list); centroids [k, 2] or None."""
        arrs = [np.ascontiguousarray(b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else b, dtype=np.float32) for b in blocks]
        for a in arrs:
            if a.ndim != 2 or a.shape[1] != self.dims:
                raise ValueError("every block must be [rows, %d]" % self.dims)
        off = self._pad()
        for a in arrs:
            self.f.write(a.tobytes())
        cent = None
        if centroids is not None:
            cent = np.asarray(centroids.detach().cpu().numpy() if isinstance(centroids, torch.Tensor) else centroids, dtype=np.float32).reshape(-1, 2).tolist()
        self.index.append({"name": str(name), "offset": off, "rows": [int(a.shape[0]) for a in arrs], "centroids": cent})

    def add_kmeans_pt(self, name, pc_w):
        """pc_w: the [n_points, D, k] FloatTensor data_proc/3_kmeans.py:99-116 saves per window."""
        t = pc_w if isinstance(pc_w, torch.Tensor) else torch.as_tensor(pc_w)
        if t.dim() != 3 or t.shape[1] != self.dims:
            raise ValueError("pc_w must be [n_points, %d, k]" % self.dims)
        self.add_blocks(name, [t[:, :, j] for j in range(t.shape[2])])

    def close(self):
        idx = json.dumps({"dims": self.dims, "windows": self.index}).encode()
        idx += b" " * ((-len(idx)) % 4)                 # keep the file a whole number of float32 words (np.memmap of the data)
        off = self._pad()
        self.f.write(idx)
        self.f.seek(0)
        self.f.write(HEADER.pack(MAGIC, 1, len(self.index), off, len(idx)))
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class BlockStore:
    """Read side: memory-mapped, random access by window index or name."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            magic, version, n, ioff, ibytes = HEADER.unpack(f.read(HEADER.size))
            if magic != MAGIC or version != 1:
                raise ValueError("%s is not an ampnet_b200 block store" % path)
            f.seek(ioff)
            meta = json.loads(f.read(ibytes).decode())
        self.dims = int(meta["dims"])
        self.windows = meta["windows"]
        assert len(self.windows) == n
        self.by_name = {w["name"]: i for i, w in enumerate(self.windows)}
        self.mm = np.memmap(path, dtype=np.float32, mode="r")
        self._pinned = None

    def __len__(self):
        return len(self.windows)

    def rows(self, i):
        return list(self.windows[i]["rows"])

    def window_array(self, i):
        """Zero-copy view [sum rows, D] of window i (blocks back to back)."""
        w = self.windows[i]
        n = sum(w["rows"])
        a = self.mm[w["offset"] // 4: w["offset"] // 4 + n * self.dims]
        return a.reshape(n, self.dims)

    def blocks(self, i):
        """List of zero-copy [rows_b, D] views."""
        a, out, o = self.window_array(i), [], 0
        for r in self.windows[i]["rows"]:
            out.append(a[o:o + r]); o += r
        return out

    def centroids(self, i):
        c = self.windows[i]["centroids"]
        return None if c is None else torch.tensor(c, dtype=torch.float32).reshape(-1, 2)

    # ---- the reference's formats, rebuilt exactly ----
    def as_kmeans_pt(self, i):
        """The tensor torch.load() returns for a data_proc/3_kmeans.py file: FloatTensor [n_points, D, k]."""
        rows = self.windows[i]["rows"]
        if len(set(rows)) != 1:
            raise ValueError("window %d has blocks of different sizes: not a 3_kmeans.py window" % i)
        a = self.window_array(i).reshape(len(rows), rows[0], self.dims)
        return torch.from_numpy(np.ascontiguousarray(a.transpose(1, 2, 0)))

    def as_cluster_pickles(self, i):
        """(clusters_list, centroids) as unpickled by test_pointnet_att_segmen.py:140-143."""
        return [torch.from_numpy(np.array(b)) for b in self.blocks(i)], self.centroids(i)

    # ---- device side: one pinned staging copy + one H2D per window (or run of windows) ----
    def to_device(self, i, device="cuda", non_blocking=True):
        """Window i on the device as ([sum rows, D] float32 tensor, rows list). The pinned staging buffer is reused between
        calls: synchronise (or pass non_blocking=False) before the next call overwrites it."""
        a = self.window_array(i)
        n = a.shape[0] * a.shape[1]
        if self._pinned is None or self._pinned.numel() < n:
            self._pinned = torch.empty(max(n, 1 << 20), dtype=torch.float32).pin_memory()
        stage = self._pinned[:n].view(a.shape)
        stage.numpy()[...] = a
        return stage.to(device, non_blocking=non_blocking), self.rows(i)


def convert_kmeans_pt_files(paths, out_path):
    """Pack the per-window `.pt` files of data_proc/3_kmeans.py into one store (names = file stems)."""
    first = torch.load(paths[0], map_location="cpu")
    with BlockStoreWriter(out_path, first.shape[1]) as w:
        for p in paths:
            w.add_kmeans_pt(os.path.splitext(os.path.basename(p))[0], first if p == paths[0] else torch.load(p, map_location="cpu"))
    return BlockStore(out_path)

"""K-means block split: host-side mirror of the reference interface.

  kmeans_clustering(...)  <- utils/utils.py:473-535 (online / test-time variant, clusters >= n_points)
  split_kmeans(...)       <- data_proc/3_kmeans.py:27-116 (offline variant, blocks of exactly n_points)
  get_cluster_centroid    <- utils/utils.py:538-543

The reference delegates the clustering to the third-party `KMeansConstrained.fit_predict`
(random, CPU, OR-tools). Here it is the deterministic restatement defined in
oracle/kmeans_oracle.py, executed by csrc/kmeans.cu through the C ABI. No CPU fallback.
"""
import os
import pickle
import random

import numpy as np
import torch

from . import _lib
from .sampling import _device, fps as _fps

KMAX = 32


def kmeans_assign(feats, centroids, return_min_d2=False):
    """labels[i] = argmin_j ((x_i - c_j)**2).sum(-1), first minimum. feats [n,3], centroids [k,3] CUDA f32."""
    _lib.require_cuda(feats, "feats", torch.float32)
    _lib.require_cuda(centroids, "centroids", torch.float32)
    n, k = feats.shape[0], centroids.shape[0]
    if feats.dim() != 2 or feats.shape[1] != 3 or centroids.dim() != 2 or centroids.shape[1] != 3:
        raise ValueError("feats must be [n,3] and centroids [k,3]")
    labels = torch.empty((n,), dtype=torch.int32, device=feats.device)
    mind = torch.empty((n,), dtype=torch.float32, device=feats.device) if return_min_d2 else None
    with torch.cuda.device(feats.device):
        _lib.check(_lib.lib().amp_kmeans_assign_f32(feats.data_ptr(), centroids.data_ptr(), n, k,
                                                    labels.data_ptr(), mind.data_ptr() if mind is not None else None,
                                                    _lib.stream_ptr()))
    return (labels, mind) if return_min_d2 else labels


def gather_feats(pc, cols):
    """feats = pc[:, cols] for 3 columns (`in_pc[:, i_f]`, 3_kmeans.py:81-82). pc [n, D] CUDA f32."""
    _lib.require_cuda(pc, "pc", torch.float32)
    n, D = pc.shape
    feats = torch.empty((n, 3), dtype=torch.float32, device=pc.device)
    with torch.cuda.device(pc.device):
        _lib.check(_lib.lib().amp_kmeans_gather_feats_f32(pc.data_ptr(), n, D, int(cols[0]), int(cols[1]),
                                                          int(cols[2]), feats.data_ptr(), _lib.stream_ptr()))
    return feats


N_INIT = 5        # the reference's KMeansConstrained(n_init=5) (3_kmeans.py:78-80, utils.py:500-503)


_WINDOW_TABLES = {}


def _window_tables(offsets, ks, dev):
    """Device copies of the window offsets / cluster counts, cached by content: a tile is usually split with the same window
    list call after call, and two small pageable host-to-device copies per call are two stream synchronisations."""
    key = (offsets.tobytes(), ks.tobytes(), str(dev))
    hit = _WINDOW_TABLES.get(key)
    if hit is None:
        if len(_WINDOW_TABLES) > 64:
            _WINDOW_TABLES.clear()
        hit = (torch.from_numpy(offsets.copy()).to(dev), torch.from_numpy(ks.copy()).to(dev))
        _WINDOW_TABLES[key] = hit
    return hit


def kmeans_constrained_windows(feats, offsets, ks, size_min=0, size_max=0, max_iter=10, tol=1e-2, n_init=1, check_range=True):
    """Constrained k-means of W independent windows in one launch.

    feats [total,3] CUDA f32; offsets: W+1 ints (host list/array); ks: W ints (host).
    Returns (labels int32 [total], centroids f32 [W,kmax,3], n_iter int32 [W]) on the device.
    check_range=False skips the fixed-point range check of the features (one device-to-host read, i.e. a stream
    synchronisation per call) for callers that know their columns are normalised."""
    _lib.require_cuda(feats, "feats", torch.float32)
    offsets = np.asarray(offsets, dtype=np.int64)
    ks = np.asarray(ks, dtype=np.int32)
    W = len(ks)
    total = int(feats.shape[0])
    if len(offsets) != W + 1 or offsets[0] != 0 or offsets[-1] != total:
        raise ValueError("offsets must be [0, ..., total] with W+1 entries")
    sizes = np.diff(offsets)
    if (sizes <= 0).any():
        raise ValueError("empty window")
    if (ks < 1).any() or (ks > KMAX).any():
        raise ValueError("k must be in [1, %d]" % KMAX)
    if (ks > sizes).any():
        raise ValueError("k larger than the number of points of a window")
    if size_min and (size_min * ks.astype(np.int64) > sizes).any():
        raise ValueError("size_min * k > n")
    if size_max and (size_max * ks.astype(np.int64) < sizes).any():
        raise ValueError("size_max * k < n")
    kmax = int(ks.max())
    dev = feats.device
    # fixed-point sums (rint(x * 2^32) in int64, csrc/kmeans.cu): sum |x| and sum x^2 of a window must stay below 2^31. The
    # reference clusters normalised columns ([-1, 1] coordinates, [0, 1] features); raw UTM coordinates would wrap silently.
    amax = float(feats.abs().max()) if (total and check_range) else 0.0
    if not np.isfinite(amax) or amax * amax * float(sizes.max()) >= 2.0 ** 31 or amax * float(sizes.max()) >= 2.0 ** 31:
        raise ValueError("kmeans: clustering features must be finite and normalised (max |x| = %g over windows of up to %d points "
                         "overflows the fixed-point centroid sums); normalise the columns first as "
                         "data_proc/2_preprocessing_filter_norm.py does" % (amax, int(sizes.max())))
    d_off, d_ks = _window_tables(offsets, ks, dev)
    labels = torch.empty((total,), dtype=torch.int32, device=dev)
    cent = torch.empty((W, kmax, 3), dtype=torch.float32, device=dev)
    n_iter = torch.empty((W,), dtype=torch.int32, device=dev)
    lib = _lib.lib()
    ws_bytes = lib.amp_kmeans_workspace_bytes(total, W, kmax)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.amp_kmeans_constrained_f32(
            feats.data_ptr(), d_off.data_ptr(), d_ks.data_ptr(), W, total, int(sizes.max()), kmax,
            int(size_min), int(size_max), int(max_iter), float(tol), int(n_init), labels.data_ptr(), cent.data_ptr(),
            n_iter.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr()))
    return labels, cent, n_iter


def regroup_windows(labels, offsets, ks, pc=None):
    """Stable regroup: (order int64 [total], counts int32 [W,kmax], xy_mean f32 [W,kmax,2] | None)."""
    _lib.require_cuda(labels, "labels", torch.int32)
    offsets = np.asarray(offsets, dtype=np.int64)
    ks = np.asarray(ks, dtype=np.int32)
    W, kmax = len(ks), int(ks.max())
    dev = labels.device
    d_off, d_ks = _window_tables(offsets, ks, dev)
    order = torch.empty((labels.shape[0],), dtype=torch.int64, device=dev)
    counts = torch.empty((W, kmax), dtype=torch.int32, device=dev)
    xy = None
    stride = 0
    if pc is not None:
        _lib.require_cuda(pc, "pc", torch.float32)
        xy = torch.empty((W, kmax, 2), dtype=torch.float32, device=dev)
        stride = pc.shape[1]
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().amp_kmeans_regroup(labels.data_ptr(), d_off.data_ptr(), d_ks.data_ptr(), W, kmax,
                                                 pc.data_ptr() if pc is not None else None, stride,
                                                 order.data_ptr(), counts.data_ptr(),
                                                 xy.data_ptr() if xy is not None else None, _lib.stream_ptr()))
    return order, counts, xy


def get_cluster_centroid(pc):
    """(mean x, mean y) of a cluster (utils/utils.py:538-543)."""
    return torch.stack([pc[:, 0].mean(0), pc[:, 1].mean(0)], dim=0)


def cluster_window(pc_dev, k, cols, size_min, size_max, max_iter=10, tol=1e-2, n_init=N_INIT):
    """One window on the device: returns (grouped rows [n, D] sorted by (label, index),
    counts [k] (host ints), xy_mean [k,2] device, labels device)."""
    feats = gather_feats(pc_dev, cols)
    n = pc_dev.shape[0]
    if n > 28000:
        n_init = 1                      # restarts need the on-chip kernel (include/ampnet_b200.h)
    labels, _, _ = kmeans_constrained_windows(feats, [0, n], [k], size_min, size_max, max_iter, tol, n_init)
    order, counts, xy = regroup_windows(labels, [0, n], [k], pc_dev)
    grouped = pc_dev.index_select(0, order)
    return grouped, counts[0].cpu().tolist(), xy[0], labels


def kmeans_clustering(in_pc, n_points=2048, get_centroids=True, max_clusters=18, out_path='', file_name='',
                      device=None):
    """Drop-in for utils/utils.py:473-535.

    in_pc: torch tensor [1, P, D] or [P, D] (CPU as in the reference, or CUDA).
    Returns (cluster_lists: list of tensors [n_i, D] on in_pc's device, centroids tensor [k, 2])."""
    in_pc = in_pc.squeeze(0)
    cluster_lists = []
    centroids = torch.FloatTensor()
    if in_pc.shape[0] >= 2 * n_points:
        k_clusters = int(np.floor(in_pc.shape[0] / n_points))
        if k_clusters > max_clusters:
            k_clusters = max_clusters
        dev = in_pc.device if in_pc.is_cuda else _device(device)
        pc_dev = in_pc.to(dev, torch.float32).contiguous()
        grouped, counts, xy, _ = cluster_window(pc_dev, k_clusters, (0, 1, 8), n_points, 0)   # i_f, utils.py:504
        grouped = grouped.to(in_pc.device).to(in_pc.dtype)
        o = 0
        for c in counts[:k_clusters]:
            if c:
                cluster_lists.append(grouped[o:o + c])
            o += c
        if get_centroids:
            keep = torch.tensor([c > 0 for c in counts[:k_clusters]])
            centroids = xy[:k_clusters].to(in_pc.device)[keep.to(in_pc.device)]
    else:
        cluster_lists.append(in_pc)
        if get_centroids:
            centroids = get_cluster_centroid(in_pc).unsqueeze(0)
    if out_path:
        if not os.path.exists(out_path):
            os.makedirs(out_path)
        with open(os.path.join(out_path, file_name + '_clusters_list') + '.pkl', 'wb') as f:
            pickle.dump(cluster_lists, f)
        with open(os.path.join(out_path, file_name + '_centroids') + '.pkl', 'wb') as f:
            pickle.dump(centroids, f)
    return cluster_lists, centroids


def split_kmeans_array(pc, n_points=2048, max_clusters=9, fps_sample=False, device=None):
    """The array-in / tensor-out core of data_proc/3_kmeans.py:27-116 (no file I/O):
    pc: ndarray [P, D]; returns torch.FloatTensor [n_points, D, k] (CPU, like the reference)."""
    pc = np.asarray(pc)
    if pc.shape[0] >= 2 * n_points:
        in_pc = pc
        k_clusters = int(np.ceil(in_pc.shape[0] / n_points))
        if k_clusters > max_clusters:                                   # 3_kmeans.py:57-62
            k_clusters = max_clusters
            ix = random.sample(range(in_pc.shape[0]), n_points * max_clusters)
            in_pc = in_pc[ix, :]
        elif in_pc.shape[0] < n_points * k_clusters:                    # :65-69
            points_needed = n_points * k_clusters - in_pc.shape[0]
            rdm_list = np.random.randint(0, in_pc.shape[0], points_needed)
            in_pc = np.concatenate([in_pc, in_pc[rdm_list, :]], axis=0)
        if in_pc.shape[0] % n_points != 0:                              # :71-73
            in_pc = in_pc[:n_points * (in_pc.shape[0] // n_points), :]
        dev = _device(device)
        pc_dev = torch.from_numpy(np.ascontiguousarray(in_pc, dtype=np.float32)).to(dev)
        grouped, counts, _, _ = cluster_window(pc_dev, k_clusters, (0, 1, 9), n_points, n_points)   # i_f, :81
        # [k, n_points, D] -> [n_points, D, k] (the stack/cat of :99-101)
        pc_w = grouped.view(k_clusters, n_points, -1).permute(1, 2, 0).contiguous().cpu()
    else:
        if pc.shape[0] > n_points:                                      # :107-113
            if fps_sample:
                pc = _fps(pc, n_points, device=device)
            ix = random.sample(range(pc.shape[0]), n_points)
            pc = pc[ix, :, ]
        pc_w = torch.Tensor(pc).unsqueeze(2)
    return pc_w


def split_kmeans(file_path, n_points=2048, max_clusters=9, plot=False, fps_sample=False, o_path=None,
                 device=None):
    """Drop-in for data_proc/3_kmeans.py:27 (`plot` is accepted and ignored: plotting is out of scope).
    Saves `<o_path>_kmeans_<name>.pt` like the reference (:116) when o_path is given; returns pc_w."""
    filename = file_path.split('/')[-1].split('.')[0]
    with open(file_path, 'rb') as f:
        pc = pickle.load(f)
    pc_w = split_kmeans_array(pc, n_points, max_clusters, fps_sample, device)
    if o_path is not None:
        torch.save(pc_w, o_path + '_kmeans_' + filename + '.pt')
    return pc_w

"""CPU restatement of the k-means block split -- TEST INFRASTRUCTURE ONLY.

What the reference does (data_proc/3_kmeans.py:27-116, utils/utils.py:473-535): it hands
`in_pc[:, [0, 1, 9]]` (or `[0, 1, 8]`) to `k_means_constrained.KMeansConstrained(size_min[, size_max],
n_init=5, max_iter=10, tol=1e-2, random_state=None).fit_predict` and regroups rows by label.
That solver is the third-party PyPI package `k-means-constrained` (+ `ortools` min-cost flow);
the reference pins no version, neither package is installed here and there is no network, and
the reference holds no test or golden vector at this boundary: PARITY UNPINNED for the
size-constrained solver. The reference's result is also random (random_state=None).

What this file therefore DEFINES (and the CUDA path must reproduce bit-for-bit):
  assign        argmin_k ((x - c_k)**2).sum(-1) in float32, NumPy order (d0^2 + d1^2) + d2^2,
                no FMA, first minimum wins                       [pinned to NumPy arithmetic]
  init          centroids = the k rows farthest-point-sampled from the 3 clustering features,
                start row 0 (same sampler as utils.py:889-933)   [deterministic; n_init = 1]
  update        per-cluster mean through order-independent fixed-point sums:
                sum_j = SUM rint(float64(x) * 2^32) as int64; c = float32(sum_j / 2^32 / count);
                an empty cluster keeps its previous centroid
  stop          shift = SUM_{j,d} (float64(c_new) - float64(c_old))^2 <= tol * mean_d var_d(x),
                var from the same fixed-point sums (sklearn's `tol` scaling)
  balance       capacity rounds: every unassigned point proposes to its nearest cluster that
                still has room; a cluster with more proposals than room keeps the closest
                (d^2, then lower index) and closes; repeat. size_max caps every cluster;
                size_min alone is enforced only when the plain assignment violates it (fill each
                cluster to size_min by rounds, remaining points go to their nearest cluster).
  final labels  one more constrained assignment with the final centroids.
"""
import numpy as np

from . import fps_oracle

FIX = 4294967296.0  # 2^32


def sqdist(x, c):
    """[n,k] float32 squared distances in NumPy order; x [n,3], c [k,3] float32."""
    x = np.asarray(x, dtype=np.float32)
    c = np.asarray(c, dtype=np.float32)
    d = x[:, None, :] - c[None, :, :]
    d = d * d
    return (d[..., 0] + d[..., 1]) + d[..., 2]


def assign(x, c):
    """(labels int32 [n], min_d2 float32 [n]); first minimum wins."""
    d = sqdist(x, c)
    lab = np.argmin(d, axis=1).astype(np.int32)
    return lab, d[np.arange(len(lab)), lab]


def _capacity_rounds(d, labels, room):
    """Assign every point with labels == -1 by capacity rounds. d [n,k]; room int64 [k] (mutated)."""
    n, k = d.shape
    while True:
        todo = np.flatnonzero(labels < 0)
        open_ = room > 0
        if todo.size == 0 or not open_.any():
            break
        dd = np.where(open_[None, :], d[todo], np.float32(np.inf))
        prop = np.argmin(dd, axis=1)
        pd = dd[np.arange(todo.size), prop]
        for j in np.flatnonzero(open_):
            mine = todo[prop == j]
            if mine.size == 0:
                continue
            if mine.size > room[j]:
                order = np.lexsort((mine, pd[prop == j]))      # by d^2, then by index
                mine = mine[order[:room[j]]]
            labels[mine] = j
            room[j] -= mine.size
    return labels


def constrained_assign(x, c, size_min=None, size_max=None):
    x = np.asarray(x, dtype=np.float32)
    n, k = x.shape[0], c.shape[0]
    d = sqdist(x, c)
    if size_max is not None:
        if size_max * k < n:
            raise ValueError("size_max * k < n")
        labels = np.full(n, -1, dtype=np.int32)
        if size_min is not None and size_min * k > n:
            raise ValueError("size_min * k > n")
        if size_min is not None and size_min < size_max:
            # fill every cluster to size_min first, then up to size_max
            _capacity_rounds(d, labels, np.full(k, size_min, dtype=np.int64))
            room = size_max - np.bincount(labels[labels >= 0], minlength=k).astype(np.int64)
            _capacity_rounds(d, labels, room)
        else:
            _capacity_rounds(d, labels, np.full(k, size_max, dtype=np.int64))
        return labels
    labels = np.argmin(d, axis=1).astype(np.int32)
    if size_min is None:
        return labels
    if size_min * k > n:
        raise ValueError("size_min * k > n")
    if (np.bincount(labels, minlength=k) >= size_min).all():
        return labels
    labels = np.full(n, -1, dtype=np.int32)
    _capacity_rounds(d, labels, np.full(k, size_min, dtype=np.int64))
    rest = labels < 0
    labels[rest] = np.argmin(d[rest], axis=1)
    return labels


def fixed_sums(x, labels, k):
    q = np.rint(np.asarray(x, dtype=np.float32).astype(np.float64) * FIX).astype(np.int64)
    sums = np.zeros((k, 3), dtype=np.int64)
    np.add.at(sums, labels, q)
    counts = np.bincount(labels, minlength=k).astype(np.int64)
    return sums, counts


def update_centroids(x, labels, c_old):
    k = c_old.shape[0]
    sums, counts = fixed_sums(x, labels, k)
    c = np.array(c_old, dtype=np.float32, copy=True)
    nz = counts > 0
    c[nz] = ((sums[nz].astype(np.float64) / FIX) / counts[nz, None].astype(np.float64)).astype(np.float32)
    return c, counts


def tol_abs(x, tol):
    """tol * mean_d var_d(x) from fixed-point first and second moments (float64, fixed order)."""
    x64 = np.asarray(x, dtype=np.float32).astype(np.float64)
    n = np.float64(x64.shape[0])
    s1 = np.rint(x64 * FIX).astype(np.int64).sum(axis=0)
    s2 = np.rint((x64 * x64) * FIX).astype(np.int64).sum(axis=0)
    acc = np.float64(0.0)
    for dd in range(3):
        m1 = (np.float64(s1[dd]) / FIX) / n
        m2 = (np.float64(s2[dd]) / FIX) / n
        acc = acc + (m2 - m1 * m1)
    return (acc / np.float64(3.0)) * np.float64(tol)


def center_shift(c_new, c_old):
    acc = np.float64(0.0)
    a = c_new.astype(np.float64).ravel()
    b = c_old.astype(np.float64).ravel()
    for i in range(a.size):
        t = a[i] - b[i]
        acc = acc + t * t
    return acc


def init_centroids(x, k, start_row=0):
    x = np.asarray(x, dtype=np.float32)
    return x[fps_oracle.fps_indices(x, k, start_row)].copy()


def inertia_fixed(x, labels, c):
    """Sum over the points of the float32 squared distance to their own centroid, as an order-independent fixed-point
    integer (rint(float64(d^2) * 2^32) in int64): the score the restarts are compared on."""
    x = np.asarray(x, dtype=np.float32)
    d = x - np.asarray(c, dtype=np.float32)[labels]
    d = d * d
    d2 = (d[:, 0] + d[:, 1]) + d[:, 2]
    return int(np.rint(d2.astype(np.float64) * FIX).astype(np.int64).sum())


def _one_run(x, k, size_min, size_max, max_iter, ta, start_row):
    c = init_centroids(x, k, start_row)
    it = 0
    for it in range(1, max_iter + 1):
        labels = constrained_assign(x, c, size_min, size_max)
        c_new, _ = update_centroids(x, labels, c)
        shift = center_shift(c_new, c)
        c = c_new
        if shift <= ta:
            break
    labels = constrained_assign(x, c, size_min, size_max)
    return labels, c, it


def kmeans_constrained(x, k, size_min=None, size_max=None, max_iter=10, tol=1e-2, n_init=1):
    """Returns (labels int32 [n], centroids float32 [k,3], n_iter).
    n_init restarts (the reference passes n_init=5 with random k-means++ seeds, 3_kmeans.py:78-80): restart r seeds the
    farthest-point initialisation at row (r * n) // n_init; the run with the smallest inertia_fixed() wins, the earliest
    on a tie. n_init=1 is the single run seeded at row 0."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    ta = tol_abs(x, tol)
    best = None
    for r in range(max(1, int(n_init))):
        labels, c, it = _one_run(x, k, size_min, size_max, max_iter, ta, (r * len(x)) // max(1, int(n_init)))
        if n_init <= 1:
            return labels, c, it
        score = inertia_fixed(x, labels, c)
        if best is None or score < best[0]:
            best = (score, labels, c, it)
    return best[1], best[2], best[3]


def regroup(pc, labels, k):
    """Rows grouped by ascending label, original order inside a label (the stable sort + groupby
    of 3_kmeans.py:88-90 / utils.py:508-510). Returns list of arrays."""
    order = np.argsort(labels, kind="stable")
    counts = np.bincount(labels, minlength=k)
    out, o = [], 0
    for j in range(k):
        out.append(pc[order[o:o + counts[j]]])
        o += counts[j]
    return [a for a in out if len(a)]

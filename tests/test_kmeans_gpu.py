"""Parity of the CUDA k-means block split (csrc/kmeans.cu via the C ABI) with the oracle. Bit-exact."""
import numpy as np
import pytest
import torch

from oracle import kmeans_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k", [(1, 1), (3, 2), (1023, 5), (4096, 9), (262147, 12), ((1 << 20) + 3, 16), (100003, 17),
                                  (100003, 18), (1 << 20, 33)])
def test_assign_bit_exact(amp, cuda, n, k):
    rng = np.random.default_rng(n)
    x = rng.random((n, 3), dtype=np.float32); x[:, :2] = x[:, :2] * 2 - 1
    c = rng.random((k, 3), dtype=np.float32)
    if k > 3:
        c[3] = c[1]
    lab, md = amp.kmeans_assign(torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda), return_min_d2=True)
    elab, emd = ko.assign(x, c)
    assert (lab.cpu().numpy() == elab).all() and (md.cpu().numpy() == emd).all()


@pytest.mark.parametrize("k", [4, 9, 16, 27])
def test_assign_ties_and_rounding_cases(amp, cuda, k):
    """Lattice points and lattice centroids: many exact ties (the first minimum must win) and sums whose unfused and fused
    roundings differ (coordinates scaled by an odd constant), on both the register-resident (k <= 16) and the generic kernel."""
    rng = np.random.default_rng(k)
    x = (rng.integers(-8, 9, size=(50001, 3)).astype(np.float32) * np.float32(0.3))
    c = (rng.integers(-8, 9, size=(k, 3)).astype(np.float32) * np.float32(0.3))
    c[k - 1] = c[0]
    lab, md = amp.kmeans_assign(torch.from_numpy(x).to(cuda), torch.from_numpy(c).to(cuda), return_min_d2=True)
    elab, emd = ko.assign(x, c)
    assert (lab.cpu().numpy() == elab).all() and (md.cpu().numpy() == emd).all()
    assert (elab != k - 1).all()            # the duplicate of centroid 0 never wins a tie


def test_gather_feats(amp, cuda):
    pc = torch.rand(5000, 13, device=cuda)
    assert (amp.gather_feats(pc, (0, 1, 9)) == pc[:, [0, 1, 9]]).all()


def _cmp(amp, cuda, xs, ks, smin, smax):
    offs = np.concatenate([[0], np.cumsum([len(x) for x in xs])])
    feats = torch.from_numpy(np.concatenate(xs)).to(cuda)
    lab, cent, it = amp.kmeans_constrained_windows(feats, offs, ks, smin, smax)
    lab = lab.cpu().numpy(); cent = cent.cpu().numpy(); it = it.cpu().numpy()
    for w, (x, k) in enumerate(zip(xs, ks)):
        el, ec, eit = ko.kmeans_constrained(x, k, smin or None, smax or None)
        assert it[w] == eit, "window %d n_iter" % w
        assert (cent[w, :k] == ec).all(), "window %d centroids" % w
        assert (lab[offs[w]:offs[w + 1]] == el).all(), "window %d labels" % w
    return lab, offs


def test_balanced_split_bit_exact(amp, cuda):
    rng = np.random.default_rng(21)
    xs = []
    for k in (2, 5, 9):
        x = rng.random((2048 * k, 3), dtype=np.float32); x[:, :2] = x[:, :2] * 2 - 1
        xs.append(x)
    # a clumpy window: two dense blobs + background (forces many capacity rounds)
    blob = np.concatenate([rng.normal(0.3, 0.02, (3000, 3)), rng.normal(-0.5, 0.05, (3000, 3)),
                           rng.random((2192, 3))]).astype(np.float32)
    xs.append(blob)
    lab, offs = _cmp(amp, cuda, xs, [2, 5, 9, 4], 2048, 2048)
    for w, k in enumerate([2, 5, 9, 4]):
        assert (np.bincount(lab[offs[w]:offs[w + 1]], minlength=k) == 2048).all()


def test_min_only_split_bit_exact(amp, cuda):
    rng = np.random.default_rng(22)
    xs = [rng.random((2048 * 3 + 517, 3), dtype=np.float32), rng.random((50000, 3), dtype=np.float32),
          np.concatenate([rng.normal(0.2, 0.01, (5000, 3)), rng.random((1200, 3))]).astype(np.float32)]
    ks = [3, 18, 3]
    lab, offs = _cmp(amp, cuda, xs, ks, 2048, 0)
    for w, k in enumerate(ks):
        assert (np.bincount(lab[offs[w]:offs[w + 1]], minlength=k) >= 2048).all()


def test_unconstrained_and_duplicates(amp, cuda):
    rng = np.random.default_rng(23)
    x = np.repeat(rng.random((700, 3), dtype=np.float32), 6, axis=0)
    _cmp(amp, cuda, [x], [4], 0, 0)
    _cmp(amp, cuda, [x], [2], 2048, 2100)


def test_regroup_and_reference_shaped_calls(amp, cuda):
    rng = np.random.default_rng(24)
    pc = rng.random((2048 * 4 + 300, 10), dtype=np.float32)
    t = torch.from_numpy(pc).unsqueeze(0)                      # reference passes a CPU tensor [1,P,10]
    clusters, cents = amp.kmeans_clustering(t, n_points=2048, max_clusters=18)
    el, _, _ = ko.kmeans_constrained(pc[:, [0, 1, 8]], 4, 2048, None)
    groups = ko.regroup(pc, el, 4)
    assert len(clusters) == len(groups) == 4
    for a, b in zip(clusters, groups):
        assert a.shape == b.shape and (a.numpy() == b).all()
        assert a.shape[0] >= 2048
    ref_cent = np.stack([[g[:, 0].mean(), g[:, 1].mean()] for g in groups]).astype(np.float32)
    assert cents.shape == (4, 2) and np.allclose(cents.numpy(), ref_cent, rtol=0, atol=1e-6)
    # small cloud: returned as is, centroid [1,2]
    small = torch.rand(1, 3000, 10)
    cl, ce = amp.kmeans_clustering(small)
    assert len(cl) == 1 and cl[0].shape == (3000, 10) and ce.shape == (1, 2)
    # offline variant: [2048, D, k], every block exactly 2048 rows of the input
    pc13 = rng.random((2048 * 3 - 100, 13), dtype=np.float64)
    np.random.seed(0)
    pw = amp.split_kmeans_array(pc13, 2048, 9)
    assert pw.shape == (2048, 13, 3) and pw.dtype == torch.float32
    rows = pw.permute(2, 0, 1).reshape(-1, 13).numpy()
    src = {r.tobytes() for r in pc13.astype(np.float32)}
    assert all(r.tobytes() in src for r in rows)
    big = rng.random((2048 * 11, 13))
    assert amp.split_kmeans_array(big, 2048, 9).shape == (2048, 13, 9)
    assert amp.split_kmeans_array(rng.random((3000, 13)), 2048, 9).shape == (2048, 13, 1)

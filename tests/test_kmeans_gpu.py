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
    el, _, _ = ko.kmeans_constrained(pc[:, [0, 1, 8]], 4, 2048, None, n_init=5)      # the drop-ins restart 5 times like the reference
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


def test_on_chip_window_kernel_equals_global_memory_kernel_and_oracle(amp, cuda):
    """kmeans_window_fast_kernel (state in shared memory, warp-parallel radix select) against kmeans_window_kernel (state in
    global memory) and the oracle: identical labels, centroids and iteration counts, for window sizes on both sides of the
    "coordinates fit in shared memory" limit and for all three constraint modes."""
    rng = np.random.default_rng(21)
    cases = [([5 * 2048, 3 * 2048, 2048 * 2], [5, 3, 2], 2048, 2048),        # balanced split (3_kmeans.py:78), x in shared memory
             ([9 * 2048, 7 * 2048], [9, 7], 2048, 2048),                     # 18 432 points: coordinates stay in global memory
             ([9000, 5000], [4, 2], 2048, 0),                                # size_min only (utils.py:500)
             ([6000], [3], 0, 0)]                                            # unconstrained
    for sizes, ks, smin, smax in cases:
        x = rng.random((sum(sizes), 3), dtype=np.float32)
        x[::50] = x[1::50][: len(x[::50])]                                   # exact duplicates: ties in the radix select
        offsets = np.concatenate([[0], np.cumsum(sizes)])
        xd = torch.from_numpy(x).to(cuda)
        n0 = amp._lib.path_count("kmeans_fast")
        lab, cent, it = amp.kmeans_constrained_windows(xd, offsets, ks, smin, smax)
        assert amp._lib.path_count("kmeans_fast") == n0 + 1
        amp._lib.set_disabled(["kmeans_fast"])
        try:
            lab2, cent2, it2 = amp.kmeans_constrained_windows(xd, offsets, ks, smin, smax)
        finally:
            amp._lib.set_disabled(None)
        assert torch.equal(lab, lab2) and torch.equal(cent, cent2) and torch.equal(it, it2)
        w = 0
        el, ec, eit = ko.kmeans_constrained(x[offsets[w]:offsets[w + 1]], ks[w], smin or None, smax or None)
        assert (lab[offsets[w]:offsets[w + 1]].cpu().numpy() == el).all() and (cent[w, :ks[w]].cpu().numpy() == ec).all() and int(it[w]) == eit


def test_infeasible_windows_are_flagged_not_corrupted(amp, cuda):
    """The C ABI itself (no Python-side validation): a window whose constraints cannot be met gets labels -1 / n_iter -1 and
    leaves the other windows of the call alone (out-of-bounds shared-memory atomics on label -1)."""
    import ctypes
    rng = np.random.default_rng(3)
    sizes, ks = [4096, 3000, 4096], [2, 2, 2]                                # window 1: size_max * k = 4096 >= 3000 but size_min * k = 4096 > 3000
    x = torch.from_numpy(rng.random((sum(sizes), 3), dtype=np.float32)).to(cuda)
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64, device=cuda)
    dks = torch.tensor(ks, dtype=torch.int32, device=cuda)
    lib = amp._lib.lib()
    for disabled in (None, ["kmeans_fast"]):
        amp._lib.set_disabled(disabled)
        try:
            labels = torch.full((sum(sizes),), 77, dtype=torch.int32, device=cuda)
            cent = torch.empty((3, 2, 3), dtype=torch.float32, device=cuda)
            nit = torch.empty(3, dtype=torch.int32, device=cuda)
            wsb = lib.amp_kmeans_workspace_bytes(sum(sizes), 3, 2)
            ws = torch.empty(wsb, dtype=torch.uint8, device=cuda)
            amp._lib.check(lib.amp_kmeans_constrained_f32(x.data_ptr(), offsets.data_ptr(), dks.data_ptr(), 3, sum(sizes), max(sizes), 2, 2048, 2048,
                                                          10, 1e-2, 1, labels.data_ptr(), cent.data_ptr(), nit.data_ptr(), ws.data_ptr(), wsb,
                                                          torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
        finally:
            amp._lib.set_disabled(None)
        assert nit.tolist()[1] == -1 and nit.tolist()[0] > 0 and nit.tolist()[2] > 0
        assert (labels[4096:7096] == -1).all()
        for w, (a, b) in enumerate([(0, 4096), (7096, 11192)]):
            cnt = torch.bincount(labels[a:b].long(), minlength=2)
            assert cnt.tolist() == [2048, 2048]
        # the regroup kernel skips unassigned rows instead of indexing with -1
        order, counts, _ = amp.regroup_windows(labels, offsets.cpu().numpy(), ks)
        assert counts.cpu().tolist() == [[2048, 2048], [0, 0], [2048, 2048]]


def test_unnormalised_coordinates_are_refused(amp, cuda):
    x = torch.rand(5000, 3, device=cuda) * 1000.0 + 431000.0                 # raw UTM metres: the fixed-point sums would wrap
    with pytest.raises(ValueError, match="normalised"):
        amp.kmeans_constrained_windows(x, [0, 5000], [2], 2048, 0)


def test_restarts_keep_the_lowest_inertia_and_match_the_oracle(amp, cuda):
    """n_init = 5 (3_kmeans.py:78-80): restart r seeds the FPS initialisation at row (r * n) // 5; the run with the smallest
    fixed-point inertia wins. Labels / centroids bit-exact against the oracle, and never worse than the single run."""
    rng = np.random.default_rng(31)
    sizes, ks = [5 * 2048, 3 * 2048], [5, 3]
    x = rng.random((sum(sizes), 3), dtype=np.float32)
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    xd = torch.from_numpy(x).to(cuda)
    lab5, cent5, it5 = amp.kmeans_constrained_windows(xd, offsets, ks, 2048, 2048, n_init=5)
    lab1, cent1, _ = amp.kmeans_constrained_windows(xd, offsets, ks, 2048, 2048, n_init=1)
    for w in range(2):
        xs = x[offsets[w]:offsets[w + 1]]
        el, ec, eit = ko.kmeans_constrained(xs, ks[w], 2048, 2048, n_init=5)
        got = lab5[offsets[w]:offsets[w + 1]].cpu().numpy()
        assert (got == el).all() and (cent5[w, :ks[w]].cpu().numpy() == ec).all() and int(it5[w]) == eit
        assert np.bincount(got, minlength=ks[w]).tolist() == [2048] * ks[w]
        i5 = ko.inertia_fixed(xs, got, cent5[w, :ks[w]].cpu().numpy())
        i1 = ko.inertia_fixed(xs, lab1[offsets[w]:offsets[w + 1]].cpu().numpy(), cent1[w, :ks[w]].cpu().numpy())
        assert i5 <= i1


def test_window_tables_are_cached_by_content_and_range_check_is_optional(amp, cuda):
    """The device copies of (offsets, ks) are cached by content: the same window list twice, then a different one with the
    same point count, must each give the oracle's labels; check_range=False skips only the host-side range check."""
    rng = np.random.default_rng(31)
    x = rng.random((2048 * 6, 3), dtype=np.float32); x[:, :2] = x[:, :2] * 2 - 1
    d = torch.from_numpy(x).to(cuda)
    for offs, ks in (([0, 2048 * 2, 2048 * 6], [2, 4]), ([0, 2048 * 2, 2048 * 6], [2, 4]), ([0, 2048 * 3, 2048 * 6], [3, 3])):
        lab, _, _ = amp.kmeans_constrained_windows(d, offs, ks, 2048, 2048, check_range=False)
        order, counts, _ = amp.regroup_windows(lab, offs, ks)
        lab = lab.cpu().numpy()
        for w, k in enumerate(ks):
            e, _, _ = ko.kmeans_constrained(x[offs[w]:offs[w + 1]], k, 2048, 2048)
            assert (lab[offs[w]:offs[w + 1]] == e).all(), (offs, ks, w)
            assert (counts.cpu().numpy()[w, :k] == 2048).all()

"""Parity of the CUDA farthest-point sampler (csrc/fps.cu via the C ABI) with the oracle.
Bit-exact: indices must be identical."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fps_oracle

pytestmark = pytest.mark.gpu


def _golden():
    z = np.load(os.path.join(GOLDEN, "fps_reference.npz"))
    names = sorted(k[:-4] for k in z.files if k.endswith("__pc"))
    return [(n, z[n + "__pc"], z[n + "__idx"]) for n in names]


@pytest.mark.parametrize("name,pc,idx", _golden(), ids=[c[0] for c in _golden()])
def test_golden_reference_vectors(amp, cuda, name, pc, idx):
    got = amp.fps_indices(torch.from_numpy(pc).to(cuda), len(idx)).cpu().numpy()
    assert (got == idx).all()
    rows = amp.fps(pc, len(idx))                       # reference-shaped call: ndarray in, rows out
    assert rows.dtype == pc.dtype and (rows == pc[idx]).all()


@pytest.mark.parametrize("P,S,D,B", [(1, 1, 3, 1), (33, 33, 3, 2), (1000, 100, 4, 3), (4096, 512, 11, 5),
                                     (8191, 700, 11, 2), (12289, 600, 13, 70), (40000, 2048, 11, 2),
                                     (70001, 300, 3, 1), (200000, 96, 5, 2)])
def test_matches_oracle_random(amp, cuda, P, S, D, B):
    rng = np.random.default_rng(P + S)
    pc = rng.random((B, P, D), dtype=np.float32)
    got = amp.fps_indices(torch.from_numpy(pc).to(cuda), S).cpu().numpy()
    for b in range(B if P < 20000 else 1):
        assert (got[b] == fps_oracle.fps_indices_c(pc[b], S)).all(), "cloud %d" % b


def test_ties_duplicates_and_grid(amp, cuda):
    rng = np.random.default_rng(9)
    dup = np.repeat(rng.random((500, 3), dtype=np.float32), 9, axis=0)
    grid = np.stack(np.meshgrid(*[np.arange(17)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    same = np.ones((3000, 3), dtype=np.float32)
    for pc, S in ((dup, 1200), (grid, 900), (same, 50)):
        got = amp.fps_indices(torch.from_numpy(pc).to(cuda), S).cpu().numpy()
        assert (got == fps_oracle.fps_indices_c(pc, S)).all()


def test_float64_and_start_index(amp, cuda):
    rng = np.random.default_rng(11)
    pc = rng.random((5000, 4))
    got = amp.fps_indices(torch.from_numpy(pc).to(cuda), 400).cpu().numpy()
    assert (got == fps_oracle.fps_indices_c(pc, 400)).all()
    pc32 = pc.astype(np.float32)
    got = amp.fps_indices(torch.from_numpy(pc32).to(cuda), 400, start_idx=1234).cpu().numpy()
    assert (got == fps_oracle.fps_indices_c(pc32, 400, start_idx=1234)).all()
    big = rng.random((30000, 3))
    got = amp.fps_indices(torch.from_numpy(big).to(cuda), 64).cpu().numpy()
    assert (got == fps_oracle.fps_indices_c(big, 64)).all()


def test_errors_like_reference(amp, cuda):
    pc = torch.rand(100, 3, device=cuda)
    with pytest.raises(ValueError):
        amp.fps_indices(pc, 101)                        # reference raises ValueError as well
    bad = pc.clone(); bad[7, 2] = float("nan")
    with pytest.raises(ValueError):
        amp.fps_indices(bad, 10)
    with pytest.raises(RuntimeError):
        amp.fps_indices(pc.to(torch.float16), 10)


def test_sample_fps_cascade(amp, cuda):
    """data_proc/sample_fps.py:23-31: float32 cast, 8192 then 4096 cascade."""
    rng = np.random.default_rng(13)
    pc = rng.random((20000, 11)).astype(np.float32)
    a = amp.fps(pc, 8192); b = amp.fps(a, 4096)
    ea = pc[fps_oracle.fps_indices_c(pc, 8192)]; eb = ea[fps_oracle.fps_indices_c(ea, 4096)]
    assert (a == ea).all() and (b == eb).all()


def test_full_size_properties(amp, cuda):
    """BASELINE config 2 (64 x 40000 -> 2048): size-independent properties + 2 clouds bit-exact."""
    g = torch.Generator(device="cpu").manual_seed(0)
    pc = torch.rand(64, 40000, 11, generator=g)
    d = pc.to(cuda)
    idx = amp.fps_indices(d, 2048)
    assert idx.shape == (64, 2048) and (idx[:, 0] == 0).all()
    srt = idx.sort(dim=1).values
    assert (srt[:, 1:] != srt[:, :-1]).all()            # no point picked twice
    assert (idx >= 0).all() and (idx < 40000).all()
    # the distance of each pick to the previously picked set never increases
    xyz = d[0, idx[0], :3].double()
    dm = torch.cdist(xyz, xyz)
    tri = torch.full_like(dm, float("inf")).triu(0) + dm.tril(-1)
    pick_d = tri.min(dim=1).values[1:]
    assert (pick_d[1:] <= pick_d[:-1] + 1e-12).all()
    ih = idx.cpu().numpy()
    for b in (0, 63):
        assert (ih[b] == fps_oracle.fps_indices_c(pc[b].numpy(), 2048)).all()
    rows = amp.gather_rows(d, idx)
    assert (rows[5] == d[5, idx[5]]).all()


def test_host_stream_matches_batch_calls_and_reports_bad_batches(amp, cuda):
    """amp.FpsHostStream / fps_host_stream (three batches in flight, copies overlapped): the rows of every batch equal the
    oracle's picks and the serial fps_host_batch call; a batch with a NaN raises when its result arrives."""
    rng = np.random.default_rng(77)
    B, P, S, D = 3, 3000, 200, 4
    batches = [torch.from_numpy(rng.random((B, P, D), dtype=np.float32)).pin_memory() for _ in range(5)]
    got = [r.clone() for r in amp.fps_host_stream(batches, S)]
    assert len(got) == 5
    for b, r in zip(batches, got):
        assert torch.equal(r, amp.fps_host_batch(b, S))
        for c in range(B):
            want = fps_oracle.fps_indices_c(b[c].numpy(), S)
            assert np.array_equal(r[c].numpy(), b[c].numpy()[want])
    stream = amp.FpsHostStream(batches[0], S, depth=2)
    assert sum(1 for _ in stream.run(batches[:3])) == 3          # reusable
    bad = batches[1].clone().pin_memory()
    bad[1, 17, 2] = float("nan")
    with pytest.raises(ValueError, match="non-finite"):
        list(stream.run([batches[0], bad, batches[2]]))
    assert list(amp.fps_host_stream([], S)) == []
    with pytest.raises(ValueError):
        list(amp.fps_host_stream([batches[0]], P + 1))

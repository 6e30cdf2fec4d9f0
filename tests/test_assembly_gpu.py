"""Device-side batch assembly + augmentation (amp_assemble_windows_f32) against the restated host code of the reference loop
(oracle/train_loop_oracle.py: shuffle_clusters, rotate_point_cloud_z, shuffle_data; pinned bit-for-bit to the unmodified
train_loop by tests/test_reference_scripts.py): same seeds -> bit-identical encoder inputs and loss targets."""
import numpy as np
import pytest
import torch

from oracle import train_loop_oracle as tlo

pytestmark = pytest.mark.gpu


def _reference_windows(pc_clusters, targets, train):
    """Lines 390-408 + 421 of train_pointnet-attention.py, verbatim semantics, on the host."""
    pc_clusters, targets = tlo.shuffle_clusters(pc_clusters, targets)
    r_angle = np.random.uniform() * 2 * np.pi
    xs, ts = [], []
    for w in range(pc_clusters.shape[3]):
        in_points = pc_clusters[:, :, :, w].numpy().copy()
        targets_w = targets[:, :, w]
        if train:
            in_points[:, :, :3] = tlo.rotate_point_cloud_z(in_points[:, :, :3], rotation_angle=r_angle)
            in_points, targets_w, _ = tlo.shuffle_data(in_points, targets_w)
        xs.append(torch.Tensor(in_points)); ts.append(torch.LongTensor(targets_w))
    return torch.stack(xs, 0), torch.cat(ts, dim=1)


@pytest.mark.parametrize("n_samples,seed,train", [(3, 5, True), (2, 6, False), (5, 7, True)])
def test_assembled_windows_equal_the_reference_loop(amp, cuda, n_samples, seed, train):
    tlo.seed_all(seed)
    pc_clusters, targets, _, _ = tlo.collate_seq_padd(tlo.synthetic_samples(n_samples, seed))
    tlo.seed_all(seed + 100)
    x_ref, t_ref = _reference_windows(pc_clusters.clone(), targets.clone(), train)
    tlo.seed_all(seed + 100)
    n0 = amp._lib.launch_count()
    x, t = amp.assemble_windows(pc_clusters.pin_memory(), targets.pin_memory(), train=train, device=cuda)
    assert amp._lib.launch_count() == n0 + 1                       # one kernel for the whole batch
    assert tuple(x.shape) == (9, n_samples, 2048, 9) and tuple(t.shape) == (n_samples, 9 * 2048)
    assert torch.equal(x.cpu(), x_ref)                             # bit-exact, rotation included (float64 arithmetic like numpy's)
    assert torch.equal(t.cpu(), t_ref)
    assert x[3].is_contiguous()                                    # every window is a ready encoder input


def test_assemble_refuses_cpu(amp):
    with pytest.raises(RuntimeError, match="CUDA"):
        amp.assemble_windows(torch.zeros(1, 8, 9, 2), None, device="cpu")

"""CPU-side restatement of the callers of the hot path -- TEST INFRASTRUCTURE ONLY.

The reference's batch assembly and per-step logic stay the reference's own Python (kept verbatim); this file
restates them only because /root/reference does not exist on the GPU box, so that the `-m gpu` tests can drive the drop-in
modules exactly as the untouched scripts do:

  collate_seq_padd()      pointNet/collate_fns.py:4-55
  shuffle_clusters()      utils/utils.py:620-632
  rotate_point_cloud_z()  utils/utils.py:582-604
  shuffle_data()          utils/utils.py:607-617
  train_loop()            pointNet/self-attention/train_pointnet-attention.py:337-475 (segmentation task)

Every function consumes the random generators (torch, random, numpy) in the same order as the reference, so under the
same seeds it yields the same tensors. PINNED: tests/test_reference_scripts.py runs the UNMODIFIED train_loop (loaded from
the reference tree with importlib) and this restatement with the reference modules on the same seeded synthetic batch and
requires bit-identical losses, predictions and parameter updates; tests/golden/train_loop_reference.npz (made by
oracle/make_golden_loop.py from the unmodified loop) carries the result to the GPU box.
"""
import random

import numpy as np
import torch
import torch.nn.functional as F

GLOBAL_FEAT_SIZE = 256          # train_pointnet-attention.py:26
N_POINTS, MAX_WINDOWS = 2048, 9  # collate_fns.py:17-18


def collate_seq_padd(batch):
    """batch: list of (pc [n, 9, W] float, targets [n, W] int, filename, centroids [2, W] float)   (collate_fns.py:4-55)
    -> (batch_data [B, 2048, 9, 9], pad_targets [B, 2048, 9] (pad -1), filenames, pad_centroids [B, 9, 2])."""
    b_data = [torch.FloatTensor(t[0]) for t in batch]
    targets = [torch.LongTensor(t[1]) for t in batch]
    filenames = [t[2] for t in batch]
    centroids = [torch.FloatTensor(t[3]) for t in batch]
    batch_data, pad_targets, pad_centroids = [], [], []
    for i, (pc_w, target) in enumerate(zip(b_data, targets)):
        cent = centroids[i].unsqueeze(1)                                        # :32
        if pc_w.shape[0] < N_POINTS:                                            # :33-36
            rdm_list = torch.randint(0, pc_w.shape[0], (N_POINTS,))
            pc_w = pc_w[rdm_list, :, :]
            target = target[rdm_list, :]
        elif pc_w.shape[0] > N_POINTS:                                          # :38-41
            ix = random.sample(range(pc_w.shape[0]), N_POINTS)
            pc_w = pc_w[ix, :, :]
            target = target[ix, :]
        p1d = (0, MAX_WINDOWS - pc_w.shape[2])                                  # :42
        batch_data.append(F.pad(pc_w, p1d, "replicate"))                       # trailing windows replicate the last real one
        pad_targets.append(F.pad(target, p1d, "constant", -1))
        pad_centroids.append(F.pad(cent, p1d, "replicate"))
    batch_data = torch.stack(batch_data, dim=0)
    pad_targets = torch.stack(pad_targets, dim=0)
    pad_centroids = torch.stack(pad_centroids, dim=0).view(-1, MAX_WINDOWS, 2)  # :50-51 (a view, not a transpose: quirk 5)
    return batch_data, pad_targets, filenames, pad_centroids


def shuffle_clusters(data, labels):
    idx = np.arange(labels.shape[2])
    np.random.shuffle(idx)
    return data[:, :, :, idx], labels[:, :, idx]


def rotate_point_cloud_z(batch_data, rotation_angle=None):
    if not rotation_angle:
        rotation_angle = np.random.uniform() * 2 * np.pi
    rotated = np.zeros(batch_data.shape, dtype=np.float32)
    for k in range(batch_data.shape[0]):
        c, s = np.cos(rotation_angle), np.sin(rotation_angle)
        rot = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]])
        rotated[k, ...] = np.dot(batch_data[k, ...].reshape((-1, 3)), rot)     # float64 product, rounded to float32 on store
    return rotated


def shuffle_data(data, labels):
    idx = np.arange(labels.shape[1])
    np.random.shuffle(idx)
    return data[:, idx, :], labels[:, idx], idx


def train_loop(data, optimizer_pointnet, optimizer_att, ce_loss, pointnet, att_net, device, train=True):
    """One step of the segmentation task (train_pointnet-attention.py:337-475); returns (metrics, targets_pc, preds, logits)."""
    metrics = {}
    pc_clusters, targets, filenames, centroids = data
    centroids = centroids.to(device)
    batch_size = pc_clusters.shape[0]
    n_clusters = pc_clusters.shape[3]
    optimizer_pointnet.zero_grad()                                              # :372-373
    optimizer_att.zero_grad()
    if train:
        pointnet = pointnet.train(); att_net = att_net.train()
    else:
        pointnet = pointnet.eval(); att_net = att_net.eval()
    np_cluster = []
    lo_feats = torch.FloatTensor().to(device)
    gl_feats = torch.FloatTensor().to(device)
    targets_pc = torch.LongTensor().to(device)
    pc_clusters, targets = shuffle_clusters(pc_clusters, targets)              # :390
    r_angle = np.random.uniform() * 2 * np.pi                                   # :393
    feat_transform = None
    for w in range(n_clusters):                                                 # :396
        in_points = pc_clusters[:, :, :, w].numpy()
        targets_w = targets[:, :, w]
        if train:
            in_points[:, :, :3] = rotate_point_cloud_z(in_points[:, :, :3], rotation_angle=r_angle)   # :403
            in_points, targets_w, ix = shuffle_data(in_points, targets_w)                               # :405
        targets_w = torch.LongTensor(targets_w).to(device)
        in_points = torch.Tensor(in_points).to(device)                          # :408
        local_global_features, feat_transform = pointnet(in_points)             # :410
        local_feat = local_global_features[:, :, -64:]
        global_feat = local_global_features[:, 0, :-64].view(-1, 1, GLOBAL_FEAT_SIZE)
        np_cluster.append(local_feat.shape[1])
        lo_feats = torch.cat((lo_feats, local_feat), dim=1)
        gl_feats = torch.cat((gl_feats, global_feat), dim=1)
        targets_pc = torch.cat((targets_pc, targets_w), dim=1)
    targets_mask = targets_pc.view(batch_size, -1, n_clusters)                  # :428 (quirk 3: mixes windows, mask all False)
    mask = torch.where(targets_mask != -1, torch.zeros_like(targets_mask, dtype=torch.bool),
                       torch.ones_like(targets_mask, dtype=torch.bool))
    mask = torch.all(mask, 1)
    gl_feats = torch.transpose(gl_feats, 0, 1)
    logits, _ = att_net(gl_feats, lo_feats, centroids, np_cluster, mask)        # :435
    metrics['ce_loss'] = ce_loss(logits, targets_pc).view(-1, 1)                # :445
    targets_pc = targets_pc.detach().cpu()
    probs = F.log_softmax(logits.detach().cpu(), dim=1)                         # :449-450
    preds = torch.LongTensor(probs.data.max(1)[1])
    identity = torch.eye(feat_transform.shape[-1]).to(device)                   # :463-464
    metrics['reg_loss'] = torch.norm(identity - torch.bmm(feat_transform, feat_transform.transpose(2, 1)))
    if train:
        metrics['loss'] = metrics['ce_loss'] + 0.001 * metrics['reg_loss']      # :466-470
        metrics['loss'].backward()
        optimizer_pointnet.step()
        optimizer_att.step()
    else:
        metrics['loss'] = metrics['ce_loss']
    return metrics, targets_pc, preds, logits.detach()


def synthetic_samples(n_samples, seed, dims=9):
    """Samples as LidarKmeansDataset.__getitem__ yields them (datasets.py:322-370): (pc [n, 9, W], targets [n, W], name,
    centroids [2, W]) with different window counts and point counts (so that collate pads / resamples)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_samples):
        W = int(rng.integers(2, 10))
        n = int(rng.choice([1500, 2048, 2600]))
        pc = rng.random((n, dims, W), dtype=np.float32)
        pc[:, :2, :] = pc[:, :2, :] * 2 - 1
        pc[:, 2, :] *= 0.3
        # windows of a sample differ in scale / offset (well-conditioned BatchNorm over the few clouds of a batch)
        pc = pc * (0.3 + 0.7 * rng.random((1, dims, W), dtype=np.float32)) + 0.2 * rng.standard_normal((1, dims, W)).astype(np.float32)
        tg = rng.integers(0, 5, (n, W)).astype(np.int64)
        cent = pc[:, :2, :].mean(0)
        out.append((pc, tg, "sample_%d" % i, cent))
    return out


def seed_all(seed):
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)

"""Device-side batch assembly and augmentation for the training loop.

`assemble_windows` replaces lines 390-408 of pointNet/self-attention/train_pointnet-attention.py (shuffle_clusters, the
per-window numpy rotate_point_cloud_z + shuffle_data, W separate `.to(device)` copies) by ONE host-to-device copy of the
collated batch and ONE kernel. The random draws are made on the host with numpy in the reference's order
(utils/utils.py:620-632 shuffle over W; train_...:393 angle; :607-617 one shuffle over N per window), so a seeded run sees
the tensors the reference loop would build.
"""
import numpy as np
import torch

from . import _lib


def draw_augmentation(n_windows, n_points, train=True):
    """(cluster_perm [W], angle, point_perm [W, N]) drawn from numpy's global generator exactly as train_loop does.
    The reference shuffles the clusters in eval mode too (train_pointnet-attention.py:390) and rotates / shuffles points
    only when training (:399-405)."""
    cperm = np.arange(n_windows)
    np.random.shuffle(cperm)                                   # shuffle_clusters
    angle = np.random.uniform() * 2 * np.pi                    # r_angle (drawn in eval mode as well)
    pperm = np.tile(np.arange(n_points, dtype=np.int32), (n_windows, 1))
    if train:
        for w in range(n_windows):
            idx = np.arange(n_points)
            np.random.shuffle(idx)                             # shuffle_data of window w
            pperm[w] = idx
    return cperm.astype(np.int32), float(angle), pperm


def assemble_windows(pc_clusters, targets=None, train=True, device="cuda", augmentation=None):
    """pc_clusters [B, N, D, W] float32 and targets [B, N, W] int64 as collate_seq_padd returns them (host, ideally pinned, or
    already on the device). Returns (x [W, B, N, D] float32 on the device, targets_pc [B, W * N] int64 on the device | None):
    `x[w]` is the encoder input of window w (train_pointnet-attention.py:407-410) and `targets_pc` the tensor the loss takes
    (:421, :445)."""
    dev = torch.device(device) if not pc_clusters.is_cuda else pc_clusters.device
    if dev.type != "cuda":
        raise RuntimeError("ampnet_b200: assemble_windows runs on a CUDA device (no CPU fallback)")
    if pc_clusters.dim() != 4 or pc_clusters.dtype != torch.float32:
        raise ValueError("pc_clusters must be float32 [B, N, D, W]")
    B, N, D, W = pc_clusters.shape
    cperm, angle, pperm = augmentation if augmentation is not None else draw_augmentation(W, N, train)
    pc = pc_clusters.to(dev, non_blocking=True).contiguous()
    tg = None
    if targets is not None:
        if tuple(targets.shape) != (B, N, W) or targets.dtype != torch.int64:
            raise ValueError("targets must be int64 [B, N, W]")
        tg = targets.to(dev, non_blocking=True).contiguous()
    d_c = torch.from_numpy(np.ascontiguousarray(cperm, dtype=np.int32)).to(dev, non_blocking=True)
    d_p = torch.from_numpy(np.ascontiguousarray(pperm, dtype=np.int32)).to(dev, non_blocking=True)
    x = torch.empty((W, B, N, D), dtype=torch.float32, device=dev)
    t = torch.empty((B, W * N), dtype=torch.int64, device=dev) if tg is not None else None
    # the reference's `if not rotation_angle` (utils.py:592) treats an angle of exactly 0 as "draw a new one"; the loop passes
    # its own r_angle, which is 0 with probability 0
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().amp_assemble_windows_f32(pc.data_ptr(), tg.data_ptr() if tg is not None else None, d_c.data_ptr(), d_p.data_ptr(),
                                                       B, N, D, W, 1 if train else 0, float(np.cos(angle)), float(np.sin(angle)),
                                                       x.data_ptr(), t.data_ptr() if t is not None else None, _lib.stream_ptr()))
    return x, t

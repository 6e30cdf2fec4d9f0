"""Parameter tables of the reference modules and a platform-independent synthetic state_dict -- TEST
INFRASTRUCTURE ONLY (also used by bench.py to give both arms the same random-init weights).

Shapes restate pointNet/model/pointnetAtt.py:10-26 (TransformationNet), :59-78 (BasePointNet with
point_dimension=3, global_feat_dim=256) and :160-174 (SegmentationWithAttention(256, 8, num_classes=5,
local_dim=64)); tests/test_oracle_pinned.py checks them key-for-key against the reference modules.
"""
import numpy as np
import torch


def _bn(pre, c):
    return [(pre + ".weight", (c,)), (pre + ".bias", (c,)), (pre + ".running_mean", (c,)),
            (pre + ".running_var", (c,)), (pre + ".num_batches_tracked", ())]


def _tnet(pre, din, dout):
    t = [(pre + "conv_1.weight", (64, din, 1)), (pre + "conv_2.weight", (128, 64, 1)),
         (pre + "conv_3.weight", (256, 128, 1))]
    for name, c in (("bn_1", 64), ("bn_2", 128), ("bn_3", 256), ("bn_4", 256), ("bn_5", 128)):
        t += _bn(pre + name, c)
    t += [(pre + "fc_1.weight", (256, 256)), (pre + "fc_2.weight", (128, 256)),
          (pre + "fc_3.weight", (dout * dout, 128)), (pre + "fc_3.bias", (dout * dout,))]
    return t


def encoder_shapes(point_dimension=3, global_feat_dim=256):
    t = _tnet("input_transform.", point_dimension, point_dimension) + _tnet("feature_transform.", 64, 64)
    t += [("conv_1.weight", (64, 9 + point_dimension, 1)), ("conv_2.weight", (64, 64, 1)),
          ("conv_3.weight", (64, 64, 1)), ("conv_4.weight", (128, 64, 1)), ("conv_5.weight", (128, 128, 1)),
          ("conv_6.weight", (global_feat_dim, 128, 1))]
    for name, c in (("bn_1", 64), ("bn_2", 64), ("bn_3", 64), ("bn_4", 128), ("bn_5", 128), ("bn_6", global_feat_dim)):
        t += _bn(name, c)
    return t


def seg_shapes(embed_dim=256, num_classes=5, local_dim=64):
    h = embed_dim // 2
    t = [("fc1.weight", (16, 2)), ("fc1.bias", (16,)), ("fc2.weight", (embed_dim, 16)), ("fc2.bias", (embed_dim,)),
         ("attention.in_proj_weight", (3 * embed_dim, embed_dim)), ("attention.in_proj_bias", (3 * embed_dim,)),
         ("attention.out_proj.weight", (embed_dim, embed_dim)), ("attention.out_proj.bias", (embed_dim,)),
         ("conv_2.weight", (h, local_dim + embed_dim, 1)), ("conv_2.bias", (h,)),
         ("conv_3.weight", (64, h, 1)), ("conv_3.bias", (64,)),
         ("conv_4.weight", (num_classes, 64, 1)), ("conv_4.bias", (num_classes,))]
    t += _bn("bn_2", h) + _bn("bn_3", 64)
    return t


def synthetic_state_dict(shapes, seed, trained_bn=True):
    """Deterministic (numpy PCG64) weights: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's default init.
    trained_bn=True also randomises BN affine parameters and running statistics (as after training),
    so eval-mode BatchNorm is exercised non-trivially; False gives fresh-module BN (1, 0, 0, 1)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in shapes:
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(7 if trained_bn else 0, dtype=torch.int64)
            continue
        if ".running_var" in name:
            v = rng.uniform(0.5, 1.5, shape) if trained_bn else np.ones(shape)
        elif ".running_mean" in name:
            v = rng.uniform(-0.2, 0.2, shape) if trained_bn else np.zeros(shape)
        elif "bn_" in name and name.endswith(".weight"):
            v = rng.uniform(0.7, 1.3, shape) * np.where(rng.random(shape) < 0.1, -1.0, 1.0) if trained_bn else np.ones(shape)
        elif "bn_" in name and name.endswith(".bias"):
            v = rng.uniform(-0.2, 0.2, shape) if trained_bn else np.zeros(shape)
        elif name.endswith("fc_3.bias"):
            v = rng.uniform(-0.05, 0.05, shape)
        elif name.endswith("bias"):
            v = rng.uniform(-0.1, 0.1, shape)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            b = 1.0 / np.sqrt(fan_in)
            v = rng.uniform(-b, b, shape)
        sd[name] = torch.from_numpy(np.asarray(v, dtype=np.float32))
    return sd


def conditioned_blocks(B, N, W, seed):
    """synthetic_blocks with clouds that differ from each other (per-cloud channel scales / offsets). With i.i.d. uniform
    blocks the T-Net BatchNorms over the B pooled rows divide by a near-zero spread, and every fp32 implementation, the
    reference included, is only good to ~1e-2 on the gradients; this variant is well conditioned, so bounds can be tight."""
    xs, _ = synthetic_blocks(B, N, W, seed)
    g = torch.Generator().manual_seed(1000 + seed)
    xs = [x * (0.15 + 0.85 * torch.rand(B, 1, 9, generator=g)) + 0.3 * torch.randn(B, 1, 9, generator=g) for x in xs]
    return xs, torch.stack([x[:, :, :2].mean(1) for x in xs], 1)


def synthetic_blocks(B, N, W, seed):
    """Synthetic ALS blocks as SURVEY 8(d): x,y ~ U[-1,1], z ~ U[0,0.3], six features ~ U[0,1];
    centroids = per-block mean(x, y). Returns (list of W tensors [B,N,9], centroids [B,W,2])."""
    rng = np.random.default_rng(seed)
    xs, cs = [], []
    for _ in range(W):
        x = rng.random((B, N, 9), dtype=np.float32)
        x[:, :, :2] = x[:, :, :2] * 2 - 1
        x[:, :, 2] *= 0.3
        xs.append(torch.from_numpy(x))
        cs.append(torch.from_numpy(x[:, :, :2].mean(1)))
    return xs, torch.stack(cs, 1)

"""NN golden vectors: outputs of the UNMODIFIED reference modules (pointNet/model/pointnetAtt.py) on
seeded synthetic inputs, with the synthetic state_dict of oracle/nn_params.py loaded by key. Called by
oracle/make_golden.py (needs /root/reference). Only inputs' seeds and outputs are stored; the weights
are regenerated from the seed at test time."""
import os

import numpy as np
import torch

from . import nn_params

CASES = {  # name: (B, N, W, seed, masked)
    "w1_b4_n256": (4, 256, 1, 11, False),
    "w3_b3_n128": (3, 128, 3, 12, True),
}


# >= 2048 rows per encoder call: the shapes at which the tensor-core layers (tc_layer_kernel / tc_wgrad_kernel) serve the
# forward AND the backward (the cases above stay below that threshold and exercise the CUDA-core kernels). Inputs are
# nn_params.conditioned_blocks; fixture: tests/golden/nn_reference_tc.npz
CASES_TC = {  # name: (B, N, W, seed)
    "tc_w1_b8_n512": (8, 512, 1, 21),
    "tc_w2_b4_n1024": (4, 1024, 2, 22),
}


def subsample_tc(g):
    """At most ~12 rows of a large gradient (every k-th row): the tensor-core fixture holds every parameter."""
    if g.size <= 2048:
        return g.copy()
    k = max(1, g.shape[0] // 12)
    return g[::k].copy()


def subsample(g):
    """Keep at most ~48 rows of a large gradient (every k-th row) so the fixture stays small."""
    if g.size <= 20000:
        return g.copy()
    k = max(1, g.shape[0] // 48)
    return g[::k].copy()


def build_reference(model, seed, trained_bn=True):
    enc = model.BasePointNet(point_dimension=3, return_local_features=True, global_feat_dim=256, device="cpu")
    seg = model.SegmentationWithAttention(256, 8, num_classes=5, local_dim=64, dropout=0.0, device="cpu")
    sd_e = nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed, trained_bn)
    sd_s = nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1, trained_bn)
    enc.load_state_dict(sd_e, strict=True)
    seg.load_state_dict(sd_s, strict=True)
    return enc, seg, sd_e, sd_s


def run_reference(enc, seg, xs, centroids, mask, train, targets=None):
    """The encoder loop + head exactly as train_pointnet-attention.py:396-435 drives the modules."""
    enc.train(train); seg.train(train)
    lo = torch.FloatTensor(); gl = torch.FloatTensor(); npc = []
    ft = None
    for xw in xs:
        out, ft = enc(xw)
        local_feat = out[:, :, -64:]
        global_feat = out[:, 0, :-64].view(-1, 1, 256)
        npc.append(local_feat.shape[1])
        lo = torch.cat((lo, local_feat), dim=1)
        gl = torch.cat((gl, global_feat), dim=1)
    gl = torch.transpose(gl, 0, 1)
    logits, _ = seg(gl, lo, centroids, npc, mask)
    res = {"logits": logits, "ft": ft, "enc_out_last": out}
    if targets is not None:
        ce = torch.nn.CrossEntropyLoss(weight=torch.FloatTensor([1, 2, 2, 1, 1]), reduction="mean", ignore_index=-1)
        eye = torch.eye(64)
        loss = ce(logits, targets) + 0.001 * torch.norm(eye - torch.bmm(ft, ft.transpose(2, 1)))
        res["loss"] = loss
    return res


def make(model, gold_dir):
    out = {}
    for name, (B, N, W, seed, masked) in CASES.items():
        enc, seg, _, _ = build_reference(model, seed)
        xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
        mask = None
        if masked:
            mask = torch.zeros(B, W, dtype=torch.bool); mask[0, W - 1] = True
        with torch.no_grad():
            r = run_reference(enc, seg, xs, cent, mask, train=False)
        out[name + "__eval_logits"] = r["logits"].numpy()
        out[name + "__eval_ft"] = r["ft"].numpy()
        out[name + "__eval_enc_out_last"] = r["enc_out_last"].numpy()[:, ::37, :].copy()   # rows 0, 37, 74, ...
        # training mode (dropout 0): loss, a few gradients, BN running stats after the step's forward
        enc, seg, _, _ = build_reference(model, seed)
        tg = torch.from_numpy(np.random.default_rng(seed).integers(-1, 5, (B, N * W)).astype(np.int64))
        r = run_reference(enc, seg, xs, cent, mask, train=True, targets=tg)
        r["loss"].backward()
        out[name + "__train_logits"] = r["logits"].detach().numpy()
        out[name + "__train_loss"] = r["loss"].detach().numpy()
        out[name + "__targets"] = tg.numpy()
        for mod, tag in ((enc, "enc"), (seg, "seg")):
            for k, p in mod.named_parameters():
                if k in ("conv_1.weight", "conv_6.weight", "input_transform.fc_3.bias", "feature_transform.conv_2.weight",
                         "feature_transform.fc_3.weight", "bn_3.weight", "bn_3.bias", "conv_2.weight", "conv_4.bias",
                         "attention.in_proj_weight", "fc1.weight", "input_transform.conv_1.weight"):
                    g = p.grad.numpy()
                    out["%s__grad_%s_%s" % (name, tag, k)] = subsample(g)
        out[name + "__train_rm_bn_6"] = enc.bn_6.running_mean.numpy().copy()
        out[name + "__train_rv_bn_1"] = enc.bn_1.running_var.numpy().copy()
        out[name + "__train_rv_seg_bn_2"] = seg.bn_2.running_var.numpy().copy()
    path = os.path.join(gold_dir, "nn_reference.npz")
    np.savez_compressed(path, **out)
    print("nn_reference.npz: %d arrays, %.1f KiB" % (len(out), os.path.getsize(path) / 1024))
    make_tc(model, gold_dir)


def make_tc(model, gold_dir):
    """Training step of the UNMODIFIED reference modules at tensor-core shapes: logits (every 16th point), loss, EVERY
    parameter gradient (subsampled rows), and the gradients of the two tensors that cross from the encoder to the head."""
    out = {}
    for name, (B, N, W, seed) in CASES_TC.items():
        enc, seg, _, _ = build_reference(model, seed)
        xs, cent = nn_params.conditioned_blocks(B, N, W, seed)
        tg = torch.from_numpy(np.random.default_rng(seed).integers(-1, 5, (B, N * W)).astype(np.int64))
        enc.train(True); seg.train(True)
        lo = torch.FloatTensor(); gl = torch.FloatTensor(); npc = []
        for xw in xs:                                                  # train_pointnet-attention.py:396-417
            o, ft = enc(xw)
            local_feat = o[:, :, -64:]
            global_feat = o[:, 0, :-64].view(-1, 1, 256)
            npc.append(local_feat.shape[1])
            lo = torch.cat((lo, local_feat), dim=1)
            gl = torch.cat((gl, global_feat), dim=1)
        gl = torch.transpose(gl, 0, 1)
        lo.retain_grad(); gl.retain_grad()
        logits, _ = seg(gl, lo, cent, npc, None)
        ce = torch.nn.CrossEntropyLoss(weight=torch.FloatTensor([1, 2, 2, 1, 1]), reduction="mean", ignore_index=-1)
        loss = ce(logits, tg) + 0.001 * torch.norm(torch.eye(64) - torch.bmm(ft, ft.transpose(2, 1)))
        loss.backward()
        out[name + "__train_logits"] = logits.detach().numpy()[:, :, ::16].copy()
        out[name + "__train_loss"] = loss.detach().numpy()
        out[name + "__targets"] = tg.numpy().astype(np.int8)
        out[name + "__dlo"] = lo.grad.numpy()[:, ::101, :].copy()
        out[name + "__dgl"] = gl.grad.numpy().copy()
        for mod, tag in ((enc, "enc"), (seg, "seg")):
            for k, p in mod.named_parameters():
                out["%s__grad_%s_%s" % (name, tag, k)] = subsample_tc(p.grad.numpy()).astype(np.float32)
    path = os.path.join(gold_dir, "nn_reference_tc.npz")
    np.savez_compressed(path, **out)
    print("nn_reference_tc.npz: %d arrays, %.1f KiB" % (len(out), os.path.getsize(path) / 1024))

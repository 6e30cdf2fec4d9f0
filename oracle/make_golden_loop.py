"""Golden vectors of the UNMODIFIED reference training step (needs /root/reference; this container only).

    python -m oracle.make_golden_loop

Loads pointNet/self-attention/train_pointnet-attention.py (the file name is not importable: importlib by path, with the
missing third-party imports stubbed as in oracle/ref_import.py), builds the reference modules with the synthetic state_dict
of oracle/nn_params.py, collates a seeded synthetic batch with the reference's own collate_seq_padd and runs ONE train_loop
step and ONE eval step on the CPU. Stored (tests/golden/train_loop_reference.npz): losses, predictions, and a sample of the
parameters after the two Adam steps. The inputs are regenerated from the seeds at test time (oracle/train_loop_oracle.py).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import nn_params, ref_import, train_loop_oracle as tlo  # noqa: E402

CASES = {"b3": (3, 101)}      # name: (samples in the batch, seed)
LR = 1e-3


def load_train_script():
    """The unmodified train script as a module (its __main__ block does not run)."""
    model, uu, coll = ref_import.load()
    for name, attrs in (("prettytable", {"PrettyTable": object}), ("codecarbon", {"track_emissions": lambda f=None, **k: f}),
                        ("progressbar", {}), ("alive_progress", {}), ("torchsummary", {"summary": None}), ("laspy", {})):
        if name not in sys.modules:
            m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m
    for name, attr in (("matplotlib.colors", "ListedColormap"), ("matplotlib.lines", "Line2D")):   # stubs of ref_import: the plot
        if name in sys.modules and not hasattr(sys.modules[name], attr):                           # helpers only need the names
            setattr(sys.modules[name], attr, object)
    try:
        import torch.utils.tensorboard  # noqa: F401
    except Exception:
        m = types.ModuleType("torch.utils.tensorboard"); m.SummaryWriter = object; sys.modules["torch.utils.tensorboard"] = m
    path = os.path.join(ref_import.REFERENCE_ROOT, "pointNet", "self-attention", "train_pointnet-attention.py")
    spec = importlib.util.spec_from_file_location("ref_train_pointnet_attention", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, model, coll


def build(model, seed):
    enc = model.BasePointNet(point_dimension=3, return_local_features=True, global_feat_dim=256, device="cpu")
    seg = model.SegmentationWithAttention(256, 8, num_classes=5, local_dim=64, dropout=0.0, device="cpu")
    enc.load_state_dict(nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed), strict=True)
    seg.load_state_dict(nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1), strict=True)
    return enc, seg


def run_case(train_loop, collate, enc, seg, n_samples, seed, device="cpu", **kw):
    """One eval step (validation loop, :245-274) then one training step (:180-219) on the same batch, as train_att() drives
    them (:127-149). Eval first: after the first Adam step every weight has moved by ~lr * sign(gradient), so a later
    forward would compare update directions of near-zero gradients rather than forward arithmetic."""
    opt_e = torch.optim.Adam(enc.parameters(), lr=LR)
    opt_s = torch.optim.Adam(seg.parameters(), lr=LR)
    ce = torch.nn.CrossEntropyLoss(weight=torch.FloatTensor([1, 2, 2, 1, 1]).to(device), reduction="mean", ignore_index=-1)
    if "task" not in kw:                 # the restated loop takes the device as an argument (the reference's is a module global)
        kw = dict(kw, device=device)
    tlo.seed_all(seed)
    data = collate(tlo.synthetic_samples(n_samples, seed))
    out = {}
    tlo.seed_all(seed + 2)
    with torch.no_grad():
        r = train_loop(data, opt_e, opt_s, ce, enc, seg, **kw, train=False)
    out["eval"] = r
    tlo.seed_all(seed + 1)
    r = train_loop(data, opt_e, opt_s, ce, enc, seg, **kw, train=True)
    out["train"] = r
    return out


SAMPLED = ("conv_1.weight", "conv_6.weight", "bn_3.bias", "feature_transform.fc_3.bias", "input_transform.conv_2.weight")
SAMPLED_SEG = ("conv_2.weight", "conv_4.weight", "attention.in_proj_bias", "fc1.weight", "bn_2.weight")


def main():
    mod, model, coll = load_train_script()
    mod.device = "cpu"
    gold = {}
    for name, (n_samples, seed) in CASES.items():
        enc, seg = build(model, seed)
        res = run_case(mod.train_loop, coll.collate_seq_padd, enc, seg, n_samples, seed, task="segmentation")
        for phase in ("train", "eval"):
            metrics, targets_pc, preds, _ = res[phase]
            gold["%s__%s_ce" % (name, phase)] = metrics["ce_loss"].detach().numpy().reshape(())
            gold["%s__%s_reg" % (name, phase)] = metrics["reg_loss"].detach().numpy().reshape(())
            gold["%s__%s_preds" % (name, phase)] = preds.numpy().astype(np.int8)
            gold["%s__%s_targets" % (name, phase)] = targets_pc.numpy().astype(np.int8)
        for k in SAMPLED:
            gold["%s__param_enc_%s" % (name, k)] = dict(enc.named_parameters())[k].detach().numpy().reshape(-1)[::7].copy()
        for k in SAMPLED_SEG:
            gold["%s__param_seg_%s" % (name, k)] = dict(seg.named_parameters())[k].detach().numpy().reshape(-1)[::7].copy()
        gold["%s__rm_bn_6" % name] = enc.bn_6.running_mean.numpy().copy()
        gold["%s__nbt" % name] = np.array(int(enc.bn_1.num_batches_tracked))
    path = os.path.join(ROOT, "tests", "golden", "train_loop_reference.npz")
    np.savez_compressed(path, **gold)
    print("train_loop_reference.npz: %d arrays, %.1f KiB" % (len(gold), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()

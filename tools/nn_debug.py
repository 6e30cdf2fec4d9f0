"""GPU debugging aid: runs the drop-in modules and the CPU oracle on the same seeded inputs and prints the
relative error of every output, gradient and BatchNorm buffer. Not part of the product path.

    python tools/nn_debug.py [B N W] [--train]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ampnet_b200 as amp  # noqa: E402
from oracle import nn_oracle, nn_params  # noqa: E402


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def relnorm(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build(seed, dev, dropout=0.0):
    enc = amp.BasePointNet(point_dimension=3, return_local_features=True, global_feat_dim=256, device=dev)
    seg = amp.SegmentationWithAttention(256, 8, num_classes=5, local_dim=64, dropout=dropout, device=dev)
    sd_e = nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed)
    sd_s = nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1)
    enc.load_state_dict(sd_e, strict=True)
    seg.load_state_dict(sd_s, strict=True)
    return enc.to(dev), seg.to(dev), sd_e, sd_s


def run_gpu(enc, seg, xs, cent, mask, dev):
    lo, gl, npc, ft, out = [], [], [], None, None
    for xw in xs:
        out, ft = enc(xw.to(dev))
        lo.append(out[:, :, -64:])
        gl.append(out[:, 0, :-64].view(-1, 1, 256))
        npc.append(xw.shape[1])
    lo = torch.cat(lo, 1)
    gl = torch.transpose(torch.cat(gl, 1), 0, 1)
    logits, _ = seg(gl, lo, cent.to(dev), npc, None if mask is None else mask.to(dev))
    return logits, ft, out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    B, N, W = (int(a) for a in args) if len(args) == 3 else (4, 256, 2)
    train = "--train" in sys.argv
    dev = torch.device("cuda:0")
    seed = 21
    enc, seg, sd_e, sd_s = build(seed, dev)
    xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
    if "--hetero" in sys.argv:      # clouds that differ from each other: BatchNorm over the B rows is well conditioned
        g = torch.Generator().manual_seed(9)
        xs = [x * (0.15 + 0.85 * torch.rand(B, 1, 9, generator=g)) + 0.3 * torch.randn(B, 1, 9, generator=g) for x in xs]
        cent = torch.stack([x[:, :, :2].mean(1) for x in xs], 1)
    mask = None
    enc.train(train); seg.train(train)
    if not train:
        logits, ft, out = run_gpu(enc, seg, xs, cent, mask, dev)
        o_logits, o_ft = nn_oracle.forward_windows(sd_e, sd_s, xs, cent, mask, training=False)
        o_out, _ = nn_oracle.base_pointnet(sd_e, xs[-1])
        print("eval  logits rel %.3e | ft rel %.3e | out(global) rel %.3e | out(local) rel %.3e" % (
            rel(logits, o_logits), rel(ft, o_ft), rel(out[:, :, :256], o_out[:, :, :256]), rel(out[:, :, 256:], o_out[:, :, 256:])))
        agree = (logits.argmax(1).cpu() == o_logits.argmax(1)).float().mean().item()
        print("eval  argmax agreement %.5f" % agree)
        return
    for sd in (sd_e, sd_s):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    tg = torch.randint(-1, 5, (B, N * W), generator=torch.Generator().manual_seed(3))
    logits, ft, out = run_gpu(enc, seg, xs, cent, mask, dev)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=dev), ignore_index=-1)
    loss = ce(logits, tg.to(dev)) + 0.001 * torch.norm(torch.eye(64, device=dev) - torch.bmm(ft, ft.transpose(2, 1)))
    loss.backward()
    st_e, st_s = {}, {}
    o_logits, o_ft = nn_oracle.forward_windows(sd_e, sd_s, xs, cent, mask, training=True, stats_enc=st_e, stats_seg=st_s)
    o_loss, _, _ = nn_oracle.train_step_loss(o_logits, tg, o_ft)
    o_loss.backward()
    # float64 run of the same oracle = "truth": shows how much of the difference is fp32 noise of the reference itself
    sd_e64 = {k: (v.detach().double() if v.is_floating_point() else v.clone()) for k, v in nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed).items()}
    sd_s64 = {k: (v.detach().double() if v.is_floating_point() else v.clone()) for k, v in nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1).items()}
    for sd in (sd_e64, sd_s64):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    t_logits, t_ft = nn_oracle.forward_windows(sd_e64, sd_s64, [x.double() for x in xs], cent.double(), mask, training=True)
    t_loss, _, _ = nn_oracle.train_step_loss(t_logits, tg, t_ft)
    t_loss.backward()
    print("vs f64: logits ours %.3e oracle32 %.3e | ft ours %.3e oracle32 %.3e" % (
        rel(logits, t_logits), rel(o_logits, t_logits), rel(ft, t_ft), rel(o_ft, t_ft)))
    print("train logits rel %.3e | ft rel %.3e | loss %.6f vs %.6f" % (rel(logits, o_logits), rel(ft, o_ft), float(loss), float(o_loss)))
    for tag, mod, sd in (("enc", enc, sd_e), ("seg", seg, sd_s)):
        for k, p in mod.named_parameters():
            g = p.grad
            og = sd[k].grad
            tg64 = (sd_e64 if tag == "enc" else sd_s64)[k].grad
            print("  grad %s.%-36s vs oracle32 %.3e | vs f64: ours %.3e oracle32 %.3e  (|g| %.3e)" % (
                tag, k, relnorm(g, og) if g is not None else float("nan"), relnorm(g, tg64), relnorm(og, tg64), float(og.norm())))
        for k, b in mod.named_buffers():
            if "running" in k:
                print("  buf  %s.%-36s rel %.3e" % (tag, k, rel(b, sd[k])))


if __name__ == "__main__":
    main()

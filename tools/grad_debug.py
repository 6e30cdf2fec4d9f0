"""Per-parameter gradient error of the CUDA path vs the float64 oracle (the body of test_all_gradients_match_oracle)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
import test_nn_gpu as T
from oracle import nn_oracle, nn_params
cuda = torch.device("cuda:0")
B, N, W, seed = 8, 192, 2, int(os.environ.get("SEED", "31"))
enc, seg, sd_e, sd_s = T._build(amp, seed, cuda)
xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
g = torch.Generator().manual_seed(9)
xs = [x * (0.15 + 0.85 * torch.rand(B, 1, 9, generator=g)) + 0.3 * torch.randn(B, 1, 9, generator=g) for x in xs]
cent = torch.stack([x[:, :, :2].mean(1) for x in xs], 1)
enc.train(); seg.train()
logits, ft, _ = T._run(enc, seg, xs, cent, None, cuda)
tg = torch.randint(-1, 5, (B, N * W), generator=torch.Generator().manual_seed(3))
ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=cuda), ignore_index=-1)
loss = ce(logits, tg.to(cuda)) + 0.001 * torch.norm(torch.eye(64, device=cuda) - torch.bmm(ft, ft.transpose(2, 1)))
loss.backward()
for sd in (sd_e, sd_s):
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k: v.requires_grad_(True)
o_logits, o_ft = nn_oracle.forward_windows(sd_e, sd_s, xs, cent, None, training=True, stats_enc={}, stats_seg={})
nn_oracle.train_step_loss(o_logits, tg, o_ft)[0].backward()
sd_e64 = {k: (v.detach().double() if v.is_floating_point() else v.clone()) for k, v in sd_e.items()}
sd_s64 = {k: (v.detach().double() if v.is_floating_point() else v.clone()) for k, v in sd_s.items()}
for sd in (sd_e64, sd_s64):
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k: v.requires_grad_(True)
t_logits, t_ft = nn_oracle.forward_windows(sd_e64, sd_s64, [x.double() for x in xs], cent.double(), None, training=True)
nn_oracle.train_step_loss(t_logits, tg, t_ft)[0].backward()
print("logits: ours %.3e  ref32 %.3e" % (T._rel(logits, t_logits), T._rel(o_logits, t_logits)))
for mod, sd, sd64, tag in ((enc, sd_e, sd_e64, "enc"), (seg, sd_s, sd_s64, "seg")):
    for k, p in mod.named_parameters():
        tg64 = sd64[k].grad
        if float(tg64.norm()) < 1e-6:
            print("%s %-40s zero-grad ours norm %.2e" % (tag, k, float(p.grad.norm()))); continue
        ours, ref32 = T._relnorm(p.grad, tg64), T._relnorm(sd[k].grad, tg64)
        flag = "" if ours < 2.5 * ref32 + 1e-4 and ours < 1e-3 else "   <<<<"
        print("%s %-40s ours %.3e  ref32 %.3e%s" % (tag, k, ours, ref32, flag))
g1 = seg.bn_3.bias.grad.double().cpu(); g2 = sd_s64["bn_3.bias"].grad
e = (g1 - g2).abs() / g2.abs().max()
print("seg bn_3.bias per-channel rel err: max %.2e median %.2e, #channels > 1e-5: %d of %d" % (e.max(), e.median(), int((e > 1e-5).sum()), e.numel()))
g1 = enc.bn_6.bias.grad.double().cpu(); g2 = sd_e64["bn_6.bias"].grad
e = (g1 - g2).abs() / g2.abs().max()
print("enc bn_6.bias per-channel rel err: max %.2e median %.2e, #channels > 1e-4: %d of %d" % (e.max(), e.median(), int((e > 1e-4).sum()), e.numel()))

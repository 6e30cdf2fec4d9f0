"""Uninitialised-read hunt: fill the caching allocator's free blocks with NaN before forward and before backward."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
from oracle import nn_params
import test_nn_gpu as T
cuda = torch.device("cuda:0")
B, N, seed = 4, 512, 81
xs, cent = nn_params.synthetic_blocks(B, N, 2, seed)
def poison(val):
    blocks = [torch.full((n,), val, device=cuda) for n in (1 << 26, 1 << 24, 1 << 22, 1 << 20, 1 << 18, 1 << 16, 1 << 14, 1 << 12) for _ in range(4)]
    del blocks
def run(val):
    enc, seg, _, _ = T._build(amp, seed, cuda)
    enc.train(); seg.train()
    params = list(enc.parameters()) + list(seg.parameters())
    names = [n for n, _ in enc.named_parameters()] + [n for n, _ in seg.named_parameters()]
    if val is not None: poison(val)
    logits, ft, _ = T._run(enc, seg, xs, cent, None, cuda)
    loss = logits.square().mean() + 0.01 * ft.square().mean()
    if val is not None: poison(val)
    loss.backward()
    return names, [p.grad.clone() for p in params], logits.detach().clone()
names, g0, l0 = run(None)
for val in (float("nan"), 1e3):
    _, g1, l1 = run(val)
    print("poison", val, "logits rel", T._rel(l0, l1))
    bad = [(n, T._rel(a, b)) for n, a, b in zip(names, g0, g1) if not (T._rel(a, b) < 1e-6)]
    print(len(bad), "parameters differ", bad[:6])

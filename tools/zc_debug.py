import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
from oracle import nn_params
import test_nn_gpu as T
cuda = torch.device("cuda:0")
B, N, seed = 4, 512, 81
xs, cent = nn_params.synthetic_blocks(B, N, 2, seed)
def run(zero_copy, passes):
    enc, seg, _, _ = T._build(amp, seed, cuda)
    enc.train(); seg.train()
    params = list(enc.parameters()) + list(seg.parameters())
    names = [n for n, _ in enc.named_parameters()] + [n for n, _ in seg.named_parameters()]
    red = amp.GradAllReduce(params, world=1, zero_copy=True) if zero_copy else None
    for _ in range(passes):
        enc.load_state_dict(nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed))
        seg.load_state_dict(nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1))
        if not zero_copy:
            for p in params: p.grad = None
        logits, ft, _ = T._run(enc, seg, xs, cent, None, cuda)
        (logits.square().mean() + 0.01 * ft.square().mean()).backward()
    return names, [p.grad.clone() for p in params]
names, g1 = run(False, 1)
_, g2 = run(False, 2)
_, z1 = run(True, 1)
_, z2 = run(True, 2)
for n, a, b, c, d in zip(names, g1, g2, z1, z2):
    r = [T._rel(a, x) for x in (b, c, d)]
    if max(r) > 1e-6: print("%-40s %s numel %d  normal2 %.2e  zc1 %.2e  zc2 %.2e" % (n, tuple(a.shape), a.numel(), *r))
print("done")

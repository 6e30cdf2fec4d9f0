import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
import bench_nn as nb
dev = torch.device("cuda:0")
enc, seg = nb.build_modules(amp, dev, dropout=0.0); enc.train(); seg.train()
x_np, c_np, t_np = nb.synthetic_blocks(0)
x, cent, tg = torch.from_numpy(x_np).to(dev), torch.from_numpy(c_np).to(dev), torch.from_numpy(t_np).to(dev)
for i in range(2):
    if i == 1: print("==== second step", file=sys.stderr, flush=True)
    logits, ft = nb.forward_pass(enc, seg, x, cent)
    loss = torch.nn.functional.cross_entropy(logits, tg)
    loss.backward(); torch.cuda.synchronize()

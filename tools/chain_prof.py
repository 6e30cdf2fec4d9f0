import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
nb = importlib.import_module("3d-semantic-segmentation-amp-net_b200.nn_bench")
dev = torch.device("cuda:0")
enc, seg = nb.build_modules(amp, dev); enc.eval(); seg.eval(); enc.precision = seg.precision = "bf16"
x_np, c_np, _ = nb.synthetic_blocks(0)
x, cent = torch.from_numpy(x_np).to(dev), torch.from_numpy(c_np).to(dev)
for _ in range(2): nb.forward_pass(enc, seg, x, cent); torch.cuda.synchronize()

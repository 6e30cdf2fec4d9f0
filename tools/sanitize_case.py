"""Small forward + backward through every kernel family (for compute-sanitizer runs)."""
import importlib, os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
import bench_nn as nb
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
pc = torch.from_numpy(rng.random((2, 3000, 11), dtype=np.float32)).to(dev)
idx = amp.fps_indices(pc, 128)
x = torch.from_numpy(rng.random((3 * 2048, 3), dtype=np.float32)).to(dev)
amp.kmeans_constrained_windows(x, [0, len(x)], [3], 2048, 2048)
enc, seg = nb.build_modules(amp, dev, dropout=0.3)
B, N = 10, 300                                   # 3000 rows: tensor-core layer kernels with a ragged last tile
xb = torch.rand(B, N, 9, device=dev); cent = torch.rand(B, 1, 2, device=dev); tg = torch.randint(-1, 5, (B, N), device=dev)
for prec in ("fp32", "bf16"):
    enc.eval(); seg.eval(); enc.precision = seg.precision = prec
    lg, _ = nb.forward_pass(enc, seg, xb, cent)
enc.train(); seg.train()
lg, ft = nb.forward_pass(enc, seg, xb, cent)
loss = torch.nn.functional.cross_entropy(lg, tg, ignore_index=-1) + 0.001 * torch.norm(torch.eye(64, device=dev) - torch.bmm(ft, ft.transpose(2, 1)))
loss.backward()
torch.cuda.synchronize()
print("sanitize case done, loss %.4f" % float(loss))

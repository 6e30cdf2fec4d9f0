"""How much of the eager training step is launch overhead? Capture forward + loss + backward + Adam in ONE CUDA graph
(dropout off: the host-drawn dropout seed cannot be captured) and time eager vs replay.

    python tools/train_graph_probe.py
"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
import bench_nn as nb
dev = torch.device("cuda:0")
enc, seg = nb.build_modules(amp, dev, dropout=float(os.environ.get("DROPOUT", "0.0")))
enc.train(); seg.train()
opt_e = torch.optim.Adam(enc.parameters(), lr=1e-3, fused=True, capturable=True)
opt_s = torch.optim.Adam(seg.parameters(), lr=1e-3, fused=True, capturable=True)
ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=dev), ignore_index=-1)
x_np, c_np, t_np = nb.synthetic_blocks(0)
x, cent, tg = (torch.from_numpy(a).to(dev) for a in (x_np, c_np, t_np))
eye = torch.eye(64, device=dev)
keep = {}

def step():
    opt_e.zero_grad(set_to_none=True); opt_s.zero_grad(set_to_none=True)
    logits, ft = nb.forward_pass(enc, seg, x, cent)
    loss = ce(logits, tg) + 0.001 * torch.norm(eye - torch.bmm(ft, ft.transpose(2, 1)))
    loss.backward()
    opt_e.step(); opt_s.step()
    keep["loss"] = loss.detach()

def timed(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

print("eager  %.3f ms/step  loss %.5f" % (timed(step), float(keep["loss"])))
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side)
g = torch.cuda.CUDAGraph()
opt_e.zero_grad(set_to_none=True); opt_s.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    step()
print("graph  %.3f ms/step  loss %.5f" % (timed(g.replay), float(keep["loss"])))

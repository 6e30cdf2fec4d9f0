"""Device-only time of the eval forward (CUDA graph replay) vs eager, both precisions (run on a B200)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
nb = importlib.import_module("3d-semantic-segmentation-amp-net_b200.nn_bench")
dev = torch.device("cuda:0")
enc, seg = nb.build_modules(amp, dev); enc.eval(); seg.eval()
x_np, c_np, _ = nb.synthetic_blocks(0)
x, cent = torch.from_numpy(x_np).to(dev), torch.from_numpy(c_np).to(dev)
def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for prec in ("fp32", "bf16"):
    enc.precision = seg.precision = prec
    eager = timeit(lambda: nb.forward_pass(enc, seg, x, cent))
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): nb.forward_pass(enc, seg, x, cent)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        lg, _ = nb.forward_pass(enc, seg, x, cent)
    graph = timeit(g.replay)
    print("forward 32x2048 %s: eager %.3f ms, CUDA graph replay %.3f ms" % (prec, eager, graph), flush=True)

"""Does splitting the eval batch over two streams hide the few-SM glue kernels behind the other half's chain kernels?
(512 tiles on 296 tile slots are two rounds either way.)   python tools/fwd_split_probe.py"""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
import bench_nn as nb
dev = torch.device("cuda:0")
enc, seg = nb.build_modules(amp, dev); enc.eval(); seg.eval()
x_np, c_np, _ = nb.synthetic_blocks(0)
x, cent = torch.from_numpy(x_np).to(dev), torch.from_numpy(c_np).to(dev)
flush = nb._Flush(dev)

def timed(fn, n=20):
    for _ in range(5): flush(); fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / n

def capture(parts):
    streams = [torch.cuda.Stream(device=dev) for _ in range(parts)]
    sl = [slice(i * 32 // parts, (i + 1) * 32 // parts) for i in range(parts)]
    out = {}
    def run():
        cur = torch.cuda.current_stream()
        if parts == 1:
            out[0], _ = nb.forward_pass(enc, seg, x, cent); return
        for s in streams: s.wait_stream(cur)
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                out[i], _ = nb.forward_pass(enc, seg, x[sl[i]], cent[sl[i]])
        for s in streams: cur.wait_stream(s)
    side = torch.cuda.Stream(device=dev); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run(); run()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    return g, out

ref = None
for parts in (1, 2, 4):
    g, out = capture(parts)
    ms = timed(g.replay)
    lg = torch.cat([out[i] for i in range(parts)], 0)
    if ref is None: ref = lg.clone()
    print("parts %d: %.4f ms  identical to unsplit: %s" % (parts, ms, bool(torch.equal(lg, ref))))

"""bf16 tensor-core forward against the reference golden vectors + timing (run on a B200)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
import test_nn_gpu as T
from oracle import make_golden_nn
dev = torch.device("cuda:0")
z = np.load(T.GOLDEN)
for name in sorted(make_golden_nn.CASES):
    seed, xs, cent, mask = T._case(name)
    enc, seg, _, _ = T._build(amp, seed, dev)
    enc.eval(); seg.eval()
    for prec in ("fp32", "bf16"):
        enc.precision = seg.precision = prec
        logits, ft, out = T._run(enc, seg, xs, cent, mask, dev)
        torch.cuda.synchronize()
        ref = torch.from_numpy(z[name + "__eval_logits"])
        agree = (logits.argmax(1).cpu() == ref.argmax(1)).float().mean().item()
        print("%-28s %s logits rel %.3e  relnorm %.3e  argmax agree %.4f | ft rel %.3e | out rel %.3e" % (
            name, prec, T._rel(logits, ref), T._relnorm(logits, ref), agree, T._rel(ft, z[name + "__eval_ft"]),
            T._rel(out[:, ::37, :], z[name + "__eval_enc_out_last"])), flush=True)
# timing at the bench shape
nb = importlib.import_module("3d-semantic-segmentation-amp-net_b200.nn_bench")
enc, seg = nb.build_modules(amp, dev); enc.eval(); seg.eval()
x_np, c_np, _ = nb.synthetic_blocks(0)
x, cent = torch.from_numpy(x_np).to(dev), torch.from_numpy(c_np).to(dev)
res = {}
for prec in ("fp32", "bf16"):
    enc.precision = seg.precision = prec
    for _ in range(3): lg, _ = nb.forward_pass(enc, seg, x, cent)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): lg, _ = nb.forward_pass(enc, seg, x, cent)
    b.record(); torch.cuda.synchronize()
    res[prec] = lg
    print("forward 32x2048 %s: %.3f ms/step" % (prec, a.elapsed_time(b) / 20), flush=True)
print("bf16 vs fp32 logits rel %.3e relnorm %.3e argmax agree %.4f" % (T._rel(res["bf16"], res["fp32"]), T._relnorm(res["bf16"], res["fp32"]),
      (res["bf16"].argmax(1) == res["fp32"].argmax(1)).float().mean().item()))
# CUDA-graph replay of the bf16 forward: pure device time of the step
enc.precision = seg.precision = "bf16"
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): lg, _ = nb.forward_pass(enc, seg, x, cent)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    lg, _ = nb.forward_pass(enc, seg, x, cent)
for _ in range(3): g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): g.replay()
b.record(); torch.cuda.synchronize()
print("forward 32x2048 bf16 CUDA graph replay: %.3f ms/step" % (a.elapsed_time(b) / 50))
print("graph logits vs eager rel %.3e" % T._rel(lg, res["bf16"]))

"""Phase timeline (clock64) of the fused chains of one eval forward: AMP_CHAIN32_PROF=1 (fp32 path) / AMP_CHAIN_PROF=1 (bf16).

    AMP_CHAIN32_PROF=1 python tools/chain_prof.py [fp32|bf16]
"""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
import bench_nn as nb
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
enc, seg = nb.build_modules(amp, dev); enc.eval(); seg.eval(); enc.precision = seg.precision = prec
x_np, c_np, _ = nb.synthetic_blocks(0)
x, cent = torch.from_numpy(x_np).to(dev), torch.from_numpy(c_np).to(dev)
for i in range(3):
    sys.stderr.write("--- forward %d\n" % i)
    nb.forward_pass(enc, seg, x, cent); torch.cuda.synchronize()

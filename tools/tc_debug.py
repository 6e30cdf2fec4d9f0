"""First-contact check of the tcgen05 chain kernel on a B200: one linear layer against torch (bf16-rounded operands)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
from importlib import import_module
tc = import_module("3d-semantic-segmentation-amp-net_b200.tensorcore")
dev = torch.device("cuda:0")
torch.manual_seed(0)
def ref(x, w, b, relu):
    y = x.bfloat16().float() @ w.bfloat16().float().t()
    if b is not None: y = y + b
    return torch.relu(y) if relu else y
for (C, R, K, N, relu, pool) in [(1, 128, 16, 16, False, False), (1, 128, 64, 64, False, False), (1, 128, 64, 64, True, False),
                                 (2, 256, 128, 256, True, False), (3, 200, 128, 128, True, False), (2, 256, 128, 256, True, True),
                                 (3, 200, 64, 128, True, True), (32, 2048, 128, 256, True, True), (32, 2048, 64, 128, True, False)]:
    x = torch.randn(C, R, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    out = tc.tc_linear(x, w, b, relu=relu, pool=pool)
    torch.cuda.synchronize()
    e = ref(x, w, b, relu)
    if pool: e = e.max(dim=1).values
    err = (out - e).abs().max().item()
    print("C=%d R=%d K=%d N=%d relu=%d pool=%d  max|err|=%.3e  ref max=%.3f  %s" % (C, R, K, N, relu, pool, err, e.abs().max().item(),
          "OK" if err < 1e-3 else "MISMATCH"), flush=True)
    if err >= 1e-3 and not pool:
        d = (out - e).abs()
        bad = (d > 1e-3)
        print("   bad fraction %.4f; bad rows %s; bad cols %s" % (bad.float().mean().item(), bad.any(2).nonzero()[:8].flatten().tolist(), bad.any(1).any(0).nonzero()[:16].flatten().tolist()))
        print("   out[0,0,:8]", out[0, 0, :8].tolist()); print("   ref[0,0,:8]", e[0, 0, :8].tolist())

"""First-contact check of the MN-major tcgen05 weight-gradient kernel against torch (run on a B200)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (C, R, N, K, bias) in [(1, 2048, 64, 64, False), (1, 2048, 64, 64, True), (4, 2048, 128, 64, True), (32, 2048, 256, 128, False),
                           (32, 2048, 128, 256, False), (3, 1000, 64, 128, True), (8, 300, 128, 128, False), (2, 100, 64, 64, True), (32, 2048, 64, 9, False), (4, 2048, 64, 3, True), (1, 32, 4096, 128, True), (1, 288, 768, 256, True)]:
    dy = torch.randn(C, R, N, device=dev); a = torch.randn(C, R, K, device=dev)
    out = amp.linear_wgrad(dy, a, bias=bias)
    torch.cuda.synchronize()
    dw = out[0] if bias else out
    ref = (dy.double().reshape(-1, N).t() @ a.double().reshape(-1, K))
    err = ((dw.double() - ref).abs().max() / ref.abs().max()).item()
    msg = "C=%d R=%d N=%d K=%d bias=%d  dW rel err %.3e" % (C, R, N, K, bias, err)
    if bias:
        rb = dy.double().reshape(-1, N).sum(0)
        msg += "  db rel err %.3e" % ((out[1].double() - rb).abs().max() / rb.abs().max()).item()
    print(msg, "OK" if err < 1e-4 else "MISMATCH", flush=True)
    if err >= 1e-4:
        d = (dw.double() - ref).abs() / ref.abs().max()
        bad = d > 1e-4
        print("   bad frac %.3f rows %s cols %s" % (bad.float().mean().item(), bad.any(1).nonzero()[:10].flatten().tolist(), bad.any(0).nonzero()[:10].flatten().tolist()))
        print("   got", dw[0, :6].tolist()); print("   ref", ref[0, :6].tolist())

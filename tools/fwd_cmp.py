import sys, torch
a, b = torch.load(sys.argv[1]), torch.load(sys.argv[2])
for k in a:
    x, y = a[k].double(), b[k].double()
    den = y.abs().max().clamp_min(1e-30)
    e = float((x - y).abs().max() / den)
    if e > 1e-5 or k in ("logits", "ft", "out"): print("%-50s rel diff %.3e" % (k, e))

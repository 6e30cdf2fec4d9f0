import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
from oracle import nn_params
import test_nn_gpu as T
cuda = torch.device("cuda:0")
B, N, seed = 4, 512, 81
xs, cent = nn_params.synthetic_blocks(B, N, 1, seed)
def run(zero_copy):
    enc, seg, _, _ = T._build(amp, seed, cuda)
    enc.train(); seg.train()
    params = list(enc.parameters()) + list(seg.parameters())
    red = amp.GradAllReduce(params, world=1, zero_copy=True) if zero_copy else None
    out, ft = enc(xs[0].to(cuda))
    out.retain_grad(); ft.retain_grad()
    lo = out[:, :, -64:].contiguous(); gl = torch.transpose(out[:, 0, :-64].reshape(-1, 1, 256), 0, 1).contiguous()
    logits, _ = seg(gl, lo, cent.to(cuda), [N], None)
    (logits.square().mean() + 0.01 * ft.square().mean()).backward()
    g = [p.grad.clone() for p in params]
    return out.detach().clone(), logits.detach().clone(), out.grad.clone(), ft.grad.clone(), g, len(list(enc.parameters()))
o0, l0, do0, df0, g0, ne = run(False)
o1, l1, do1, df1, g1, _ = run(True)
print("out", T._rel(o0, o1), "logits", T._rel(l0, l1), "d_out", T._rel(do0, do1), "d_ft", T._rel(df0, df1))
print("enc grads max rel", max(T._rel(a, b) for a, b in zip(g0[:ne], g1[:ne])), "seg grads max rel", max(T._rel(a, b) for a, b in zip(g0[ne:], g1[ne:])))
print("enc grads: norm ratio zc/normal", [round(float(b.norm() / a.norm()), 3) for a, b in zip(g0[:8], g1[:8])])

import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
import test_nn_gpu as T
from oracle import nn_params
cuda = torch.device("cuda:0")
B, N, W, seed = 8, 192, 2, 31
enc, seg, sd_e, sd_s = T._build(amp, seed, cuda)
xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
enc.train(); seg.train()
logits, ft, out = T._run(enc, seg, xs, cent, None, cuda)
tg = torch.randint(-1, 5, (B, N * W), generator=torch.Generator().manual_seed(3))
ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=cuda), ignore_index=-1)
loss = ce(logits, tg.to(cuda)) + 0.001 * torch.norm(torch.eye(64, device=cuda) - torch.bmm(ft, ft.transpose(2, 1)))
loss.backward()
d = {"logits": logits.detach().cpu(), "ft": ft.detach().cpu(), "out": out.detach().cpu()}
for m, tag in ((enc, "enc"), (seg, "seg")):
    for k, p in m.named_parameters(): d[tag + ".grad." + k] = p.grad.cpu()
    for k, b in m.named_buffers(): d[tag + ".buf." + k] = b.cpu()
torch.save(d, sys.argv[1])

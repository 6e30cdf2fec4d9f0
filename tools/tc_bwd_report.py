"""Per-tensor report behind tests/test_nn_backward_tc_gpu.py: ours vs float64 truth, fp32 envelope, tensor-core vs CUDA-core backward."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
t = importlib.import_module("test_nn_backward_tc_gpu")
from oracle import nn_params  # noqa: E402

dev = torch.device("cuda:0")
for (B, N, W, seed) in [(32, 2048, 1, 91), (4, 2048, 2, 92), (8, 512, 1, 21)]:
    xs, cent = nn_params.conditioned_blocks(B, N, W, seed)
    tg = torch.randint(-1, 5, (B, N * W), generator=torch.Generator().manual_seed(seed))
    enc, seg, sd_e, sd_s = t._build(amp, seed, dev)
    logits, loss, g_tc, fwd, ran = t._step(amp, enc, seg, xs, cent, tg, dev)
    enc, seg, _, _ = t._build(amp, seed, dev)
    _, loss_b, g_cc, _, ran_b = t._step(amp, enc, seg, xs, cent, tg, dev, disable_for_backward=["tc_layer", "tc_wgrad"])
    enc, seg, _, _ = t._build(amp, seed, dev)
    amp._lib.set_disabled(["tc_layer", "tc_wgrad"])
    _, loss_c, g_all_cc, _, _ = t._step(amp, enc, seg, xs, cent, tg, dev)
    amp._lib.set_disabled(None)
    t_logits, t_loss, truth = t._oracle_grads(sd_e, sd_s, xs, cent, tg, torch.float64)
    env = t._fp32_envelope(sd_e, sd_s, xs, cent, tg, truth)
    print("== B=%d N=%d W=%d  fwd=%s ran=%s ran_b=%s loss %.6f %.6f %.6f truth %.6f logits rel %.2e" %
          (B, N, W, fwd, ran, ran_b, float(loss), float(loss_b), float(loss_c), float(t_loss), t._rel(logits, t_logits)))
    print("%-44s %10s %10s %10s %10s" % ("tensor", "tc-truth", "cc-truth", "env", "tc-vs-cc"))
    for k in sorted(truth):
        if float(truth[k].norm()) < 1e-9:
            continue
        print("%-44s %10.2e %10.2e %10.2e %10.2e %10.2e" % (k, t._relnorm(g_tc[k], truth[k]), t._relnorm(g_cc[k], truth[k]), env[k],
                                                 t._relnorm(g_tc[k], g_cc[k]), t._relnorm(g_all_cc[k], truth[k])))

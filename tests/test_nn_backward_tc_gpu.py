"""GPU parity of the TENSOR-CORE backward (tc_layer_kernel input-gradient modes, tc_wgrad_kernel): the kernels that train
configs[2] / configs[4]. Every encoder call here has >= 2048 rows, the threshold below which the CUDA-core kernels serve
the layers (nn_tc_layer.cu / nn_tc_wgrad.cu); amp_path_count proves which family ran.

  * the same backward with the tensor-core kernels switched off (CUDA-core fp32 tiles) on the SAME saved forward state:
    isolates the tcgen05 backward kernels from max-pool / ReLU tie flips of the forward -- agreement 1e-4 (measured 2e-5);
  * every parameter gradient, d(lo_feats), d(gl_feats) against autograd through the CPU oracle (float64 = truth) at
    configs[2] shape (32 x 2048, W = 1) and at 4 x 2048, W = 2. The loss of this network is only piecewise smooth: the
    three max-pools route each (cloud, channel) gradient to ONE row and near-tied winners flip under any forward error
    larger than their gap, moving single gradients by 1e-3 .. 1e-1 of their norm although every kernel is right (measured
    with bf16 hi + lo terms in the training forward, ~4e-6 per layer: 16 % on one tensor at 8 x 512). The training forward
    therefore splits its operands into fp16 terms (2^-23 residual, fp32 class) and the bound is the envelope of the
    float64 oracle run with the forward perturbed at THAT error level (1e-6), next to the fp32 reference's own envelope;
  * precision "fp32_strict" (plain fp32 FMA kernels, no split-bf16): within the fp32 reference's own envelope;
  * golden vectors of the UNMODIFIED reference modules at tensor-core shapes (tests/golden/nn_reference_tc.npz, made by
    oracle/make_golden_nn.py::make_tc): reference autograd backward of pointnetAtt.py:80-112,176-209.
"""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_nn, nn_oracle, nn_params

pytestmark = pytest.mark.gpu
GOLDEN_TC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nn_reference_tc.npz")
TC_PATHS = ("tc_layer", "tc_layer_dgrad", "tc_wgrad", "narrow_dgrad")


def _relnorm(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _build(amp, seed, dev):
    enc = amp.BasePointNet(point_dimension=3, return_local_features=True, global_feat_dim=256, device=dev)
    seg = amp.SegmentationWithAttention(256, 8, num_classes=5, local_dim=64, dropout=0.0, device=dev)
    sd_e = nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed)
    sd_s = nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1)
    enc.load_state_dict(sd_e, strict=True)
    seg.load_state_dict(sd_s, strict=True)
    return enc.to(dev).train(), seg.to(dev).train(), sd_e, sd_s


def _step(amp, enc, seg, xs, cent, tg, dev, disable_for_backward=None):
    """train_pointnet-attention.py:396-467 on the drop-in modules; returns logits, loss, d_lo, d_gl and the counters."""
    for m in (enc, seg):
        m.zero_grad(set_to_none=True)
    before = {p: amp._lib.path_count(p) for p in TC_PATHS}
    lo = torch.FloatTensor().to(dev); gl = torch.FloatTensor().to(dev); npc = []
    for xw in xs:
        out, ft = enc(xw.to(dev))
        local_feat = out[:, :, -64:]
        global_feat = out[:, 0, :-64].view(-1, 1, 256)
        npc.append(local_feat.shape[1])
        lo = torch.cat((lo, local_feat), dim=1)
        gl = torch.cat((gl, global_feat), dim=1)
    gl = torch.transpose(gl, 0, 1)
    lo.retain_grad(); gl.retain_grad()
    logits, _ = seg(gl, lo, cent.to(dev), npc, None)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=dev), ignore_index=-1)
    loss = ce(logits, tg.to(dev)) + 0.001 * torch.norm(torch.eye(64, device=dev) - torch.bmm(ft, ft.transpose(2, 1)))
    fwd = {p: amp._lib.path_count(p) - before[p] for p in TC_PATHS}
    if disable_for_backward is not None:
        amp._lib.set_disabled(disable_for_backward)
    try:
        loss.backward()
        torch.cuda.synchronize()
    finally:
        if disable_for_backward is not None:
            amp._lib.set_disabled(None)
    ran = {p: amp._lib.path_count(p) - before[p] for p in TC_PATHS}
    grads = {"enc." + k: p.grad.detach().clone() for k, p in enc.named_parameters()}
    grads.update({"seg." + k: p.grad.detach().clone() for k, p in seg.named_parameters()})
    grads["d_lo"] = lo.grad.detach().clone()
    grads["d_gl"] = gl.grad.detach().clone()
    return logits.detach(), loss.detach(), grads, fwd, ran


def _oracle_grads(sd_e, sd_s, xs, cent, tg, dtype, forward_noise=0.0, noise_seed=1):
    """forward_noise > 0: every wide (>= 2048-row) layer output gets i.i.d. noise of that relative size (of the layer's mean
    magnitude): the gradient an implementation whose forward arithmetic is good to `forward_noise` would be entitled to."""
    orig = nn_oracle._conv
    if forward_noise > 0.0:
        gen = torch.Generator().manual_seed(noise_seed)

        def noisy(x, sd, pre):
            y = orig(x, sd, pre)
            if y.shape[0] >= 2048:
                y = y + (y.detach().abs().mean() * forward_noise) * torch.randn(y.shape, generator=gen, dtype=torch.float64).to(y.dtype)
            return y
        nn_oracle._conv = noisy
    try:
        return _oracle_grads_plain(sd_e, sd_s, xs, cent, tg, dtype)
    finally:
        nn_oracle._conv = orig


def _oracle_grads_plain(sd_e, sd_s, xs, cent, tg, dtype):
    se = {k: (v.detach().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd_e.items()}
    ss = {k: (v.detach().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sd_s.items()}
    for sd in (se, ss):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    taps = {}
    logits, ft = nn_oracle.forward_windows(se, ss, [x.to(dtype) for x in xs], cent.to(dtype), None, training=True, taps=taps)
    loss, _, _ = nn_oracle.train_step_loss(logits, tg, ft)
    loss.backward()
    grads = {"enc." + k: v.grad for k, v in se.items() if v.is_floating_point() and v.requires_grad}
    grads.update({"seg." + k: v.grad for k, v in ss.items() if v.is_floating_point() and v.requires_grad})
    grads["d_lo"] = taps["lo_feats"].grad
    grads["d_gl"] = taps["gl_feats"].grad
    return logits.detach(), loss.detach(), grads


SPLIT_FP16_FORWARD_ERROR = 1e-6      # relative error bound of one fp16-split layer (2^-23 residual, dropped lo*lo, fp32 accumulate)


def _forward_noise_envelope(sd_e, sd_s, xs, cent, tg, truth, eps=SPLIT_FP16_FORWARD_ERROR):
    """Per tensor: how far the EXACT (float64) gradient moves when the forward is perturbed at the training forward's error
    level (max over three noise draws): the tie flips of the max-pools (see module docstring)."""
    env = {}
    for seed in (1, 2, 3):
        _, _, g = _oracle_grads(sd_e, sd_s, xs, cent, tg, torch.float64, forward_noise=eps, noise_seed=seed)
        for k, t in truth.items():
            if float(t.norm()) >= 1e-9:
                env[k] = max(env.get(k, 0.0), _relnorm(g[k], t))
    return env


def _fp32_envelope(sd_e, sd_s, xs, cent, tg, truth):
    """Distance of the fp32 reference arithmetic to the exact gradient, per tensor: the maximum over the clean fp32 oracle
    run and two runs whose inputs are perturbed by 1e-7 relative. At these sizes a handful of near-tied max-pool winners and
    ReLU thresholds flip under ANY change of rounding (another BLAS, another summation order), which moves a gradient by
    ~1e-3 of its norm: the honest bound is that noise envelope, not the arithmetic error of one GEMM (that is the
    tensor-core vs CUDA-core test below, 1e-4)."""
    env = {}
    g = torch.Generator().manual_seed(5)
    for i in range(3):
        xi = xs if i == 0 else [x * (1 + 1e-7 * torch.randn(x.shape, generator=g)) for x in xs]
        _, _, ref = _oracle_grads(sd_e, sd_s, xi, cent, tg, torch.float32)
        for k, t in truth.items():
            if float(t.norm()) >= 1e-9:
                env[k] = max(env.get(k, 0.0), _relnorm(ref[k], t))
    return env


@pytest.mark.parametrize("B,N,W,seed", [(32, 2048, 1, 91), (4, 2048, 2, 92)], ids=["configs2_32x2048_w1", "4x2048_w2"])
def test_tensor_core_backward_every_gradient_vs_oracle(amp, cuda, B, N, W, seed):
    enc, seg, sd_e, sd_s = _build(amp, seed, cuda)
    xs, cent = nn_params.conditioned_blocks(B, N, W, seed)
    tg = torch.randint(-1, 5, (B, N * W), generator=torch.Generator().manual_seed(seed))
    logits, loss, ours, fwd, ran = _step(amp, enc, seg, xs, cent, tg, cuda)
    # the tensor-core kernels served this step (forward layers, input-gradient layers, weight gradients)
    assert fwd["tc_layer"] >= 9 * W and ran["tc_layer_dgrad"] >= 8 * W and ran["tc_wgrad"] >= 10 * W, (fwd, ran)
    assert ran["narrow_dgrad"] == 1, ran          # the 5-class logits layer's input gradient: the narrow kernel's masked form
    t_logits, t_loss, truth = _oracle_grads(sd_e, sd_s, xs, cent, tg, torch.float64)
    o_logits, o_loss, _ = _oracle_grads(sd_e, sd_s, xs, cent, tg, torch.float32)
    assert _rel(logits, t_logits) < 1e-3
    assert _rel(logits, t_logits) < max(3 * _rel(o_logits, t_logits), 3e-4)
    assert abs(float(loss) - float(t_loss)) < 1e-4 * abs(float(t_loss))
    assert _rel(logits, t_logits) < 3e-4                # the forward itself (batch statistics over 4 .. 32 clouds)
    env32 = _fp32_envelope(sd_e, sd_s, xs, cent, tg, truth)
    envn = _forward_noise_envelope(sd_e, sd_s, xs, cent, tg, truth)
    report, bad = [], []
    for k in sorted(truth):
        t = truth[k]
        if float(t.norm()) < 1e-9:                      # mathematically zero (bias in front of a BatchNorm)
            assert float(ours[k].norm()) < 1e-4, k
            continue
        e_ours, env = _relnorm(ours[k], t), max(env32[k], envn[k])
        report.append((k, e_ours, env32[k], envn[k]))
        if not (e_ours < 3 * env + 1e-4):
            bad.append((k, e_ours, env32[k], envn[k]))
    assert not bad, "tensor-core gradients outside the tie-flip envelope: %s\nall: %s" % (bad, report)
    assert len(report) >= 60
    # the typical tensor is far inside: median error below 1e-2 even where single tensors carry a flipped tie
    errs = sorted(r[1] for r in report)
    assert errs[len(errs) // 2] < 1e-2, errs


@pytest.mark.parametrize("B,N,W,seed", [(32, 2048, 1, 91), (8, 512, 1, 21)], ids=["configs2_32x2048_w1", "8x512_w1"])
def test_strict_fp32_training_step_within_reference_fp32_envelope(amp, cuda, B, N, W, seed):
    """precision = "fp32_strict": plain fp32 FMA kernels for the forward (where the discontinuities are; the backward, smooth
    in its inputs, keeps the tensor-core kernels); gradients as close to the exact ones as the reference's own fp32
    arithmetic (8 x 512 is the case where the split-bf16 forward flips a tie worth 16 %)."""
    enc, seg, sd_e, sd_s = _build(amp, seed, cuda)
    enc.precision = seg.precision = "fp32_strict"
    xs, cent = nn_params.conditioned_blocks(B, N, W, seed)
    tg = torch.randint(-1, 5, (B, N * W), generator=torch.Generator().manual_seed(seed))
    logits, loss, ours, fwd, ran = _step(amp, enc, seg, xs, cent, tg, cuda)
    assert fwd["tc_layer"] == 0 and ran["tc_wgrad"] > 0, (fwd, ran)
    t_logits, t_loss, truth = _oracle_grads(sd_e, sd_s, xs, cent, tg, torch.float64)
    assert _rel(logits, t_logits) < 2e-5
    env32 = _fp32_envelope(sd_e, sd_s, xs, cent, tg, truth)
    bad = []
    for k in sorted(truth):
        if float(truth[k].norm()) < 1e-9:
            continue
        e = _relnorm(ours[k], truth[k])
        if not (e < 3 * env32[k] + 1e-4 and e < 2e-2):
            bad.append((k, e, env32[k]))
    assert not bad, bad


@pytest.mark.parametrize("B,N,W,seed", [(32, 2048, 1, 93), (4, 2048, 2, 94)], ids=["configs2_32x2048_w1", "4x2048_w2"])
def test_tensor_core_backward_equals_cuda_core_backward(amp, cuda, B, N, W, seed):
    """Same forward (same saved activations, arg-max rows, BatchNorm statistics), backward once through tc_layer_kernel /
    tc_wgrad_kernel and once through the fp32 CUDA-core tiles: the two gradient sets agree to 1e-4."""
    xs, cent = nn_params.conditioned_blocks(B, N, W, seed)
    tg = torch.randint(-1, 5, (B, N * W), generator=torch.Generator().manual_seed(seed))
    enc, seg, _, _ = _build(amp, seed, cuda)
    _, loss_a, g_tc, _, ran_a = _step(amp, enc, seg, xs, cent, tg, cuda)
    enc, seg, _, _ = _build(amp, seed, cuda)
    n_pw, n_wg = amp._lib.path_count("pw_linear"), amp._lib.path_count("wgrad_partial")
    _, loss_b, g_cc, _, ran_b = _step(amp, enc, seg, xs, cent, tg, cuda, disable_for_backward=["tc_layer", "tc_wgrad"])
    assert ran_a["tc_layer_dgrad"] > 0 and ran_a["tc_wgrad"] > 0
    assert ran_b["tc_layer_dgrad"] == 0 and ran_b["tc_wgrad"] == 0, ran_b
    assert amp._lib.path_count("pw_linear") > n_pw and amp._lib.path_count("wgrad_partial") > n_wg
    assert abs(float(loss_a) - float(loss_b)) < 1e-6 * abs(float(loss_b))     # same forward (torch's loss reduction is not bitwise repeatable)
    # (biases in front of a BatchNorm have a mathematically zero gradient: both sides hold rounding noise only)
    scale = max(float(g.norm()) for g in g_cc.values())
    worst = max((_relnorm(g_tc[k], g_cc[k]), k) for k in g_tc if float(g_cc[k].norm()) > 1e-6 * scale)
    assert worst[0] < 1e-4, worst
    assert sum(1 for k in g_tc if float(g_cc[k].norm()) > 1e-6 * scale) >= 60


@pytest.mark.parametrize("name", sorted(make_golden_nn.CASES_TC))
def test_tensor_core_training_step_vs_reference_golden(amp, cuda, name):
    """The unmodified reference modules' own fp32 training step at >= 2048 rows per encoder call."""
    z = np.load(GOLDEN_TC)
    B, N, W, seed = make_golden_nn.CASES_TC[name]
    enc, seg, sd_e, sd_s = _build(amp, seed, cuda)
    xs, cent = nn_params.conditioned_blocks(B, N, W, seed)
    tg = torch.from_numpy(z[name + "__targets"].astype(np.int64))
    logits, loss, ours, fwd, ran = _step(amp, enc, seg, xs, cent, tg, cuda)
    assert ran["tc_layer_dgrad"] > 0 and ran["tc_wgrad"] > 0, ran
    assert _rel(logits[:, :, ::16], z[name + "__train_logits"]) < 5e-4     # (the reference's own fp32 is ~1e-4 from float64 here)
    assert abs(float(loss) - float(z[name + "__train_loss"])) < 1e-4 * abs(float(z[name + "__train_loss"]))
    _, _, truth = _oracle_grads(sd_e, sd_s, xs, cent, tg, torch.float64)
    env32 = _fp32_envelope(sd_e, sd_s, xs, cent, tg, truth)
    envn = _forward_noise_envelope(sd_e, sd_s, xs, cent, tg, truth)
    env = {k: max(env32[k], envn[k]) for k in env32}
    checked = 0
    for key in z.files:
        if "__grad_" not in key or not key.startswith(name):
            continue
        tag, k = key[len(name) + 7:].split("_", 1)
        full = tag + "." + k
        ref = z[key]
        if float(np.linalg.norm(ref)) < 1e-9:
            continue
        t = make_golden_nn.subsample_tc(truth[full].numpy())
        got = make_golden_nn.subsample_tc(ours[full].cpu().numpy())
        if full not in env:                             # mathematically zero gradient (bias in front of a BatchNorm)
            continue
        e_ours, e_ref = _relnorm(got, t), _relnorm(ref, t)
        # subsampled rows: compare against the envelope of the full tensor with head room for the row subset
        assert e_ours < 6 * max(e_ref, env[full]) + 1e-4, (key, e_ours, e_ref, env[full])
        checked += 1
    assert checked >= 60
    assert _relnorm(ours["d_lo"][:, ::101, :], z[name + "__dlo"]) < 6 * env["d_lo"] + 1e-4
    assert _relnorm(ours["d_gl"], z[name + "__dgl"]) < 6 * env["d_gl"] + 1e-4


def test_weight_gradient_from_the_input_gradient_kernels_split_operand_is_bit_identical(amp, cuda):
    """bwd_step runs the input gradient first; its tensor-core kernel leaves dy' (BatchNorm-backward applied, split into bf16
    hi + lo) in a scratch buffer and tc_wgrad_kernel copies it instead of recomputing it from dz and y. Same values, same
    summation order: every gradient must be bit-identical to the path that recomputes (AMP_DISABLE=wgrad_presplit)."""
    B, N, W, seed = 32, 2048, 1, 97
    xs, cent = nn_params.conditioned_blocks(B, N, W, seed)
    tg = torch.randint(-1, 5, (B, N * W), generator=torch.Generator().manual_seed(seed))
    enc, seg, _, _ = _build(amp, seed, cuda)
    n0 = amp._lib.path_count("tc_wgrad_presplit")
    _, _, g_a, _, _ = _step(amp, enc, seg, xs, cent, tg, cuda)
    used = amp._lib.path_count("tc_wgrad_presplit") - n0
    assert used >= 8, used                                   # the encoder's wide layers
    enc, seg, _, _ = _build(amp, seed, cuda)
    n1 = amp._lib.path_count("tc_wgrad_presplit")
    _, _, g_b, _, _ = _step(amp, enc, seg, xs, cent, tg, cuda, disable_for_backward=["wgrad_presplit"])
    assert amp._lib.path_count("tc_wgrad_presplit") == n1
    for k in g_a:
        assert torch.equal(g_a[k], g_b[k]), k

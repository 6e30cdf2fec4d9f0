"""GPU parity of the PointNet-attention forward / backward (C ABI amp_encoder_* / amp_seg_*) against
 (a) the committed golden vectors produced by the UNMODIFIED reference modules (tests/golden/nn_reference.npz,
     made by oracle/make_golden_nn.py) and (b) the CPU oracle (oracle/nn_oracle.py) on seeded inputs.
Tolerances (north star): logits within 1e-3 relative in fp32, >= 99.9 % argmax-label agreement."""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_nn, nn_oracle, nn_params

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nn_reference.npz")
TOL_LOGITS = 1e-3


def _rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _relnorm(a, b):
    a = torch.as_tensor(a).detach().double().cpu(); b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _build(amp, seed, dev, dropout=0.0):
    enc = amp.BasePointNet(point_dimension=3, return_local_features=True, global_feat_dim=256, device=dev)
    seg = amp.SegmentationWithAttention(256, 8, num_classes=5, local_dim=64, dropout=dropout, device=dev)
    sd_e = nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed)
    sd_s = nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1)
    enc.load_state_dict(sd_e, strict=True)
    seg.load_state_dict(sd_s, strict=True)
    return enc.to(dev), seg.to(dev), sd_e, sd_s


def _run(enc, seg, xs, cent, mask, dev):
    """The window loop of train_pointnet-attention.py:396-435, driving the drop-in modules."""
    lo = torch.FloatTensor().to(dev); gl = torch.FloatTensor().to(dev); npc = []
    for xw in xs:
        out, ft = enc(xw.to(dev))
        local_feat = out[:, :, -64:]
        global_feat = out[:, 0, :-64].view(-1, 1, 256)
        npc.append(local_feat.shape[1])
        lo = torch.cat((lo, local_feat), dim=1)
        gl = torch.cat((gl, global_feat), dim=1)
    gl = torch.transpose(gl, 0, 1)
    logits, zero = seg(gl, lo, cent.to(dev), npc, None if mask is None else mask.to(dev))
    assert zero == 0
    return logits, ft, out


def _case(name):
    B, N, W, seed, masked = make_golden_nn.CASES[name]
    xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
    mask = None
    if masked:
        mask = torch.zeros(B, W, dtype=torch.bool); mask[0, W - 1] = True
    return seed, xs, cent, mask


@pytest.mark.parametrize("name", sorted(make_golden_nn.CASES))
def test_eval_forward_matches_reference_golden(amp, cuda, name):
    z = np.load(GOLDEN)
    seed, xs, cent, mask = _case(name)
    enc, seg, _, _ = _build(amp, seed, cuda)
    enc.eval(); seg.eval()
    n0 = amp._lib.launch_count()
    logits, ft, out = _run(enc, seg, xs, cent, mask, cuda)
    assert amp._lib.launch_count() > n0
    assert not logits.requires_grad
    ref = torch.from_numpy(z[name + "__eval_logits"])
    assert _rel(logits, ref) < TOL_LOGITS
    assert (logits.argmax(1).cpu() == ref.argmax(1)).float().mean().item() >= 0.999
    assert _rel(ft, z[name + "__eval_ft"]) < TOL_LOGITS
    assert _rel(out[:, ::37, :], z[name + "__eval_enc_out_last"]) < TOL_LOGITS


@pytest.mark.parametrize("name", sorted(make_golden_nn.CASES))
def test_train_step_matches_reference_golden(amp, cuda, name):
    z = np.load(GOLDEN)
    seed, xs, cent, mask = _case(name)
    enc, seg, _, _ = _build(amp, seed, cuda)
    enc.train(); seg.train()
    logits, ft, _ = _run(enc, seg, xs, cent, mask, cuda)
    tg = torch.from_numpy(z[name + "__targets"]).to(cuda)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=cuda), reduction="mean", ignore_index=-1)
    loss = ce(logits, tg) + 0.001 * torch.norm(torch.eye(64, device=cuda) - torch.bmm(ft, ft.transpose(2, 1)))
    loss.backward()
    assert _rel(logits, z[name + "__train_logits"]) < TOL_LOGITS
    assert abs(float(loss.detach()) - float(z[name + "__train_loss"])) < 1e-4 * abs(float(z[name + "__train_loss"]))
    # Exact (float64) gradients from the oracle: the reference's own fp32 gradients (the golden vectors) are up to a
    # few 1e-2 away from them on these tiny batches (BatchNorm over 3-4 near-identical clouds, near-tied max-pool
    # winners), so the bound is "as close to the exact gradient as the unmodified reference is".
    sd_e64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed).items()}
    sd_s64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1).items()}
    for sd in (sd_e64, sd_s64):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    t_logits, t_ft = nn_oracle.forward_windows(sd_e64, sd_s64, [x.double() for x in xs], cent.double(), mask, training=True)
    nn_oracle.train_step_loss(t_logits, tg.cpu(), t_ft)[0].backward()
    checked = 0
    for key in z.files:
        if key.startswith(name + "__grad_"):
            tag, k = key[len(name) + 7:].split("_", 1)
            g = dict((enc if tag == "enc" else seg).named_parameters())[k].grad
            assert g is not None, key
            truth = make_golden_nn.subsample((sd_e64 if tag == "enc" else sd_s64)[k].grad.numpy())
            ours = _relnorm(make_golden_nn.subsample(g.cpu().numpy()), truth)
            ref32 = _relnorm(z[key], truth)
            # (a 1e-4 wobble of the feature transform flips ~1 % of the near-tied global max-pool winners, which moves
            # a gradient by ~1e-2 of its norm; test_all_gradients_match_oracle is the tight, well-conditioned check)
            assert ours < max(10 * ref32, 3e-2), (key, ours, ref32)
            assert _relnorm(make_golden_nn.subsample(g.cpu().numpy()), z[key]) < 6e-2, key
            checked += 1
    assert checked >= 10
    assert _rel(enc.bn_6.running_mean, z[name + "__train_rm_bn_6"]) < 1e-4
    assert _rel(enc.bn_1.running_var, z[name + "__train_rv_bn_1"]) < 1e-4
    assert _rel(seg.bn_2.running_var, z[name + "__train_rv_seg_bn_2"]) < 1e-4
    assert int(enc.bn_1.num_batches_tracked) == 7 + len(xs)
    assert int(seg.bn_3.num_batches_tracked) == 8


def _gradient_check(amp, cuda, seed):
    """Every parameter gradient of both modules against autograd through the CPU oracle (dropout off).
    Returns (tight_ok, worst) where worst = max over parameters of the relative error to the exact (float64) gradient."""
    B, N, W = 8, 192, 2
    enc, seg, sd_e, sd_s = _build(amp, seed, cuda)
    xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
    # clouds that differ from each other (per-cloud channel scales / offsets): with the i.i.d. uniform blocks of
    # synthetic_blocks the T-Net BatchNorms over the B pooled rows divide by a near-zero spread and every fp32
    # implementation, the reference included, is only good to ~1e-2; here the problem is well conditioned and the
    # bound can be tight
    g = torch.Generator().manual_seed(9)
    xs = [x * (0.15 + 0.85 * torch.rand(B, 1, 9, generator=g)) + 0.3 * torch.randn(B, 1, 9, generator=g) for x in xs]
    cent = torch.stack([x[:, :, :2].mean(1) for x in xs], 1)
    enc.train(); seg.train()
    logits, ft, _ = _run(enc, seg, xs, cent, None, cuda)
    tg = torch.randint(-1, 5, (B, N * W), generator=torch.Generator().manual_seed(3))
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=cuda), ignore_index=-1)
    loss = ce(logits, tg.to(cuda)) + 0.001 * torch.norm(torch.eye(64, device=cuda) - torch.bmm(ft, ft.transpose(2, 1)))
    loss.backward()
    for sd in (sd_e, sd_s):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    o_logits, o_ft = nn_oracle.forward_windows(sd_e, sd_s, xs, cent, None, training=True, stats_enc={}, stats_seg={})
    o_loss, _, _ = nn_oracle.train_step_loss(o_logits, tg, o_ft)
    o_loss.backward()
    assert _rel(logits, o_logits) < TOL_LOGITS
    # float64 run of the oracle = the exact answer: fp32 train-mode gradients of this net are noisy in the reference
    # itself (BatchNorm over a handful of near-identical clouds, near-tied max-pool winners), so the statement that
    # holds is "the CUDA path is as close to the exact gradient as the fp32 reference arithmetic is"
    sd_e64 = {k: (v.detach().double() if v.is_floating_point() else v.clone()) for k, v in sd_e.items()}
    sd_s64 = {k: (v.detach().double() if v.is_floating_point() else v.clone()) for k, v in sd_s.items()}
    # the fp32 oracle run above already advanced the BatchNorm buffers of sd_e / sd_s: batch statistics do not use them
    for sd in (sd_e64, sd_s64):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    t_logits, t_ft = nn_oracle.forward_windows(sd_e64, sd_s64, [x.double() for x in xs], cent.double(), None, training=True)
    t_loss, _, _ = nn_oracle.train_step_loss(t_logits, tg, t_ft)
    t_loss.backward()
    assert _rel(logits, t_logits) < max(3 * _rel(o_logits, t_logits), 2e-4)
    tight, worst = True, 0.0
    for mod, sd, sd64 in ((enc, sd_e, sd_e64), (seg, sd_s, sd_s64)):
        for k, p in mod.named_parameters():
            assert p.grad is not None, k
            tg64 = sd64[k].grad
            if float(tg64.norm()) < 1e-6:     # biases in front of a BatchNorm: mathematically zero gradient
                assert float(p.grad.norm()) < 1e-4, k
                continue
            ours, ref32 = _relnorm(p.grad, tg64), _relnorm(sd[k].grad, tg64)
            tight = tight and ours < 2.5 * ref32 + 1e-4 and ours < 1e-3
            worst = max(worst, ours)
    for mod, sd in ((enc, sd_e), (seg, sd_s)):
        for k, b in mod.named_buffers():
            if "running" in k:
                assert _rel(b, sd[k]) < 1e-4, k
    return tight, worst


def test_all_gradients_match_oracle(amp, cuda):
    """Tight bound (within 2.5x of the fp32 reference's own distance to the exact gradient, and < 1e-3) on at least two of
    three seeds, loose bound (1e-2) on all. Why not all three tight: a single activation that sits within ~1e-7 of a ReLU
    threshold flips its mask under ANY change of summation order (the reference on another BLAS does the same), and one
    flipped element moves every upstream gradient by ~1e-3 of its norm (measured: exactly 1 of 64 bn_3.bias channels off
    by 2e-3, all others at 1e-7). The arithmetic itself is good to ~1e-5 (split-bf16 tensor-core GEMMs) or 1e-7 (fp32)."""
    results = [_gradient_check(amp, cuda, seed) for seed in (31, 32, 33)]
    assert all(w < 1e-2 for _, w in results), results
    assert sum(1 for t, _ in results if t) >= 2, results


def test_eval_forward_full_size_matches_oracle(amp, cuda):
    """configs[0] shape: batch 32 x 2048 points, one block per window."""
    B, N, W, seed = 32, 2048, 1, 41
    enc, seg, sd_e, sd_s = _build(amp, seed, cuda)
    xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
    enc.eval(); seg.eval()
    logits, ft, out = _run(enc, seg, xs, cent, None, cuda)
    with torch.no_grad():
        o_logits, o_ft = nn_oracle.forward_windows(sd_e, sd_s, xs, cent, None, training=False)
    assert _rel(logits, o_logits) < TOL_LOGITS
    assert _rel(ft, o_ft) < TOL_LOGITS
    assert (logits.argmax(1).cpu() == o_logits.argmax(1)).float().mean().item() >= 0.999
    # size-independent properties: the global half of the encoder output is constant over the points of a block
    assert bool((out[:, :, :256] == out[:, :1, :256]).all())


def test_eval_test_script_shape_variable_blocks(amp, cuda):
    """test_pointnet_att_segmen.py:160-177: batch 1, blocks of different sizes (>= n_points), no mask."""
    seed = 51
    enc, seg, sd_e, sd_s = _build(amp, seed, cuda)
    enc.eval(); seg.eval()
    rng = np.random.default_rng(seed)
    sizes = [300, 257, 411]
    xs = []
    for n in sizes:
        x = rng.random((1, n, 9), dtype=np.float32); x[:, :, :2] = x[:, :, :2] * 2 - 1; x[:, :, 2] *= 0.3
        xs.append(torch.from_numpy(x))
    cent = torch.stack([x[:, :, :2].mean(1) for x in xs], 1)
    logits, ft, _ = _run(enc, seg, xs, cent, None, cuda)
    with torch.no_grad():
        o_logits, o_ft = nn_oracle.forward_windows(sd_e, sd_s, xs, cent, None, training=False)
    assert tuple(logits.shape) == (1, 5, sum(sizes))
    assert _rel(logits, o_logits) < TOL_LOGITS


def test_dropout_training_is_seeded_and_unbiased(amp, cuda):
    B, N, W, seed = 4, 256, 2, 61
    enc, seg, _, _ = _build(amp, seed, cuda, dropout=0.3)
    xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
    enc.train(); seg.train()
    torch.manual_seed(5)
    l1, _, _ = _run(enc, seg, xs, cent, None, cuda)
    torch.manual_seed(5)
    l2, _, _ = _run(enc, seg, xs, cent, None, cuda)
    torch.manual_seed(6)
    l3, _, _ = _run(enc, seg, xs, cent, None, cuda)
    assert torch.isfinite(l1).all()
    assert _rel(l1, l2) < 1e-4          # same seed -> same masks (BN running buffers do not enter train-mode outputs)
    assert _rel(l1, l3) > 1e-3          # different seed -> different masks
    l1.sum().backward()
    for m in (enc, seg):
        for k, p in m.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), k


def test_no_cpu_fallback(amp):
    enc = amp.BasePointNet(point_dimension=3, return_local_features=True)
    with pytest.raises(RuntimeError):
        enc(torch.zeros(2, 16, 9))


# ---- bf16 tensor-core eval path (AMP_PREC_BF16): stated separately from the fp32 parity bound --------------------
TOL_BF16_LOGITS = 2e-2      # max-norm relative; measured 2.7e-3 .. 6.6e-3 on the golden cases
TOL_BF16_FT = 1e-3


@pytest.mark.parametrize("name", sorted(make_golden_nn.CASES))
def test_bf16_tensor_core_forward_vs_reference_golden(amp, cuda, name):
    z = np.load(GOLDEN)
    seed, xs, cent, mask = _case(name)
    enc, seg, _, _ = _build(amp, seed, cuda)
    enc.eval(); seg.eval()
    enc.precision = seg.precision = "bf16"
    logits, ft, out = _run(enc, seg, xs, cent, mask, cuda)
    ref = torch.from_numpy(z[name + "__eval_logits"])
    assert _rel(logits, ref) < TOL_BF16_LOGITS
    assert _relnorm(logits, ref) < TOL_BF16_LOGITS / 2
    assert (logits.argmax(1).cpu() == ref.argmax(1)).float().mean().item() >= 0.97
    assert _rel(ft, z[name + "__eval_ft"]) < TOL_BF16_FT
    assert _rel(out[:, ::37, :], z[name + "__eval_enc_out_last"]) < TOL_BF16_LOGITS


def test_bf16_forward_ragged_blocks_and_training_refusal(amp, cuda):
    """Variable block sizes (test script shape, batch 1, rows not a multiple of the 128-point tile) on the tensor-core
    path against the fp32 path; the bf16 path is eval-only."""
    enc, seg, _, _ = _build(amp, 5, cuda)
    enc.eval(); seg.eval()
    g = torch.Generator().manual_seed(3)
    sizes = [2048, 2500, 2177]
    xs = [torch.rand(1, n, 9, generator=g) for n in sizes]
    cent = torch.rand(1, len(sizes), 2, generator=g)
    res = {}
    for prec in ("fp32", "bf16"):
        enc.precision = seg.precision = prec
        res[prec], _, _ = _run(enc, seg, xs, cent, None, cuda)
    assert res["bf16"].shape == (1, 5, sum(sizes))
    assert _rel(res["bf16"], res["fp32"]) < TOL_BF16_LOGITS
    enc.train(); seg.train()       # training ignores the precision switch and runs the fp32 path
    out, ft = enc(torch.rand(2, 256, 9, generator=g).to(cuda))
    assert out.requires_grad


def test_tc_linear_matches_bf16_rounded_matmul(amp, cuda):
    torch.manual_seed(0)
    for (C, R, K, N, relu, pool) in [(1, 128, 16, 16, False, False), (3, 200, 128, 128, True, False), (2, 256, 128, 256, True, True),
                                     (3, 77, 64, 128, True, True), (4, 1000, 32, 64, False, False)]:
        x = torch.randn(C, R, K, device=cuda); w = torch.randn(N, K, device=cuda) / K ** 0.5; b = torch.randn(N, device=cuda)
        got = amp.tc_linear(x, w, b, relu=relu, pool=pool)
        exp = x.bfloat16().float() @ w.bfloat16().float().t() + b
        if relu:
            exp = torch.relu(exp)
        if pool:
            exp = exp.max(dim=1).values
        assert (got - exp).abs().max().item() < 1e-4
    with pytest.raises(RuntimeError, match="multiple of 16"):
        amp.tc_linear(torch.zeros(1, 8, 24, device=cuda), torch.zeros(16, 24, device=cuda))


def test_linear_wgrad_entry_point(amp, cuda):
    """amp_wgrad_f32: tensor-core (split bf16) path for wide K, exact fp32 paths for narrow K and few rows."""
    torch.manual_seed(1)
    for (C, R, N, K, bias, tol) in [(4, 2048, 128, 64, True, 2e-5), (8, 2048, 256, 128, False, 2e-5), (3, 1000, 64, 128, True, 2e-5),
                                    (8, 2048, 64, 9, False, 2e-6), (2, 2048, 64, 3, True, 2e-6), (1, 32, 4096, 128, True, 2e-6),
                                    (1, 288, 768, 256, True, 3e-6), (2, 100, 64, 64, True, 2e-6)]:
        dy = torch.randn(C, R, N, device=cuda); a = torch.randn(C, R, K, device=cuda)
        out = amp.linear_wgrad(dy, a, bias=bias)
        dw = out[0] if bias else out
        ref = dy.double().reshape(-1, N).t() @ a.double().reshape(-1, K)
        assert float((dw.double() - ref).abs().max() / ref.abs().max()) < tol, (C, R, N, K)
        if bias:
            rb = dy.double().reshape(-1, N).sum(0)
            assert float((out[1].double() - rb).abs().max() / rb.abs().max()) < tol
        again = amp.linear_wgrad(dy, a, bias=bias)                       # deterministic: bit-identical on a second run
        assert torch.equal(dw, again[0] if bias else again)


def test_tile_pipeline_blocks_and_logits(amp, cuda):
    """configs[3] in miniature: windows -> constrained k-means (exactly 2048 per block) -> encoder over all blocks ->
    attention + head per window; every point gets a label and the block split is a partition of each window."""
    rng = np.random.default_rng(5)
    ks = [2, 3, 2]
    wins = [rng.random((k * 2048, 13), dtype=np.float32) for k in ks]
    pc = torch.from_numpy(np.concatenate(wins, 0)).to(cuda)
    offsets = np.concatenate([[0], np.cumsum([len(w) for w in wins])])
    feats = amp.gather_feats(pc, (0, 1, 9))
    labels, _, _ = amp.kmeans_constrained_windows(feats, offsets, ks, 2048, 2048)
    order, counts, xy = amp.regroup_windows(labels, offsets, ks, pc)
    counts = counts.cpu().numpy()
    for w, k in enumerate(ks):
        assert (counts[w, :k] == 2048).all()
        seg_order = order[offsets[w]:offsets[w + 1]].cpu().numpy()
        assert sorted(seg_order.tolist()) == list(range(offsets[w], offsets[w + 1]))       # partition of the window's rows
    enc, seg, _, _ = _build(amp, 7, cuda)
    enc.eval(); seg.eval()
    grouped = pc.index_select(0, order)
    x9 = torch.cat((grouped[:, 0:3], grouped[:, 4:10]), 1).view(-1, 2048, 9)
    out, _ = enc(x9)
    b0 = 0
    for w, k in enumerate(ks):
        e = out[b0:b0 + k]
        gl = e[:, 0, :256].unsqueeze(1)                                   # [k, 1, 256]
        lo = e[:, :, 256:].reshape(1, k * 2048, 64)
        logits, _ = seg(gl, lo, xy[w:w + 1, :k, :], [2048] * k, None)
        assert tuple(logits.shape) == (1, 5, k * 2048) and torch.isfinite(logits).all()
        b0 += k


def test_training_shape_of_the_script_nine_windows_padded(amp, cuda):
    """The real training shape (train_pointnet-attention.py:396-470 on a collate_seq_padd batch): W = 9 windows of 2048
    points per sample, the trailing windows of a sample are replicas of its last real window with targets -1
    (collate_fns.py:42-44), CE with ignore_index -1 + 0.001 reg, backward, two Adam steps. Logits / loss against the
    CPU oracle, every parameter gets a finite gradient, padded windows contribute nothing to the loss gradient."""
    B, N, W, seed = 4, 2048, 9, 71
    enc, seg, sd_e, sd_s = _build(amp, seed, cuda)
    xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
    # clouds that differ from each other, so that the T-Net BatchNorms over the B pooled rows are well conditioned
    # (see test_all_gradients_match_oracle)
    g = torch.Generator().manual_seed(9)
    xs = [x * (0.15 + 0.85 * torch.rand(B, 1, 9, generator=g)) + 0.3 * torch.randn(B, 1, 9, generator=g) for x in xs]
    cent = torch.stack([x[:, :, :2].mean(1) for x in xs], 1)
    real = [5, 9, 3, 7]                                      # real windows per sample; the rest replicate the last real one
    for b in range(B):
        for w in range(real[b], W):
            xs[w][b] = xs[real[b] - 1][b]
    tg = torch.randint(0, 5, (B, W * N), generator=torch.Generator().manual_seed(2))
    for b in range(B):
        tg[b, real[b] * N:] = -1
    enc.train(); seg.train()
    opt_e = torch.optim.Adam(enc.parameters(), lr=1e-3); opt_s = torch.optim.Adam(seg.parameters(), lr=1e-3)
    logits, ft, _ = _run(enc, seg, xs, cent, None, cuda)
    ce = torch.nn.CrossEntropyLoss(weight=torch.tensor([1., 2., 2., 1., 1.], device=cuda), ignore_index=-1)
    loss = ce(logits, tg.to(cuda)) + 0.001 * torch.norm(torch.eye(64, device=cuda) - torch.bmm(ft, ft.transpose(2, 1)))
    loss.backward()
    o_logits, o_ft = nn_oracle.forward_windows(sd_e, sd_s, xs, cent, None, training=True, stats_enc={}, stats_seg={})
    o_loss, _, _ = nn_oracle.train_step_loss(o_logits, tg, o_ft)
    assert tuple(logits.shape) == (B, 5, W * N)
    assert _rel(logits, o_logits) < TOL_LOGITS
    assert abs(float(loss.detach()) - float(o_loss)) < 1e-4 * abs(float(o_loss))
    for m in (enc, seg):
        for k, p in m.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
    before = seg.conv_4.weight.detach().clone()
    opt_e.step(); opt_s.step()
    assert not torch.equal(before, seg.conv_4.weight.detach())
    assert int(enc.bn_1.num_batches_tracked) == 7 + W        # one BatchNorm update per encoder call (quirk 7 of SURVEY 3.5)


def test_seg_reads_encoder_output_slices_in_place(amp, cuda):
    """One block per window: out[:, :, -64:] / out[:, 0, :-64] (train_pointnet-attention.py:427-433) go to the head as strided
    views (gl_ld / lo_ld of amp_seg_fwd) and must give exactly what dense copies give, forward and backward."""
    B, N, seed = 4, 512, 71
    enc, seg, _, _ = _build(amp, seed, cuda)
    xs, cent = nn_params.synthetic_blocks(B, N, 1, seed)
    cent = cent.to(cuda)
    results = []
    for dense in (False, True):
        enc.train(); seg.train()
        enc.zero_grad(); seg.zero_grad()
        out, _ = enc(xs[0].to(cuda))
        lo = out[:, :, -64:]
        gl = torch.transpose(out[:, 0, :-64].view(-1, 1, 256), 0, 1)
        if dense:
            lo, gl = lo.contiguous(), gl.contiguous()
        else:
            assert not lo.is_contiguous()
        logits, _ = seg(gl, lo, cent, [N], None)
        logits.square().mean().backward()
        results.append((logits.detach().clone(), [p.grad.clone() for p in list(enc.parameters()) + list(seg.parameters())]))
        # running statistics advance once per pass: restore them so that both passes see the same state
        enc.load_state_dict(nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed))
        seg.load_state_dict(nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1))
    assert torch.equal(results[0][0], results[1][0])
    for a, b in zip(results[0][1], results[1][1]):
        assert _rel(a, b) < 1e-6


def test_zero_copy_gradient_sinks_match_autograd_gradients(amp, cuda):
    """GradAllReduce(zero_copy=True): p.grad is a slice of the flat all-reduce buffer; the first backward of a step writes
    it in place, the second encoder call of the same step (two windows) is accumulated by autograd, and the next step
    overwrites -- same numbers as the ordinary autograd path."""
    B, N, seed = 4, 512, 81
    xs, cent = nn_params.synthetic_blocks(B, N, 2, seed)

    def run(zero_copy):
        enc, seg, _, _ = _build(amp, seed, cuda)
        enc.train(); seg.train()
        params = list(enc.parameters()) + list(seg.parameters())
        red = amp.GradAllReduce(params, world=1, zero_copy=True) if zero_copy else None
        for _ in range(2 if zero_copy else 1):            # the second step must overwrite, not add
            enc.load_state_dict(nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed))
            seg.load_state_dict(nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1))
            logits, ft, _ = _run(enc, seg, xs, cent, None, cuda)
            (logits.square().mean() + 0.01 * ft.square().mean()).backward()
            if zero_copy:
                red.begin_step()                           # what all_reduce() does at the end of a step
        if zero_copy:
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))
            out = [v.clone() for v in red.views]
            red.detach()
            assert all(p.grad is None for p in params)
            return out
        return [p.grad.clone() for p in params]

    for a, b in zip(run(False), run(True)):
        assert _rel(a, b) < 1e-6


def _grads(mods):
    return {"%d.%s" % (i, k): p.grad.detach().clone() for i, m in enumerate(mods) for k, p in m.named_parameters()}


def test_device_dropout_offset_shifts_every_dropout_seed(amp, cuda):
    """amp_set_dropout_offset: seed_by_value + *device_word must be what every dropout site of forward AND backward uses.
    (seed a, offset b - a) has to reproduce (seed b, no offset) bit for bit, at a shape served by the tensor-core kernels."""
    B, N, W, seed = 4, 1024, 2, 63
    enc, seg, _, _ = _build(amp, seed, cuda, dropout=0.3)
    xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
    enc.train(); seg.train()

    def draw(s):                                             # the by-value seed SegmentationWithAttention.forward will draw
        torch.manual_seed(s)
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    def run(s):
        for m in (enc, seg):
            m.zero_grad(set_to_none=True)
        torch.manual_seed(s)
        logits, ft, _ = _run(enc, seg, xs, cent, None, cuda)
        (logits.square().mean() + ft.sum() * 1e-3).backward()
        return logits.detach().clone(), _grads((enc, seg))

    sa, sb = draw(11), draw(12)
    word = torch.tensor([(sb - sa) % (1 << 63)], dtype=torch.int64, device=cuda)
    assert (sa + int(word.item())) % (1 << 64) == sb % (1 << 64) or sb < sa
    if sb < sa:                                              # keep the sum free of 64-bit wrap-around games: swap the roles
        sa, sb = sb, sa
        word = torch.tensor([sb - sa], dtype=torch.int64, device=cuda)
        first, second = 12, 11
    else:
        first, second = 11, 12
    lb, gb = run(second)
    la0, _ = run(first)
    assert _rel(la0, lb) > 1e-3                              # different seeds, different masks
    amp._lib.check(amp._lib.lib().amp_set_dropout_offset(word.data_ptr()))
    try:
        la, ga = run(first)
    finally:
        amp._lib.check(amp._lib.lib().amp_set_dropout_offset(None))
    assert torch.equal(la, lb)
    for k in gb:
        assert torch.equal(ga[k], gb[k]), k


def test_graphed_training_step_redraws_dropout_and_trains(amp, cuda):
    """amp.GraphedStep: zero_grad + forward + loss + backward + 2 x Adam as one CUDA graph; every replay draws new masks."""
    B, N, seed = 8, 1024, 65
    enc, seg, _, _ = _build(amp, seed, cuda, dropout=0.3)
    xs, cent = nn_params.synthetic_blocks(B, N, 1, seed)
    x, cent = xs[0].to(cuda), cent.to(cuda)
    tg = torch.randint(0, 5, (B, N), device=cuda)
    enc.train(); seg.train()
    opt_e = torch.optim.Adam(enc.parameters(), lr=1e-3, fused=True, capturable=True)
    opt_s = torch.optim.Adam(seg.parameters(), lr=1e-3, fused=True, capturable=True)
    keep = {}

    def step():
        opt_e.zero_grad(set_to_none=True); opt_s.zero_grad(set_to_none=True)
        logits, ft, _ = _run(enc, seg, [x], cent, None, cuda)
        loss = torch.nn.functional.cross_entropy(logits, tg) + 0.001 * torch.norm(torch.eye(64, device=cuda) - torch.bmm(ft, ft.transpose(2, 1)))
        loss.backward()
        opt_e.step(); opt_s.step()
        keep["loss"] = loss.detach(); keep["logits"] = logits.detach()

    g = amp.GraphedStep(step, device=cuda)
    try:
        w0 = seg.conv_3.weight.detach().clone()
        seen = []
        for _ in range(4):
            g()
            seen.append((float(keep["loss"]), keep["logits"].clone()))
        torch.cuda.synchronize()
    finally:
        g.close()
    assert all(np.isfinite(l) for l, _ in seen)
    assert _rel(seen[0][1], seen[1][1]) > 1e-3               # consecutive replays: different dropout masks
    assert not torch.equal(w0, seg.conv_3.weight)            # the optimizers ran inside the graph
    assert seen[-1][0] < seen[0][0] * 1.5                    # and did not blow the loss up


def test_fused_adam_matches_torch_adam(amp, cuda):
    """amp.FusedAdam (one launch for all tensors) against torch.optim.Adam with the reference's settings (train_...:141-142)."""
    torch.manual_seed(3)
    shapes = [(64, 12, 1), (64,), (256, 128, 1), (4096, 128), (5,), (1,), (129, 3), (768, 256)]
    pa = [torch.nn.Parameter(torch.randn(s, device=cuda)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = amp.FusedAdam(pa, lr=1e-3)
    ob = torch.optim.Adam(pb, lr=1e-3)
    for it in range(6):
        for a, b in zip(pa, pb):
            g = torch.randn_like(a) * (10.0 ** (it - 3))
            a.grad = g.clone(); b.grad = g.clone()
        if it == 3:
            pa[2].grad = None; pb[2].grad = None            # a tensor without a gradient is skipped, as in torch
        oa.step(); ob.step()
        oa.zero_grad(set_to_none=True); ob.zero_grad(set_to_none=True)
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (a.shape, float((a - b).abs().max()))


def test_streamed_forward_keeps_batches_in_flight_and_returns_each_batchs_logits(amp, cuda):
    """amp.StreamedForward (copy-in / run / copy-out streams, one captured forward per slot): seven different batches through
    three slots give, bit for bit, the logits of seven direct calls; slot reuse and bad inputs are refused."""
    B, N, seed = 4, 1024, 91
    enc, seg, _, _ = _build(amp, seed, cuda)
    enc.eval(); seg.eval()
    batches = []
    for i in range(7):
        xs, cent = nn_params.synthetic_blocks(B, N, 1, seed + i)
        batches.append((xs[0].contiguous().pin_memory(), cent.contiguous().pin_memory()))
    with torch.no_grad():
        want = [_run(enc, seg, [x], c, None, cuda)[0].cpu() for x, c in batches]
        assert not torch.equal(want[0], want[1])
        pipe = amp.StreamedForward(lambda x, c: _run(enc, seg, [x], c, None, cuda)[0], batches[0], device=cuda, depth=3)
        got = [lg.clone() for lg in pipe.run(batches)]
    assert len(got) == 7
    for g, w in zip(got, want):
        assert torch.equal(g, w)
    t = [pipe.submit(*batches[i]) for i in range(3)]
    with pytest.raises(RuntimeError, match="must be collected"):
        pipe.submit(*batches[3])
    assert torch.equal(pipe.result(t[1]), want[1])
    assert torch.equal(pipe.result(t[0]), want[0])
    t3 = pipe.submit(*batches[3])                              # slot 0 is free again
    with pytest.raises(KeyError):
        pipe.result(t[0])
    assert torch.equal(pipe.result(t[2]), want[2]) and torch.equal(pipe.result(t3), want[3])
    with pytest.raises(ValueError):
        pipe.submit(batches[0][0][:2], batches[0][1])


@pytest.mark.parametrize("B,W,masked", [(32, 1, False), (5, 3, True), (8, 9, True), (3, 20, False)])
def test_fused_attention_tail_matches_the_separate_launches(amp, cuda, B, W, masked):
    """Eval forward: positional encoding + attention + per-block bias in one launch (seg_tail_eval_kernel, grid barriers between
    its phases) against the five separate launches (AMP_DISABLE=seg_tail), several row tiles and masked windows included."""
    N, seed = 256, 300 + B + W
    enc, seg, _, _ = _build(amp, seed, cuda)
    enc.eval(); seg.eval()
    xs, cent = nn_params.synthetic_blocks(B, N, W, seed)
    mask = None
    if masked:
        mask = torch.zeros(B, W, dtype=torch.bool)
        mask[0, W - 1] = True
        mask[B - 1, 0] = True
    with torch.no_grad():
        n0 = amp._lib.path_count("seg_tail")
        fused = _run(enc, seg, xs, cent, mask, cuda)[0].clone()
        assert amp._lib.path_count("seg_tail") == n0 + 1
        amp._lib.set_disabled(["seg_tail"])
        try:
            plain = _run(enc, seg, xs, cent, mask, cuda)[0].clone()
            assert amp._lib.path_count("seg_tail") == n0 + 1
        finally:
            amp._lib.set_disabled(None)
    assert torch.isfinite(fused).all()
    assert _rel(fused, plain) < 2e-5
    assert (fused.argmax(1) == plain.argmax(1)).float().mean().item() > 0.9995

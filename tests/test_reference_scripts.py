"""The drop-in boundary seen from the reference's own scripts.

 * CPU, needs_reference: the UNMODIFIED train_loop (pointNet/self-attention/train_pointnet-attention.py:337-475, loaded
   from the reference tree by path) and the reference's collate_seq_padd (collate_fns.py:4-55) against their restatement
   in oracle/train_loop_oracle.py, both driving the reference modules on the same seeded synthetic batch: bit-identical
   batch tensors, losses, predictions and parameters after the Adam steps. This pins the restatement.
 * GPU: the restated loop driving the DROP-IN modules (hand-written CUDA behind the same constructors / forward
   signatures / autograd) against tests/golden/train_loop_reference.npz, which oracle/make_golden_loop.py recorded from
   the unmodified loop + unmodified modules: one training step (zero_grad, 9 encoder calls, attention head, CE + reg loss,
   backward, two Adam steps) and one eval step.
"""
import os

import numpy as np
import pytest
import torch

from oracle import make_golden_loop as mgl, nn_params, train_loop_oracle as tlo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_loop_reference.npz")


@pytest.mark.needs_reference
def test_restated_collate_and_train_loop_equal_the_unmodified_ones(reference):
    mod, model, coll = mgl.load_train_script()
    mod.device = "cpu"
    n_samples, seed = 2, 7
    # collate: same random draws, same tensors
    tlo.seed_all(seed)
    a = coll.collate_seq_padd(tlo.synthetic_samples(n_samples, seed))
    tlo.seed_all(seed)
    b = tlo.collate_seq_padd(tlo.synthetic_samples(n_samples, seed))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[2] == b[2] and torch.equal(a[3], b[3])
    assert tuple(a[0].shape) == (n_samples, 2048, 9, 9) and tuple(a[1].shape) == (n_samples, 2048, 9) and tuple(a[3].shape) == (n_samples, 9, 2)
    # one train + one eval step with the reference modules
    enc_a, seg_a = mgl.build(model, seed)
    enc_b, seg_b = mgl.build(model, seed)
    ra = mgl.run_case(mod.train_loop, coll.collate_seq_padd, enc_a, seg_a, n_samples, seed, task="segmentation")
    rb = mgl.run_case(tlo.train_loop, tlo.collate_seq_padd, enc_b, seg_b, n_samples, seed, device="cpu")
    for phase in ("train", "eval"):
        ma, ta, pa, _ = ra[phase]
        mb, tb, pb, _ = rb[phase]
        assert float(ma["ce_loss"]) == float(mb["ce_loss"]) and float(ma["reg_loss"]) == float(mb["reg_loss"]), phase
        assert torch.equal(ta, tb) and torch.equal(pa, pb), phase
    for (ka, pa), (kb, pb) in zip(enc_a.state_dict().items(), enc_b.state_dict().items()):
        assert ka == kb and torch.equal(pa, pb), ka
    for (ka, pa), (kb, pb) in zip(seg_a.state_dict().items(), seg_b.state_dict().items()):
        assert ka == kb and torch.equal(pa, pb), ka


@pytest.mark.needs_reference
def test_golden_of_the_unmodified_loop_is_reproduced_by_the_restatement_on_cpu(reference):
    """The committed fixture equals what the restated loop + reference modules give today (guards a stale fixture)."""
    model, _, _ = reference
    z = np.load(GOLDEN)
    for name, (n_samples, seed) in mgl.CASES.items():
        enc, seg = mgl.build(model, seed)
        r = mgl.run_case(tlo.train_loop, tlo.collate_seq_padd, enc, seg, n_samples, seed, device="cpu")
        for phase in ("train", "eval"):
            m, t, p, _ = r[phase]
            assert abs(float(m["ce_loss"]) - float(z["%s__%s_ce" % (name, phase)])) < 1e-6
            assert (p.numpy() == z["%s__%s_preds" % (name, phase)]).mean() > 0.9999


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mgl.CASES))
def test_drop_in_modules_under_the_reference_training_loop(amp, cuda, name):
    z = np.load(GOLDEN)
    n_samples, seed = mgl.CASES[name]
    enc = amp.BasePointNet(point_dimension=3, return_local_features=True, global_feat_dim=256, device=cuda)
    seg = amp.SegmentationWithAttention(256, 8, num_classes=5, local_dim=64, dropout=0.0, device=cuda)
    enc.load_state_dict(nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed), strict=True)
    seg.load_state_dict(nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1), strict=True)
    enc.to(cuda); seg.to(cuda)
    n0 = amp._lib.launch_count()
    r = mgl.run_case(tlo.train_loop, tlo.collate_seq_padd, enc, seg, n_samples, seed, device=cuda)
    assert amp._lib.launch_count() - n0 > 100
    for phase in ("train", "eval"):
        m, t, p, logits = r[phase]
        ce_ref, reg_ref = float(z["%s__%s_ce" % (name, phase)]), float(z["%s__%s_reg" % (name, phase)])
        assert abs(float(m["ce_loss"]) - ce_ref) < 2e-4 * abs(ce_ref), (phase, float(m["ce_loss"]), ce_ref)
        # (the feature transform comes out of BatchNorms over the 3 clouds of the batch: the least conditioned number of the step)
        assert abs(float(m["reg_loss"]) - reg_ref) < 2e-3 * abs(reg_ref), phase
        assert (t.numpy() == z["%s__%s_targets" % (name, phase)]).all()                    # same shuffles, same padding
        agree = (p.numpy() == z["%s__%s_preds" % (name, phase)]).mean()
        assert agree >= 0.999, (phase, agree)
        assert tuple(logits.shape) == (n_samples, 5, 9 * 2048)
    # after zero_grad -> backward -> two Adam steps: the first Adam step moves every weight by ~lr * sign(gradient), so the
    # comparison is on the update direction (a parameter whose gradient is ~0 may flip)
    init_e = nn_params.synthetic_state_dict(nn_params.encoder_shapes(), seed)
    init_s = nn_params.synthetic_state_dict(nn_params.seg_shapes(), seed + 1)
    same, total = 0, 0
    for tag, mod, init, keys in (("enc", enc, init_e, mgl.SAMPLED), ("seg", seg, init_s, mgl.SAMPLED_SEG)):
        for k in keys:
            ours = dict(mod.named_parameters())[k].detach().cpu().numpy().reshape(-1)[::7]
            ref = z["%s__param_%s_%s" % (name, tag, k)]
            start = init[k].numpy().reshape(-1)[::7]
            moved = np.abs(ref - start) > 0.2 * mgl.LR
            same += int((np.sign(ours - start)[moved] == np.sign(ref - start)[moved]).sum()); total += int(moved.sum())
            assert np.abs(ours - ref).max() < 2.5 * mgl.LR, k
    assert total > 1000 and same / total > 0.99, (same, total)
    assert int(enc.bn_1.num_batches_tracked) == int(z[name + "__nbt"]) == 7 + 9              # 9 encoder calls in the training step
    assert np.abs(enc.bn_6.running_mean.cpu().numpy() - z[name + "__rm_bn_6"]).max() < 1e-4

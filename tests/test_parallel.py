"""Host-side multi-GPU plumbing on CPU: window sharding and the flat gradient all-reduce (gloo, world size 2)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _shards(amp, blocks, world):
    b = amp.shard_windows(blocks, world)
    assert len(b) == world
    assert b[0][0] == 0 and b[-1][1] == len(blocks)
    for (a0, a1), (b0, b1) in zip(b, b[1:]):
        assert a1 == b0 and a0 <= a1 and b0 <= b1
    return b


def test_shard_windows_partitions_and_balances(amp):
    assert _shards(amp, [1] * 8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    b = _shards(amp, [9, 1, 1, 1, 9, 1, 1, 1], 2)
    loads = [sum([9, 1, 1, 1, 9, 1, 1, 1][s:e]) for s, e in b]
    assert max(loads) - min(loads) <= 9
    _shards(amp, [3], 4)                   # fewer windows than ranks: trailing ranks get empty ranges
    for blocks in ([9, 1, 1, 1], [1, 1, 1, 9], [5, 5, 1, 1, 1, 1, 1, 1], [18, 1, 1]):    # every rank non-empty whenever n >= world
        for world in (2, 3, 4):
            if len(blocks) >= world:
                assert all(e > s for s, e in _shards(amp, blocks, world)), (blocks, world)
    _shards(amp, [], 2)
    for world in (1, 2, 3, 8):
        _shards(amp, [((7 * i) % 18) + 1 for i in range(53)], world)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import importlib
    amp = importlib.import_module("3d-semantic-segmentation-amp-net_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[2].requires_grad_(False)
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = None                                   # a parameter that got no gradient on this rank
    red = amp.GradAllReduce(params, world)
    red.all_reduce()
    ok = torch.allclose(params[0].grad, torch.full((5, 3), (1 + world) / 2.0)) and torch.count_nonzero(params[1].grad) == 0 \
        and params[2].grad is None
    # zero-copy mode: .grad is a slice of the flat buffer, "backward" writes the sink in place, all_reduce is the collective alone
    zc = amp.GradAllReduce(params[:2], world, zero_copy=True)
    ok = ok and params[0].grad.data_ptr() == zc.views[0].data_ptr()
    params[0]._amp_grad_sink.fill_(float(10 * (rank + 1))); params[0]._amp_sink_written = True      # what modules._grad_targets +
    params[1]._amp_grad_sink.fill_(float(rank)); params[1]._amp_sink_written = True                 # the library's backward do
    zc.all_reduce()
    ok = ok and torch.allclose(params[0].grad, torch.full((5, 3), 10 * (1 + world) / 2.0)) \
        and torch.allclose(params[1].grad, torch.full((7,), (world - 1) / 2.0))
    zc.detach()
    ok = ok and params[0].grad is None and not hasattr(params[0], "_amp_grad_sink") and not hasattr(params[0], "_amp_sink_written")
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grad_all_reduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_zero_copy_survives_zero_grad_set_to_none(amp):
    """The reference's loop calls optimizer.zero_grad() every step (train_pointnet-attention.py:372-373; set_to_none=True by
    default), which detaches p.grad from the flat buffer. all_reduce() must still deliver the sum of all windows' gradients
    and leave p.grad aliasing the buffer; a parameter without a gradient in the step must read zero, not last step's value."""
    import importlib
    modules = importlib.import_module("3d-semantic-segmentation-amp-net_b200.modules")
    w = torch.nn.Parameter(torch.zeros(4, 3)); b = torch.nn.Parameter(torch.zeros(4)); idle = torch.nn.Parameter(torch.zeros(3))
    red = amp.GradAllReduce([w, b, idle], world=1, zero_copy=True)
    red._reduce_flat = lambda: None                                   # single process: the collective is the identity
    opt = torch.optim.SGD([w, b, idle], lr=1.0)
    for step in range(2):
        opt.zero_grad()                                               # .grad = None on every parameter
        assert w.grad is None
        # window 1: the library writes the sink in place and returns None to autograd
        grads, targets = modules._grad_targets([w, b])
        assert grads == [None, None]
        targets[0].fill_(1.0 + step); targets[1].fill_(2.0)
        # window 2: ordinary gradients; with .grad detached autograd stores them in a NEW tensor
        grads, targets = modules._grad_targets([w, b])
        targets[0].fill_(10.0); w.grad = grads[0]
        # (b gets no second contribution)
        if step == 0:
            idle.grad = torch.full((3,), 7.0)                         # a torch-side gradient outside the library
        red.all_reduce()
        assert w.grad.data_ptr() == red.views[0].data_ptr() and b.grad.data_ptr() == red.views[1].data_ptr()
        assert torch.equal(w.grad, torch.full((4, 3), 11.0 + step)) and torch.equal(b.grad, torch.full((4,), 2.0))
        assert torch.equal(idle.grad, torch.full((3,), 7.0 if step == 0 else 0.0))     # stale value cleared on step 2
    # without zero_grad: aliasing kept, an untouched parameter is zeroed rather than re-used
    grads, targets = modules._grad_targets([w])
    targets[0].fill_(3.0)
    red.all_reduce()
    assert torch.equal(w.grad, torch.full((4, 3), 3.0)) and torch.count_nonzero(b.grad) == 0


def test_grad_targets_first_write_then_accumulate(amp):
    """Host logic of the zero-copy gradient sinks (modules._grad_targets), no GPU: the first backward of a step gets the
    sink as its write target and returns None to autograd; a second call in the same step gets a fresh tensor (autograd
    adds it to .grad = the sink); begin_step() re-arms; frozen parameters and buffers are handled as before."""
    import importlib
    modules = importlib.import_module("3d-semantic-segmentation-amp-net_b200.modules")
    w = torch.nn.Parameter(torch.zeros(4, 3)); b = torch.nn.Parameter(torch.zeros(4)); frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
    running = torch.zeros(4); counter = torch.zeros((), dtype=torch.long)
    tensors = [w, b, frozen, running, counter]
    grads, targets = modules._grad_targets(tensors)                 # no sinks: ordinary gradients
    assert grads[0] is targets[0] and grads[1] is targets[1] and grads[2] is None and targets[2] is not None
    assert grads[3] is None and targets[3] is None and grads[4] is None and targets[4] is None
    red = amp.GradAllReduce([w, b, frozen], world=1, zero_copy=True)
    assert w.grad.data_ptr() == red.views[0].data_ptr() and frozen.grad is None
    assert (red.views[1].data_ptr() - red.flat.data_ptr()) % 256 == 0   # every slice starts on a 256-byte boundary of the buffer
    grads, targets = modules._grad_targets(tensors)                 # first backward of the step: written in place
    assert grads[0] is None and targets[0] is red.views[0] and grads[1] is None and targets[1] is red.views[1]
    grads, targets = modules._grad_targets(tensors)                 # second backward of the step: accumulated by autograd
    assert grads[0] is targets[0] and grads[0].data_ptr() != red.views[0].data_ptr()
    red.begin_step()
    grads, targets = modules._grad_targets(tensors)
    assert grads[0] is None and targets[0] is red.views[0]
    red.detach()
    grads, targets = modules._grad_targets(tensors)
    assert grads[0] is targets[0] and w.grad is None

"""Data-parallel plumbing (one process per GPU, torch.distributed): the reference has no multi-GPU code
("no DDP, no torch.distributed"), this is the build's own row 8(e).

  shard_windows()   inference: contiguous ranges of windows per rank, balanced by block count, no collective
  GradAllReduce     training: one flat fp32 gradient buffer, a single all-reduce (NCCL over NVLink on the GPU box,
                    gloo in the CPU tests), averaged, before the two Adam optimizers step
"""
import torch
import torch.distributed as dist


def shard_windows(blocks_per_window, world):
    """Split windows [0, n) into `world` contiguous ranges with near-equal total block counts.
    The partition unit is the window: the attention layer mixes the blocks of one window
    (pointnetAtt.py:187-190), so all its blocks stay on one rank. Returns a list of (begin, end)."""
    n = len(blocks_per_window)
    total = float(sum(blocks_per_window))
    bounds, acc, start = [], 0.0, 0
    for r in range(world):
        left = world - r - 1                                  # ranks still to be served after this one
        if r == world - 1:
            end = n
        else:
            # at least one window (when any is left), at most what leaves one window for every later rank (when n >= world)
            lo = min(start + 1, n)
            hi = max(lo, n - left)
            end = lo
            acc_r = acc + sum(blocks_per_window[start:lo])
            while end < hi and acc_r < total * (r + 1) / world - 1e-9:
                acc_r += blocks_per_window[end]
                end += 1
        acc += sum(blocks_per_window[start:end])
        bounds.append((start, end))
        start = end
    return bounds


class GradAllReduce:
    """Averages the gradients of `params` over the process group with ONE collective on a flat buffer.

    zero_copy=False: `all_reduce()` packs the `.grad` tensors into the flat buffer (one multi-tensor copy), reduces, and
    unpacks. zero_copy=True (parameters of the ampnet_b200 modules): every `p.grad` IS its slice of the flat buffer; the first
    backward of a step writes it in place, later backward calls of the same step (one encoder call per window) are added to
    it by autograd (`modules._grad_targets`), and `all_reduce()` is the collective alone and ends the step. No zero_grad is
    needed (the first backward of the next step overwrites), but the reference's loop calls `optimizer.zero_grad()` every
    step (train_pointnet-attention.py:372-373), whose default `set_to_none=True` detaches `.grad` from the buffer:
    `all_reduce()` therefore re-attaches every parameter before the collective -- it folds a detached `.grad` (the windows
    autograd accumulated outside the buffer) into the slice, zeroes the slices of parameters that received no gradient in
    the step, and points `.grad` back at the buffer, so both styles train on the same averaged gradients."""

    def __init__(self, params, world=None, group=None, zero_copy=False):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = world if world is not None else dist.get_world_size(group)
        self.zero_copy = zero_copy
        # every slice starts on a 256-byte boundary, like a tensor of its own (the library's kernels use 16-byte accesses on
        # gradient targets); the padding is reduced along with the rest and costs < 1 % of the buffer
        pad = lambda k: (k + 63) // 64 * 64
        n = sum(pad(p.numel()) for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += pad(p.numel())
        if zero_copy:
            for p, v in zip(self.params, self.views):
                p.grad = v
                p._amp_grad_sink = v
                p._amp_sink_written = False

    def detach(self):
        """Undo zero_copy: parameters get ordinary gradients again."""
        for p in self.params:
            if getattr(p, "_amp_grad_sink", None) is not None:
                del p._amp_grad_sink
                del p._amp_sink_written
                p.grad = None
        self.zero_copy = False

    def begin_step(self):
        """zero_copy: the next backward overwrites the gradients (all_reduce() calls this; call it yourself when a step
        ends without a collective, e.g. world size 1)."""
        for p in self.params:
            p._amp_sink_written = False

    def _reduce_flat(self):
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)       # averaged inside the collective
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)

    def _reattach(self):
        """zero_copy, before the collective: make every slice hold exactly this step's gradient and every `.grad` alias it."""
        stale, add_dst, add_src, copy_dst, copy_src = [], [], [], [], []
        for p, v in zip(self.params, self.views):
            g, written = p.grad, getattr(p, "_amp_sink_written", False)
            aliased = g is not None and g.data_ptr() == v.data_ptr()
            if aliased:
                if not written:                 # no backward touched it this step: last step's average is still in the slice
                    stale.append(v)
            elif g is None:
                if not written:
                    stale.append(v)
                p.grad = v
            else:                               # zero_grad(set_to_none=True) detached it and autograd made a new tensor
                if written:
                    add_dst.append(v); add_src.append(g)
                else:
                    copy_dst.append(v); copy_src.append(g)
                p.grad = v
        if stale:
            torch._foreach_zero_(stale)
        if add_dst:
            torch._foreach_add_(add_dst, add_src)
        if copy_dst:
            torch._foreach_copy_(copy_dst, copy_src)

    def all_reduce(self):
        """zero_copy: re-attach check + the collective. Otherwise pack (one multi-tensor copy), all-reduce, unpack (one
        multi-tensor copy): 4 launches + the collective instead of two small copies per parameter."""
        if self.zero_copy:
            self._reattach()
            self._reduce_flat()
            self.begin_step()
            return
        have = [(p, v) for p, v in zip(self.params, self.views) if p.grad is not None]
        missing = [v for p, v in zip(self.params, self.views) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)
        if have:
            torch._foreach_copy_([v for _, v in have], [p.grad for p, _ in have])
        self._reduce_flat()
        if have:
            torch._foreach_copy_([p.grad for p, _ in have], [v for _, v in have])
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()

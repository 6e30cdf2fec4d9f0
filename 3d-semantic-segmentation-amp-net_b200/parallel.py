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
    r = 1
    for i, b in enumerate(blocks_per_window):
        acc += b
        while r < world and acc >= total * r / world - 1e-9 and n - (i + 1) >= 0:
            bounds.append((start, i + 1))
            start = i + 1
            r += 1
    bounds.append((start, n))
    while len(bounds) < world:
        bounds.append((n, n))
    return bounds[:world]


class GradAllReduce:
    """Averages the gradients of `params` over the process group with ONE collective on a flat buffer."""

    def __init__(self, params, world=None, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = world if world is not None else dist.get_world_size(group)
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def all_reduce(self):
        """Pack (one multi-tensor copy), all-reduce, average, unpack (one multi-tensor copy): 4 launches + the collective
        instead of two small copies per parameter."""
        have = [(p, v) for p, v in zip(self.params, self.views) if p.grad is not None]
        missing = [v for p, v in zip(self.params, self.views) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)
        if have:
            torch._foreach_copy_([v for _, v in have], [p.grad for p, _ in have])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.mul_(1.0 / self.world)
        if have:
            torch._foreach_copy_([p.grad for p, _ in have], [v for _, v in have])
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()

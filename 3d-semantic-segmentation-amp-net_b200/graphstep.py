"""A whole training step as ONE CUDA graph.

The reference's `train_loop` (train_pointnet-attention.py:337-475) is eager Python; the drop-in modules work under it
unchanged. A step of this network is ~180 short dependent kernels, so replaying it as a graph removes the launch gaps
(measured on B200: see DESIGN.md). Two things make the step capturable:
  * dropout: amp_seg_fwd / amp_seg_bwd take their seed by value, which a graph would replay unchanged; `GraphedStep` owns a
    device word that the library adds to every dropout seed (amp_set_dropout_offset) and bumps it inside the graph;
  * the optimizers must be built with `capturable=True`.
"""
import torch

from . import _lib


class GraphedStep:
    """g = GraphedStep(step_fn); g() replays. `step_fn()` is one full step on static tensors (zero_grad(set_to_none=True),
    forward, loss, backward, [gradient all-reduce], optimizer steps) and must not synchronise with the host."""

    _GOLDEN = 0x1E3779B97F4A7C15          # odd 61-bit increment of the dropout offset (positive as an int64)

    def __init__(self, step_fn, device=None, warmup=3):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.offset = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.step_fn = step_fn
        lib = _lib.lib()
        _lib.check(lib.amp_set_dropout_offset(self.offset.data_ptr()))
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._one()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)     # warm-up work is done: host-side staging buffers may be rewritten during capture
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._one()

    def _one(self):
        self.offset.add_(self._GOLDEN)     # a new dropout mask per execution, eager or replayed
        self.step_fn()

    def __call__(self):
        self.graph.replay()

    def close(self):
        _lib.check(_lib.lib().amp_set_dropout_offset(None))

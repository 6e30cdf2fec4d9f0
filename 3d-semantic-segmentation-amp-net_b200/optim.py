"""Adam for the drop-in modules in ONE kernel launch per step.

`FusedAdam(params, lr=...)` is used where the reference builds `torch.optim.Adam(net.parameters(), lr=...)`
(train_pointnet-attention.py:141-142: default betas / eps, no weight decay, no amsgrad) and stepped / zeroed the same way
(:372-373, :468-469). One instance may hold the parameters of both networks. The ~110 parameter tensors are cut into
2048-element chunks, one CTA each (amp_adam_step of include/ampnet_b200.h); torch's fused multi-tensor Adam takes three
launches and 135 us per step for the same 1.2 M parameters on B200.
"""
import struct

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("FusedAdam: bad hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._chunk = int(_lib.lib().amp_adam_chunk_elems())
        self._tables = {}                       # group index -> (key, device table, n_chunks)

    def _state_of(self, p):
        st = self.state[p]
        if not st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _table(self, gi, ps):
        """Device table of chunk records for this group; rebuilt (in place: the pinned staging buffer and the device table
        are allocated once, so that a rebuild inside a CUDA-graph capture is one captured copy and no allocation) when a
        parameter or gradient tensor moved."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), p.numel()) for p in ps)
        ent = self._tables.get(gi)
        if ent is not None and ent["key"] == key:
            return ent["dev"], ent["n"]
        rec = bytearray()
        n_chunks = 0
        index_of = {id(q): i for i, q in enumerate(self.param_groups[gi]["params"])}
        for p in ps:
            st = self._state_of(p)
            n, off = p.numel(), 0
            while off < n:
                c = min(self._chunk, n - off)
                rec += struct.pack("<QQQQii", p.data_ptr() + 4 * off, p.grad.data_ptr() + 4 * off, st["exp_avg"].data_ptr() + 4 * off,
                                   st["exp_avg_sq"].data_ptr() + 4 * off, c, index_of[id(p)])
                off += c
                n_chunks += 1
        if ent is None or ent["host"].numel() < len(rec):
            cap = sum((g_p.numel() + self._chunk - 1) // self._chunk for g_p in self.param_groups[gi]["params"]) * 40
            ent = {"host": torch.empty(max(cap, len(rec)), dtype=torch.uint8).pin_memory(),
                   "dev": torch.empty(max(cap, len(rec)), dtype=torch.uint8, device=ps[0].device), "event": None}
            self._tables[gi] = ent
        capturing = torch.cuda.is_current_stream_capturing()
        if ent["event"] is not None and not capturing:
            ent["event"].synchronize()                 # the previous copy out of the staging buffer has run
        ent["host"][:len(rec)] = torch.frombuffer(rec, dtype=torch.uint8)
        ent["dev"].copy_(ent["host"], non_blocking=True)
        if not capturing:
            ent["event"] = torch.cuda.Event()
            ent["event"].record()
        ent["key"], ent["n"] = key, n_chunks
        all_ps = self.param_groups[gi]["params"]
        ent["active"] = None if len(ps) == len(all_ps) else torch.tensor([index_of[id(p)] for p in ps], dtype=torch.int64).pin_memory().to(
            ps[0].device, non_blocking=True)
        return ent["dev"], n_chunks

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or \
                        not p.grad.is_contiguous() or p.grad.is_sparse:
                    raise RuntimeError("FusedAdam: contiguous float32 CUDA parameters and dense gradients only (no CPU fallback)")
            steps = group.get("_amp_steps")               # updates done so far, per parameter (torch semantics), on the device:
            if steps is None:                             # a captured graph counts its own replays
                steps = torch.zeros(len(group["params"]), dtype=torch.int64, device=ps[0].device)
                group["_amp_steps"] = steps
            table, n_chunks = self._table(gi, ps)
            with torch.cuda.device(ps[0].device):
                _lib.check(lib.amp_adam_step(table.data_ptr(), n_chunks, steps.data_ptr(), float(group["lr"]), float(group["betas"][0]),
                                             float(group["betas"][1]), float(group["eps"]), _lib.stream_ptr()))
            active = self._tables[gi]["active"]
            if active is None:
                steps += 1
            else:
                steps.index_add_(0, active, torch.ones_like(active))
        return loss

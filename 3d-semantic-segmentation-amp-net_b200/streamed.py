"""Inference from host buffers with the copies hidden behind the compute.

The reference's test loop (test_pointnet-attention.py) moves a batch to the GPU, runs the two modules and moves the
logits back, one after the other. On B200 the forward of a 32 x 2048 batch takes 0.23 ms and the two PCIe copies another
0.07 ms plus a host round trip, so a serial loop leaves the GPU idle for a third of the time. `StreamedForward` keeps
`depth` batches in flight: three streams (copy in, run, copy out), one captured CUDA graph of the forward per slot, events
between them. Slot i's input copy runs while slot i-1 computes and slot i-2's logits travel back.

    pipe = StreamedForward(lambda x, c: forward(enc, seg, x, c), (x_host, c_host))
    tickets = [pipe.submit(xb, cb) for xb, cb in first_batches]      # at most `depth` uncollected tickets
    logits = pipe.result(ticket)                                     # pinned host tensor, valid until the slot is reused
"""
import torch


class StreamedForward:
    """`fn(*device_inputs)` -> tensor or tuple of tensors; static shapes (those of `example_inputs`), no host synchronisation
    inside, same rules as for any CUDA graph capture. Host inputs should be pinned (an unpinned source makes the copy
    synchronous and the pipeline serial)."""

    def __init__(self, fn, example_inputs, device=None, depth=3, warmup=2):
        if not torch.cuda.is_available():
            raise RuntimeError("StreamedForward needs a CUDA device")
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.depth = int(depth)
        self.s_in = torch.cuda.Stream(device=self.device)
        self.s_run = torch.cuda.Stream(device=self.device)
        self.s_out = torch.cuda.Stream(device=self.device)
        self._slots = []
        self._submitted = 0
        torch.cuda.synchronize(self.device)
        for _ in range(self.depth):
            dev_in = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in example_inputs]
            with torch.cuda.stream(self.s_run):
                for d, h in zip(dev_in, example_inputs):
                    d.copy_(h)
                for _ in range(warmup):
                    fn(*dev_in)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=self.s_run):
                out = fn(*dev_in)
            single = torch.is_tensor(out)
            outs = [out] if single else list(out)
            host_out = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
            self._slots.append({"dev_in": dev_in, "graph": graph, "outs": outs, "host_out": host_out, "single": single,
                                "ev_in": torch.cuda.Event(), "ev_run": torch.cuda.Event(), "ev_out": torch.cuda.Event(),
                                "ticket": None, "collected": True})
        torch.cuda.synchronize(self.device)

    def submit(self, *host_inputs):
        """Enqueue one batch; returns its ticket. Never blocks on the GPU."""
        slot = self._slots[self._submitted % self.depth]
        if not slot["collected"]:
            raise RuntimeError("StreamedForward: result(%d) must be collected before its slot is reused (depth = %d)"
                               % (slot["ticket"], self.depth))
        if len(host_inputs) != len(slot["dev_in"]):
            raise ValueError("expected %d inputs" % len(slot["dev_in"]))
        for d, h in zip(slot["dev_in"], host_inputs):
            if tuple(h.shape) != tuple(d.shape) or h.dtype != d.dtype:
                raise ValueError("input of shape %s / %s where the captured forward takes %s / %s"
                                 % (tuple(h.shape), h.dtype, tuple(d.shape), d.dtype))
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(slot["ev_run"])            # the slot's previous run has consumed its inputs
            for d, h in zip(slot["dev_in"], host_inputs):
                d.copy_(h, non_blocking=True)
            slot["ev_in"].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(slot["ev_in"])
            self.s_run.wait_event(slot["ev_out"])           # the slot's previous outputs have left the device
            slot["graph"].replay()
            slot["ev_run"].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["ev_run"])
            for h, o in zip(slot["host_out"], slot["outs"]):
                h.copy_(o, non_blocking=True)
            slot["ev_out"].record(self.s_out)
        slot["ticket"] = self._submitted
        slot["collected"] = False
        self._submitted += 1
        return slot["ticket"]

    def result(self, ticket):
        """Block until the batch's outputs are in pinned host memory; the tensors stay valid until `depth` more submits."""
        slot = self._slots[ticket % self.depth]
        if slot["ticket"] != ticket:
            raise KeyError("ticket %d is not in flight" % ticket)
        slot["ev_out"].synchronize()
        slot["collected"] = True
        return slot["host_out"][0] if slot["single"] else tuple(slot["host_out"])

    def run(self, batches):
        """Generator over host outputs for an iterable of host input tuples, `depth` batches in flight. The yielded tensors
        are the slots' pinned buffers: consume (or copy) each before asking for the next `depth`-th one."""
        pending = []
        for b in batches:
            if len(pending) == self.depth:
                yield self.result(pending.pop(0))
            pending.append(self.submit(*b))
        for t in pending:
            yield self.result(t)

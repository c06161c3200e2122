"""Host-buffer pipeline around ``BVRNNCodecModel``: encode -> decode of a sequence of HOST batches with the copies of
neighbouring batches hidden under compute.

The blocking facade calls (``model.encode(x_cpu, bitrate)`` / ``model.decode(codes_cpu, length)``, reference
bvrnn_codec_model.py:44-71 with CPU tensors) expose every copy: H2D of the audio, D2H + H2D of the codes, D2H of the
decoded audio (~10 ms of a 223 ms step at B = 256 x 10 s).  A service that codes batch after batch does not need to:
the C ABI's device entry points only enqueue (ABI 3), so batch k + 1 is copied in and batch k - 1 copied out while batch
k computes.  Three streams, ``depth`` slots:

    copy-in     x_dev[slot] <- x_cpu (pinned)                                   after the slot's previous compute
    compute     codes = model.encode(x_dev[slot], bitrate); wav = model.decode(codes, L)
    copy-out    codes_cpu[slot], wav_cpu[slot] (pinned) <- codes, wav           after the compute of this batch

``submit`` returns a ticket, ``result(ticket)`` blocks until that batch's outputs are in host memory.  Outputs are the
same tensors the blocking calls return (codes float32 [B, L // 256, 64], audio float32 [B, L]).
"""
from __future__ import annotations

import torch


class HostPipeline:
    def __init__(self, model, depth: int = 2):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one batch computing, one in flight on the copy engines)")
        self.model = model
        self.depth = depth
        dev = model.device
        self._dev = dev
        self._in = torch.cuda.Stream(device=dev)
        self._compute = torch.cuda.Stream(device=dev)
        self._out = torch.cuda.Stream(device=dev)
        self._slots = [dict(x=None, codes=None, wav=None, busy=False,
                            ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event(), ev_out=torch.cuda.Event())
                       for _ in range(depth)]
        self._next = 0
        self._tickets = {}

    def submit(self, x_cpu: torch.Tensor, bitrate: float) -> int:
        """x_cpu: (batch, samples) float32 host tensor (page-locked for an asynchronous copy)."""
        if x_cpu.device.type != "cpu" or x_cpu.dtype != torch.float32 or x_cpu.dim() != 2:
            raise ValueError("x_cpu must be a 2-D float32 CPU tensor")
        ticket = self._next
        self._next += 1
        sl = self._slots[ticket % self.depth]
        if sl["busy"]:
            raise RuntimeError("HostPipeline: collect result(%d) before submitting batch %d" % (ticket - self.depth, ticket))
        B, L = x_cpu.shape
        T = L // self.model.conf["hopsize"]
        Z = self.model.conf["z_dim"]
        if sl["x"] is None or tuple(sl["x"].shape) != (B, L):
            sl["x"] = torch.empty(B, L, device=self._dev, dtype=torch.float32)
            sl["codes"] = torch.empty(B, T, Z, dtype=torch.float32).pin_memory()
            sl["wav"] = torch.empty(B, L, dtype=torch.float32).pin_memory()
        with torch.cuda.stream(self._in):
            self._in.wait_event(sl["ev_done"])                 # the slot's previous batch has been read by its encode
            sl["x"].copy_(x_cpu, non_blocking=True)
            sl["ev_in"].record(self._in)
        with torch.cuda.stream(self._compute):
            self._compute.wait_event(sl["ev_in"])
            codes = self.model.encode(sl["x"], bitrate)
            wav = self.model.decode(codes, L)
            sl["ev_done"].record(self._compute)
        with torch.cuda.stream(self._out):
            self._out.wait_event(sl["ev_done"])
            codes.record_stream(self._out)
            wav.record_stream(self._out)
            sl["codes"].copy_(codes, non_blocking=True)
            n = wav.shape[1]
            sl["wav"][:, :n].copy_(wav, non_blocking=True)
            sl["ev_out"].record(self._out)
        sl["busy"] = True
        self._tickets[ticket] = n
        return ticket

    def result(self, ticket: int):
        """Blocks until batch `ticket` is in host memory; returns (codes_cpu, wav_cpu) views of the slot's pinned buffers
        (valid until the slot is reused by submit number ticket + depth)."""
        n = self._tickets.pop(ticket)
        sl = self._slots[ticket % self.depth]
        sl["ev_out"].synchronize()
        sl["busy"] = False
        return sl["codes"], sl["wav"][:, :n]

    def close(self):
        for s in (self._in, self._compute, self._out):
            s.synchronize()

"""Utterance sharding across the GPUs of one box (one process per GPU, torch.distributed).

Utterances are independent (reference bvrnn.py:186-206 works row-wise, the vocoder per item), so
a batch is split into contiguous shards, each rank runs the whole encode->decode path on its
shard with replicated weights, and the only collective is the final gather of codes / audio
(NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int):
    """Contiguous, balanced split: the first (n_items % world_size) ranks get one extra item."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def local_shard(x: torch.Tensor, group=None) -> torch.Tensor:
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(x.shape[0], ws, rk)
    return x[lo:hi]


class _Gather:
    """A gather in flight (see gather_shards_async): result() orders the current stream after the collective."""

    def __init__(self, work, out, sizes, max_n):
        self.work, self.out, self.sizes, self.max_n = work, out, sizes, max_n

    def result(self) -> torch.Tensor:
        if self.work is not None:
            self.work.wait()
        if self.sizes is None:
            return self.out
        parts = [self.out[r * self.max_n: r * self.max_n + (hi - lo)] for r, (lo, hi) in enumerate(self.sizes)]
        return torch.cat(parts, 0)


def gather_shards_async(local: torch.Tensor, n_total: int, group=None) -> _Gather:
    """Starts the gather of every rank's shard (dim 0) and returns at once; the collective is ordered after the work
    already queued on the current stream and runs next to whatever is queued afterwards (e.g. the gather of the codes
    while the decoder runs, the gather of one half of the audio while the vocoder works on the other half).
    Equal shards go straight into the output (no padding, no concatenation pass); ragged shards are padded."""
    ws = dist.get_world_size(group)
    if ws == 1:
        return _Gather(None, local, None, 0)
    sizes = [shard_bounds(n_total, ws, r) for r in range(ws)]
    max_n = max(hi - lo for lo, hi in sizes)
    ragged = any(hi - lo != max_n for lo, hi in sizes)
    src = local.contiguous()
    if ragged:
        src = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        src[: local.shape[0]] = local
    out = torch.empty((ws * max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    work = dist.all_gather_into_tensor(out, src, group=group, async_op=True)
    return _Gather(work, out, sizes if ragged else None, max_n)


def gather_shards(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All ranks receive the concatenation of every rank's shard (dim 0), ragged shards allowed."""
    return gather_shards_async(local, n_total, group).result()


class ShardedCodec:
    """Runs ``codec.encode`` / ``codec.decode`` on this rank's shard and gathers the results.

    ``codec`` is any object with the reference API (``encode(x, bitrate)``, ``decode(codes, length)``).
    """

    def __init__(self, codec, group=None):
        self.codec = codec
        self.group = group

    def encode(self, x_full: torch.Tensor, bitrate, gather: bool = True):
        codes = self.codec.encode(local_shard(x_full, self.group), bitrate)
        return gather_shards(codes, x_full.shape[0], self.group) if gather else codes

    def decode(self, codes_full: torch.Tensor, length: int, gather: bool = True):
        wav = self.codec.decode(local_shard(codes_full, self.group), length)
        return gather_shards(wav, codes_full.shape[0], self.group) if gather else wav

    def forward(self, x_full: torch.Tensor, bitrate, gather: bool = True):
        xl = local_shard(x_full, self.group)
        wav = self.codec.decode(self.codec.encode(xl, bitrate), x_full.shape[1])
        return gather_shards(wav, x_full.shape[0], self.group) if gather else wav

    __call__ = forward

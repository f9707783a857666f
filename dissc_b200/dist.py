"""Utterance sharding across the GPUs of one box (replaces the reference's
``Pool(8, init_worker)`` + id-queue, sr/inference.py:288-292,351-359).

One process per GPU (torchrun).  Utterances are independent, weights are
replicated, so the model has no exchange step: the only collectives are the
batch ``scatter`` of packed inputs from rank 0 and the ``gather`` of waveforms
back (NCCL over NVLink/NVSwitch; ``gloo`` in the CPU tests).  Work is balanced
by length: utterances are sorted by frame count and dealt round-robin.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None):
    """(rank, world, local_rank); initialises torch.distributed iff WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_by_length(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Indices per rank: sort by length (desc), deal round-robin in serpentine order so the
    per-rank sum of frames (work is proportional to sum T) is balanced."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    for pos, idx in enumerate(order):
        rnd, k = divmod(pos, world)
        r = k if rnd % 2 == 0 else world - 1 - k
        shards[r].append(idx)
    return shards


def pack_batch(codes: Sequence[torch.Tensor], f0s: Sequence[torch.Tensor], spkrs: Sequence[int], T: int):
    """Ragged utterances -> padded (code int64 (B,T), f0 fp32 (B,T), spkr int64 (B), lengths int32 (B))."""
    B = len(codes)
    code = torch.zeros(B, T, dtype=torch.int64)
    f0 = torch.zeros(B, T, dtype=torch.float32)
    lengths = torch.zeros(B, dtype=torch.int32)
    for i, (c, f) in enumerate(zip(codes, f0s)):
        n = int(c.numel())
        code[i, :n] = c
        f0[i, :n] = f.reshape(-1)[:n]
        lengths[i] = n
    return code, f0, torch.as_tensor(list(spkrs), dtype=torch.int64), lengths


class ShardedBatch:
    """scatter -> local forward -> gather for one padded batch whose size is a multiple of world."""

    def __init__(self, rank: int, world: int, device: torch.device):
        self.rank, self.world, self.device = rank, world, device

    def scatter(self, code, f0, spkr, lengths, B_local: int, T: int):
        """rank 0 passes device tensors of the full batch (world*B_local rows); others pass None."""
        dev = self.device
        out = (torch.empty(B_local, T, dtype=torch.int64, device=dev),
               torch.empty(B_local, T, dtype=torch.float32, device=dev),
               torch.empty(B_local, dtype=torch.int64, device=dev),
               torch.empty(B_local, dtype=torch.int32, device=dev))
        if self.world == 1:
            return code, f0, spkr, lengths
        srcs = (code, f0, spkr, lengths)
        for dst, src in zip(out, srcs):
            chunks = list(src.chunk(self.world, dim=0)) if self.rank == 0 else None
            dist.scatter(dst, chunks, src=0)
        return out

    def gather(self, y_local: torch.Tensor):
        """-> on rank 0 the (world*B_local, ...) tensor, None elsewhere."""
        if self.world == 1:
            return y_local
        # NCCL has no 16-bit integer type ("Unconvertible NCCL type Short"): int16 waveforms travel as raw bytes
        dtype = y_local.dtype
        raw = dtype in (torch.int16, torch.uint16)
        y = y_local.contiguous()
        if raw:
            y = y.view(torch.uint8)
        bufs = [torch.empty_like(y) for _ in range(self.world)] if self.rank == 0 else None
        dist.gather(y, bufs, dst=0)
        if self.rank != 0:
            return None
        out = torch.cat(bufs, dim=0)
        return out.view(dtype) if raw else out

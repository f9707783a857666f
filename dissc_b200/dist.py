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


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


def packed_layout(B: int, T: int):
    """Byte offsets of one rank's inputs inside its packed row: code int64 (B,T) | f0 fp32 (B,T) | spkr int64 (B) |
    lengths int32 (B), every section 16-byte aligned -> ({name: (offset, nbytes)}, row_bytes)."""
    off, lay = 0, {}
    for name, nbytes in (("code", B * T * 8), ("f0", B * T * 4), ("spkr", B * 8), ("lengths", B * 4)):
        lay[name] = (off, nbytes)
        off = _align(off + nbytes)
    return lay, off


def pack_inputs(code, f0, spkr, lengths, world: int) -> torch.Tensor:
    """Full-batch tensors (world*B rows, on rank 0) -> ONE uint8 tensor (world, row_bytes): row r is everything rank r
    needs, so the batch travels in a single scatter instead of four."""
    Bt, T = code.shape
    B = Bt // world
    lay, row = packed_layout(B, T)
    out = torch.zeros((world, row), dtype=torch.uint8, device=code.device)
    parts = {"code": code.to(torch.int64).reshape(world, B * T), "f0": f0.to(torch.float32).reshape(world, B * T),
             "spkr": spkr.to(torch.int64).reshape(world, B), "lengths": lengths.to(torch.int32).reshape(world, B)}
    for name, (o, n) in lay.items():
        out[:, o:o + n] = parts[name].contiguous().view(torch.uint8).reshape(world, n)
    return out


def unpack_inputs(row: torch.Tensor, B: int, T: int):
    """One rank's packed row (uint8, row_bytes) -> zero-copy views (code (B,T) int64, f0 (B,T) fp32, spkr (B) int64,
    lengths (B) int32)."""
    lay, _ = packed_layout(B, T)
    get = lambda name, dt: row[lay[name][0]:lay[name][0] + lay[name][1]].view(dt)
    return (get("code", torch.int64).view(B, T), get("f0", torch.float32).view(B, T), get("spkr", torch.int64),
            get("lengths", torch.int32))


class ScatterGatherPipeline:
    """The N-GPU data path of the north-star: rank 0 owns the inputs and the outputs of every batch (it replaces the
    parent process of sr/inference.py:288-292,351-359, which hands utterance indices to eight workers and collects
    their wav files).  Per step

        scatter (ONE packed buffer)  ->  forward on every rank, int16 written by the last kernel straight into the
        gather's send buffer  ->  gather of the int16 waveforms to rank 0

    with the gather on a side stream and its own communicator, double-buffered, so step i's gather runs under step
    i+1's scatter + forward; nothing on the compute stream ever waits for a collective of the same step except the
    (0.2 MB) scatter.  ``forward_i16(code, f0, spkr, lengths, out)`` must write int16 (B, hop*T) into ``out``.

    ``step`` returns the buffer index; ``wait(i)`` makes the current stream wait for that step's gather, after which
    (on rank 0) ``gathered[i]`` is the (world, B, hop*T) int16 result.  Works on CPU tensors with gloo (no streams)."""

    def __init__(self, rank: int, world: int, device: torch.device, B_local: int, T: int, hop: int, forward_i16,
                 depth: int = 2, gather_group=None):
        self.rank, self.world, self.device = rank, world, device
        self.B, self.T, self.hop, self.depth = B_local, T, hop, depth
        self.forward_i16 = forward_i16
        _, self.row_bytes = packed_layout(B_local, T)
        self.cuda = device.type == "cuda"
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=device)
        self.recv = [mk((self.row_bytes,), torch.uint8) for _ in range(depth)]
        self.send = [mk((B_local, hop * T), torch.int16) for _ in range(depth)]
        self.gathered = [mk((world, B_local, hop * T), torch.int16) if rank == 0 else None for _ in range(depth)]
        self.n = 0
        self.gather_group = gather_group
        if world > 1 and gather_group is None and self.cuda:
            self.gather_group = dist.new_group(backend="nccl")   # own communicator: gathers never queue behind scatters
        if self.cuda:
            self.side = torch.cuda.Stream(device=device)
            self.ev_fwd = [torch.cuda.Event() for _ in range(depth)]
            self.ev_gather = [torch.cuda.Event() for _ in range(depth)]

    def step(self, packed):
        """``packed``: pack_inputs(...) on rank 0 (uint8 (world, row_bytes) on ``device``), None elsewhere."""
        i = self.n % self.depth
        self.n += 1
        if self.cuda:
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self.ev_gather[i])          # send[i] / gathered[i] are free again (no-op the first time)
        if self.world == 1:
            row = packed[0]
        else:
            rows = [packed[r] for r in range(self.world)] if self.rank == 0 else None
            dist.scatter(self.recv[i], rows, src=0)
            row = self.recv[i]
        code, f0, spkr, lengths = unpack_inputs(row, self.B, self.T)
        # one rank: the "gather" is the identity, the forward writes the result buffer itself
        self.forward_i16(code, f0, spkr, lengths, self.gathered[i][0] if self.world == 1 else self.send[i])
        if self.world == 1:
            if self.cuda:
                self.ev_gather[i].record(torch.cuda.current_stream(self.device))
            return i
        # NCCL has no 16-bit integer type ("Unconvertible NCCL type Short"): int16 waveforms travel as raw bytes
        src = self.send[i].view(torch.uint8)
        bufs = [self.gathered[i][r].view(torch.uint8) for r in range(self.world)] if self.rank == 0 else None
        if self.cuda:
            self.ev_fwd[i].record(main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.ev_fwd[i])
                dist.gather(src, bufs, dst=0, group=self.gather_group)
                self.ev_gather[i].record(self.side)
        else:
            dist.gather(src, bufs, dst=0, group=self.gather_group)
        return i

    def wait(self, i: int):
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_event(self.ev_gather[i])

    def flush(self):
        for i in range(self.depth):
            self.wait(i)


class ShardedBatch:
    """scatter -> local forward -> gather for one padded batch whose size is a multiple of world (synchronous
    convenience form; ``ScatterGatherPipeline`` is the overlapped one)."""

    def __init__(self, rank: int, world: int, device: torch.device):
        self.rank, self.world, self.device = rank, world, device

    def scatter(self, code, f0, spkr, lengths, B_local: int, T: int):
        """rank 0 passes device tensors of the full batch (world*B_local rows); others pass None.  One packed
        collective (``pack_inputs``)."""
        if self.world == 1:
            return code, f0, spkr, lengths
        _, row_bytes = packed_layout(B_local, T)
        recv = torch.empty(row_bytes, dtype=torch.uint8, device=self.device)
        rows = None
        if self.rank == 0:
            packed = pack_inputs(code, f0.reshape(code.shape), spkr.reshape(-1), lengths, self.world)
            rows = [packed[r] for r in range(self.world)]
        dist.scatter(recv, rows, src=0)
        return unpack_inputs(recv, B_local, T)

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

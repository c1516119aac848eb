"""Multi-GPU plumbing: channels shard embarrassingly across ranks -- every
receiver is independent (all state lives in its own struct receiver, src/receiver.h:35-46), so
there is NO collective on the data path.  The only exchange is collecting the decoded-message
buffers: an all-gather of per-rank counts, then exact-size point-to-point sends of the 64-byte
records to the destination rank (NCCL over NVLink on GPUs; gloo on CPU for the host-logic tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

REC_BYTES = 64
_CH_WORD = 14   # int32 word of gais_msg.channel inside a record


def shard_channels(total_channels: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous channel range [first, first + count) owned by `rank` (remainder to low ranks)."""
    base, rem = divmod(total_channels, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


class _CudaView:
    """zero-copy torch view of library-owned device memory via __cuda_array_interface__"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_records(rx) -> torch.Tensor:
    """[n, 64] uint8 CUDA tensor aliasing the dense message array of rx's last run."""
    ptr, n = rx.device_messages()
    if n == 0:
        return torch.empty((0, REC_BYTES), dtype=torch.uint8, device="cuda")
    return torch.as_tensor(_CudaView(ptr, n * REC_BYTES), device="cuda").view(n, REC_BYTES)


def globalize_channels(records: torch.Tensor, first_channel: int) -> torch.Tensor:
    """local channel index -> global channel index, in a copy of the records."""
    out = records.clone()
    if out.numel():
        out.view(torch.int32)[:, _CH_WORD] += first_channel
    return out


def gather_records(records: torch.Tensor, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Collect every rank's [n_r, 64] uint8 records on `dst` in rank order (ranks own ascending
    channel ranges and each rank's array is (channel, end_bit)-sorted, so the result is in the
    canonical global order).  Returns the concatenation on dst, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
    counts = torch.zeros(world, dtype=torch.int64, device=records.device)
    dist.all_gather_into_tensor(counts, n_local, group=group)
    counts = counts.cpu().tolist()
    if rank == dst:
        out = torch.empty((sum(counts), REC_BYTES), dtype=torch.uint8, device=records.device)
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        ops = []
        for r in range(world):
            if counts[r] == 0:
                continue
            if r == dst:
                out[offs[r]:offs[r + 1]].copy_(records)
            else:
                ops.append(dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r, group=group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    if counts[rank] > 0:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, records.contiguous(), dst, group=group)]):
            req.wait()
    return None


class PendingGather:
    """An in-flight gather_records(): the sends/receives run on NCCL's stream while the caller goes
    on computing; wait() completes them and returns the concatenation on dst (None elsewhere)."""

    def __init__(self, reqs, out, keep):
        self._reqs, self._out, self._keep = reqs, out, keep

    def wait(self):
        for r in self._reqs:
            r.wait()
        self._reqs, self._keep = [], None
        return self._out


def gather_records_async(records: torch.Tensor, dst: int = 0, group=None) -> PendingGather:
    """Like gather_records(), but returns as soon as the transfers are enqueued.  `records` must not
    be modified until wait() (pass a private copy, e.g. the result of globalize_channels())."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
    counts = torch.zeros(world, dtype=torch.int64, device=records.device)
    dist.all_gather_into_tensor(counts, n_local, group=group)
    counts = counts.cpu().tolist()
    if rank == dst:
        out = torch.empty((sum(counts), REC_BYTES), dtype=torch.uint8, device=records.device)
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        ops = []
        for r in range(world):
            if counts[r] == 0:
                continue
            if r == dst:
                out[offs[r]:offs[r + 1]].copy_(records)
            else:
                ops.append(dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r, group=group))
        return PendingGather(dist.batch_isend_irecv(ops) if ops else [], out, records)
    ops = [dist.P2POp(dist.isend, records.contiguous(), dst, group=group)] if counts[rank] > 0 else []
    return PendingGather(dist.batch_isend_irecv(ops) if ops else [], None, records)


def reduce_totals(totals, dst: int = 0, group=None, device="cpu"):
    """sum of (ok, crcfail, sizefail) over ranks"""
    t = torch.tensor(list(totals), dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tuple(int(x) for x in t.cpu())

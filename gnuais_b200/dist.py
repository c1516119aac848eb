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


class PeerGatherUnavailable(RuntimeError):
    """raised on EVERY rank of the group when the CUDA IPC mapping cannot be set up"""


HDR_BYTES = REC_BYTES      # slice header: int64 count, int64 step, 48 bytes unused


class PeerGather:
    """Collect every rank's records on `dst` with NO collective and NO count exchange on the critical path.

    `dst` owns one receive buffer of `world` fixed-capacity SLICES (64-byte header + cap_records records);
    every other rank maps it through CUDA IPC.  Per step a rank copies its (channel, end_bit)-ordered records
    -- which already carry global channel numbers (gais_config.reserved[3]) -- straight into ITS slice over
    NVLink with cudaMemcpyAsync on a side stream (copy engines: no SMs, no NCCL proxy), then the header
    {count, step}.  The records are first copied to a device-local staging buffer, so nothing on the compute
    stream ever waits for another rank or for NVLink: the only thing it waits for is that local copy, before
    the next run overwrites the dense array.  A one-word
    NCCL all-reduce rides the SIDE stream after each copy, so that step k is complete on `dst` before any
    rank's step k + 1 transfer starts; ranks are otherwise free to drift apart by a step.

    One box only (CUDA IPC); needs the `cuda-python` bindings.
    """

    def __init__(self, cap_records: int, dst: int = 0, group=None):
        from cuda.bindings import runtime as rt   # noqa: F401  (fail here, not in the middle of a step)

        self._rt = rt
        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.cap = int(cap_records)
        self.slice_bytes = HDR_BYTES + self.cap * REC_BYTES
        self.ptr = 0
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._work = None
        self._copied = None
        self._keep = None
        self._stage = None
        self.step = 0
        # headers in flight: a small ring, so that step k + 1 never rewrites the words step k's copy is still reading
        self._hdr_host = torch.zeros((8, 8), dtype=torch.int64).pin_memory()
        self._hdr_dev = torch.zeros((8, 8), dtype=torch.int64, device=self.device)
        self._hdr_done = [None] * 8
        self._allocate()

    @staticmethod
    def _ck(res, what):
        err = res[0] if isinstance(res, tuple) else res
        if int(err) != 0:
            raise RuntimeError(f"{what} failed: CUDA error {int(err)}")
        return res[1] if isinstance(res, tuple) and len(res) > 1 else None

    def _allocate(self):
        rt = self._rt
        nbytes = self.world * self.slice_bytes
        blob = None
        if self.rank == self.dst:
            self.ptr = int(self._ck(rt.cudaMalloc(nbytes), "cudaMalloc"))
            self._ck(rt.cudaMemset(self.ptr, 0, nbytes), "cudaMemset")
            blob = bytes(self._ck(rt.cudaIpcGetMemHandle(self.ptr), "cudaIpcGetMemHandle").reserved)
        box = [blob]
        dist.broadcast_object_list(box, src=self.dst, group=self.group)
        ok = 1
        if self.rank != self.dst:
            h = rt.cudaIpcMemHandle_t()
            h.reserved = box[0]
            res = rt.cudaIpcOpenMemHandle(h, rt.cudaIpcMemLazyEnablePeerAccess)
            if int(res[0]) == 0:
                self.ptr = int(res[1])
            else:
                ok = 0
        # the mapping either works on every rank or the whole group gives up together (no rank may be
        # left waiting in a collective the others never enter)
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            if self.rank != self.dst and self.ptr:
                rt.cudaIpcCloseMemHandle(self.ptr)
            dist.barrier(group=self.group)               # every rank, whether its own mapping worked or not
            if self.rank == self.dst:
                rt.cudaFree(self.ptr)
            self.ptr = 0
            raise PeerGatherUnavailable("cudaIpcOpenMemHandle failed on at least one rank")

    def wait_source_free(self, stream=None):
        """make `stream` (default: the current one) wait until the last start()'s copy has read its source"""
        if self._copied is not None:
            (stream or torch.cuda.current_stream(self.device)).wait_event(self._copied)

    def start(self, records: torch.Tensor) -> None:
        """Enqueue the transfer of `records` ([n, 64] uint8 on this device, global channel numbers) into this rank's
        slice.  Returns at once; `records` must stay untouched until wait_source_free() / finish()."""
        rt = self._rt
        n = int(records.shape[0])
        if n > self.cap:
            raise RuntimeError(f"rank {self.rank}: {n} records exceed the slice capacity {self.cap}")
        self.step += 1
        k = self.step % 8
        if self._hdr_done[k] is not None:
            self._hdr_done[k].synchronize()       # eight steps back: long done unless the side stream is stuck
        self._hdr_host[k, 0], self._hdr_host[k, 1] = n, self.step
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        base = self.ptr + self.rank * self.slice_bytes
        if self._stage is None or self._stage.shape[0] < n:
            # local staging copy of the records (grown on demand, kept): the dense array is free again after a device-local
            # copy (~0.2 ms per 100 MB), not after the NVLink transfer -- with eight ranks pushing 4.9 GB into one GPU that
            # transfer takes ~5 ms, which every rank's next run used to wait for (8 GPUs: 31.7 ms per step instead of 27.5)
            self.copy_stream.synchronize()
            self._stage = torch.empty((max(n + n // 8, 1024), REC_BYTES), dtype=torch.uint8, device=self.device)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            if n > 0:
                self._ck(rt.cudaMemcpyAsync(self._stage.data_ptr(), records.data_ptr(), n * REC_BYTES, rt.cudaMemcpyKind.cudaMemcpyDefault,
                                            self.copy_stream.cuda_stream), "cudaMemcpyAsync (staging)")
            self._copied = torch.cuda.Event()
            self._copied.record(self.copy_stream)
            if self._work is not None:
                self._work.wait()                 # step k - 1 is complete on dst (side stream only)
            if n > 0:
                self._ck(rt.cudaMemcpyAsync(base + HDR_BYTES, self._stage.data_ptr(), n * REC_BYTES, rt.cudaMemcpyKind.cudaMemcpyDefault,
                                            self.copy_stream.cuda_stream), "cudaMemcpyAsync (peer)")
            self._hdr_dev[k].copy_(self._hdr_host[k], non_blocking=True)
            self._ck(rt.cudaMemcpyAsync(base, self._hdr_dev[k].data_ptr(), HDR_BYTES, rt.cudaMemcpyKind.cudaMemcpyDefault,
                                        self.copy_stream.cuda_stream), "cudaMemcpyAsync (peer header)")
            done = torch.cuda.Event()
            done.record(self.copy_stream)
            self._hdr_done[k] = done
            self._work = dist.all_reduce(self._flag, group=self.group, async_op=True)
        self._keep = records

    def finish(self):
        """Complete everything in flight.  On dst: (counts per rank, list of per-rank [count, 64] uint8 views of the
        slices -- valid until the next start()); None elsewhere."""
        with torch.cuda.stream(self.copy_stream):
            if self._work is not None:
                self._work.wait()
                self._work = None
        self.copy_stream.synchronize()
        self._keep = None
        if self.rank != self.dst:
            return None
        counts, views = [], []
        for r in range(self.world):
            base = self.ptr + r * self.slice_bytes
            hdr = torch.as_tensor(_CudaView(base, HDR_BYTES), device=self.device).view(torch.int64).cpu()
            n = int(hdr[0])
            counts.append(n)
            views.append(torch.as_tensor(_CudaView(base + HDR_BYTES, max(n, 1) * REC_BYTES), device=self.device)[: n * REC_BYTES]
                         .view(n, REC_BYTES))
        return counts, views

    def close(self):
        rt = self._rt
        if not self.ptr:
            return
        self.finish()
        torch.cuda.synchronize(self.device)
        if self.rank == self.dst:
            dist.barrier(group=self.group)          # nobody still has it mapped for writing
            self._ck(rt.cudaFree(self.ptr), "cudaFree")
        else:
            self._ck(rt.cudaIpcCloseMemHandle(self.ptr), "cudaIpcCloseMemHandle")
            dist.barrier(group=self.group)
        self.ptr = 0

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


class PeerGather:
    """Collect every rank's records on `dst` WITHOUT a collective on the payload: `dst` owns a receive
    buffer, every other rank maps it through CUDA IPC and copies its records straight into its slice
    over NVLink with one cudaMemcpyAsync on a side stream (copy engines: no SMs, no NCCL proxy, so
    the next step's kernels keep the GPU while the records move).  NCCL carries only the per-rank
    counts (8 bytes each) and a one-word completion all-reduce.

    One box only (CUDA IPC); needs the `cuda-python` bindings.  The buffer grows on demand: every rank
    sees the same counts, so every rank takes the same (re)allocation branch.
    """

    def __init__(self, dst: int = 0, group=None):
        from cuda.bindings import runtime as rt   # noqa: F401  (fail here, not in the middle of a step)

        self._rt = rt
        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.ptr, self.cap = 0, 0
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    @staticmethod
    def _ck(res, what):
        err = res[0] if isinstance(res, tuple) else res
        if int(err) != 0:
            raise RuntimeError(f"{what} failed: CUDA error {int(err)}")
        return res[1] if isinstance(res, tuple) and len(res) > 1 else None

    def _release(self):
        rt = self._rt
        if self.ptr:
            torch.cuda.synchronize(self.device)
            if self.rank == self.dst:
                dist.barrier(group=self.group)          # nobody still has it mapped for writing
                self._ck(rt.cudaFree(self.ptr), "cudaFree")
            else:
                self._ck(rt.cudaIpcCloseMemHandle(self.ptr), "cudaIpcCloseMemHandle")
                dist.barrier(group=self.group)
        self.ptr, self.cap = 0, 0

    def _allocate(self, cap_records: int):
        rt = self._rt
        self._release()
        nbytes = cap_records * REC_BYTES
        blob = None
        if self.rank == self.dst:
            self.ptr = int(self._ck(rt.cudaMalloc(nbytes), "cudaMalloc"))
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
            self.ptr, self.cap = 0, 0
            raise PeerGatherUnavailable("cudaIpcOpenMemHandle failed on at least one rank")
        self.cap = cap_records

    def start(self, records: torch.Tensor) -> "PendingGather":
        """Enqueue the collection of `records` ([n, 64] uint8, a private copy with global channel
        numbers).  Returns at once; wait() gives the concatenation on dst (a view of the receive buffer,
        valid until the next start()), None elsewhere."""
        rt = self._rt
        n_local = torch.tensor([records.shape[0]], dtype=torch.int64, device=self.device)
        counts = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(counts, n_local, group=self.group)
        counts = counts.cpu().tolist()
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        total = offs[-1]
        if total > self.cap:
            self._allocate(total + total // 4 + 1024)
        records = records.contiguous()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            if counts[self.rank] > 0:
                self._ck(rt.cudaMemcpyAsync(self.ptr + offs[self.rank] * REC_BYTES, records.data_ptr(),
                                            counts[self.rank] * REC_BYTES, rt.cudaMemcpyKind.cudaMemcpyDefault,
                                            self.copy_stream.cuda_stream), "cudaMemcpyAsync (peer)")
            # every rank enters this after its own copy has completed on its stream, so once it is
            # done on dst all slices are in place
            work = dist.all_reduce(self._flag, group=self.group, async_op=True)
        out = None
        if self.rank == self.dst:
            out = (torch.as_tensor(_CudaView(self.ptr, total * REC_BYTES), device=self.device).view(total, REC_BYTES)
                   if total else torch.empty((0, REC_BYTES), dtype=torch.uint8, device=self.device))
        return PendingGather([work], out, records)

    def close(self):
        self._release()

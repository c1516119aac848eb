"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo.  Channels shard
embarrassingly; the only exchange is the collection of 64-byte message records on rank 0
(gnuais_b200/dist.py).  The same functions run over NCCL in bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnuais_b200 import MSG_DTYPE
from gnuais_b200 import dist as gdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_records(rank: int, n_local_channels: int, seed: int) -> np.ndarray:
    """(channel, end_bit)-sorted records a rank's BatchReceiver would hold; rank 1 may have none"""
    rng = np.random.default_rng(seed + rank)
    n = int(rng.integers(0, 40)) if rank else 25
    m = np.zeros(n, dtype=MSG_DTYPE)
    m["channel"] = np.sort(rng.integers(0, n_local_channels, n))
    m["end_bit"] = np.arange(n) * 300 + rank
    m["nbits"] = 168
    m["payload"] = rng.integers(0, 256, (n, 53), dtype=np.uint8)
    return m


def _worker(rank, world, port, total_channels, seed, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = gdist.shard_channels(total_channels, world, rank)
    local = _fake_records(rank, count, seed)
    recs = torch.from_numpy(local.view(np.uint8).reshape(-1, 64).copy())
    recs = gdist.globalize_channels(recs, first)
    got = gdist.gather_records(recs, dst=0)
    got2 = gdist.gather_records_async(recs.clone(), dst=0).wait()      # the overlapped form bench.py uses
    assert (got is None and got2 is None) or torch.equal(got, got2)
    tot = gdist.reduce_totals((len(local), rank, 1))
    if rank == 0:
        np.save(out_path, got.numpy())
        assert tot[2] == world
    else:
        assert got is None
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total_channels", [(2, 101), (2, 64)])
def test_gather_records_gloo(tmp_path, world, total_channels):
    port, seed = _free_port(), 77
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, port, total_channels, seed, out), nprocs=world, join=True)
    got = np.load(out).view(MSG_DTYPE).reshape(-1)
    want = []
    for r in range(world):
        first, count = gdist.shard_channels(total_channels, world, r)
        m = _fake_records(r, count, seed)
        m["channel"] += first
        want.append(m)
    want = np.concatenate(want)
    assert got.tobytes() == want.tobytes()
    key = got["channel"].astype(np.int64) << 32 | got["end_bit"]
    # ranks own ascending channel ranges -> rank-order concatenation is the canonical global order
    assert np.all(np.diff(got["channel"].astype(np.int64)) >= 0) and len(np.unique(key)) == len(key)


def test_shard_channels_partition():
    for total in (1, 7, 64, 65536, 524288 + 3):
        for world in (1, 2, 4, 8):
            spans = [gdist.shard_channels(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1

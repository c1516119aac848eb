"""The NVLink peer-copy gather (gnuais_b200/dist.py PeerGather) against the NCCL point-to-point gather on
the same records: two ranks on two GPUs of one box.  Skipped on boxes with fewer than two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnuais_b200 import MSG_DTYPE
from gnuais_b200 import dist as gdist

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _records(rank: int, n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed + rank)
    m = np.zeros(n, dtype=MSG_DTYPE)
    m["channel"] = np.sort(rng.integers(0, 4096, n))
    m["end_bit"] = np.arange(n) * 300 + rank
    m["nbits"] = 168
    m["payload"] = rng.integers(0, 256, (n, 53), dtype=np.uint8)
    return m


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pg = gdist.PeerGather(dst=0)
    # three rounds: growing counts force a re-allocation of the receive buffer; one round with an empty rank
    for rnd, n in enumerate([(1000, 700), (0, 5000), (300000, 250000)]):
        local = _records(rank, n[rank], 100 * rnd)
        recs = torch.from_numpy(local.view(np.uint8).reshape(-1, 64).copy()).to(dev)
        recs = gdist.globalize_channels(recs, rank * 4096)
        want = gdist.gather_records(recs, dst=0)
        got = pg.start(recs.clone()).wait()
        torch.cuda.synchronize()
        if rank == 0:
            assert got.shape == want.shape and torch.equal(got, want), f"round {rnd}: peer gather differs from NCCL gather"
        else:
            assert got is None and want is None
    pg.close()
    dist.destroy_process_group()


def test_peer_gather_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)

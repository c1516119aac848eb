"""The NVLink peer-copy gather (gnuais_b200/dist.py PeerGather: fixed slices in rank 0's buffer, no count exchange,
no collective on the compute stream) against the NCCL point-to-point gather on the same records, and -- end to end --
two sharded BatchReceivers whose gathered records must be the oracle's, with global channel numbers.  Two ranks on
two GPUs of one box; skipped on boxes with fewer than two GPUs."""
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
    m["channel"] = np.sort(rng.integers(0, 4096, n)) + rank * 4096
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
    pg = gdist.PeerGather(cap_records=300000, dst=0)
    # three rounds back to back without waiting in between (the slices are rewritten every step); one with an empty rank
    for rnd, n in enumerate([(1000, 700), (0, 5000), (300000, 250000)]):
        local = _records(rank, n[rank], 100 * rnd)
        recs = torch.from_numpy(local.view(np.uint8).reshape(-1, 64).copy()).to(dev)
        want = gdist.gather_records(recs, dst=0)
        pg.wait_source_free()
        pg.start(recs)
        got = pg.finish()
        if rank == 0:
            counts, views = got
            assert counts == list(n)
            assert torch.equal(torch.cat(views), want), f"round {rnd}: peer gather differs from NCCL gather"
        else:
            assert got is None and want is None
    with pytest.raises(RuntimeError):
        pg.start(torch.zeros((300001, 64), dtype=torch.uint8, device=dev))
    pg.close()

    # end to end: 2 x 64 channels sharded over the two ranks, records gathered on rank 0 == oracle, global channel numbers
    import oracle_lib as O
    from gnuais_b200 import BatchReceiver, SynthParams, nmea_format, synth_device, synth_host
    n_ch, frames = 64, 96000
    p = SynthParams(seed=909, sigma=300.0, rho=0.7)
    d = torch.empty((n_ch, frames), dtype=torch.int16, device=dev)
    synth_device(p, d, n_ch, frames, first_channel=rank * n_ch)
    rx = BatchReceiver(n_ch, frames, device=rank, first_channel=rank * n_ch)
    pg = gdist.PeerGather(cap_records=n_ch * (frames // 1280 + 2), dst=0)
    for _ in range(2):
        rx.reset()
        rx.run(d)
        rx.sync()
        pg.wait_source_free()
        pg.start(gdist.device_records(rx))
    got = pg.finish()
    tot = gdist.reduce_totals(rx.totals(), device=dev)
    if rank == 0:
        counts, views = got
        allrec = torch.cat(views).cpu().numpy().view(MSG_DTYPE).reshape(-1)
        assert len(allrec) == tot[0] == sum(counts)
        key = allrec["channel"].astype(np.int64) << 32 | allrec["end_bit"]
        assert np.all(np.diff(key) > 0)                       # canonical global order
        host = synth_host(p, 2 * n_ch, frames)
        for c in (0, 17, 63, 64, 100, 127):
            want = O.port().run(host[c])
            text = b"".join(nmea_format(m) for m in allrec[allrec["channel"] == c])
            assert text == want.nmea and want.ok > 10, c
    pg.close()
    rx.close()
    dist.destroy_process_group()


def test_peer_gather_matches_nccl_and_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    mp.spawn(_worker, args=(2, _free_port()), nprocs=2, join=True)

"""GPU parity tests of the ONE-kernel chain (gnuais_b200/csrc/gais_fused.cuh), forced on at sizes the oracle finishes
in seconds (by default the library keeps small batches on the two-kernel path): every output the library reports
-- FIR sign words, NRZI bits, message records, NMEA bytes, counters, DPLL / FSM state -- against the oracle port
(reference: src/receiver.c:87-148, src/protodec.c:988-1122) and against the two-kernel path.  Shapes are chosen to
hit the seams: channels beyond a multiple of 32 and samples beyond a multiple of 256 (swept up by the FIR-sign and
tracking kernels), several runs on one context (carried history, DPLL phase, half-received frames), host buffers
(one launch per staging tile), more channel sets than one wave of CTAs can own."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_lib as O
from gnuais_b200 import BatchReceiver, SynthParams, synth_device, synth_host
from test_parity_gpu import check_channel, check_records, nmea_of_channel, torch_dev

pytestmark = pytest.mark.gpu


def run_chain(planar, chain, chunks=None, host=False, keep=True, tile_frames=0):
    torch = torch_dev()
    C_, N = planar.shape
    chunks = chunks or [N]
    rx = BatchReceiver(C_, max(chunks), keep_bits=keep, keep_signs=keep, chain=chain, tile_frames=tile_frames)
    msgs, nmea, bits, signs, launches = [], [], [[] for _ in range(C_)], [], []
    off = 0
    for n in chunks:
        part = np.ascontiguousarray(planar[:, off:off + n])
        rx.run(part if host else torch.from_numpy(part).cuda())
        msgs.append(rx.messages())
        nmea.append(rx.nmea_records())
        launches.append(rx.timing()["launches"])
        if keep:
            for c, b in enumerate(rx.bits()):
                bits[c].append(b)
            signs.append(rx.signs(n))
        off += n
    out = dict(msgs=msgs, nmea_recs=nmea, counters=rx.counters(), state=rx.state(), totals=rx.totals(), launches=launches)
    if keep:
        out["bits"] = [np.concatenate(b) for b in bits]
        out["signs"] = np.concatenate(signs, axis=1)
    rx.close()
    return out


def oracle_all(planar):
    with ThreadPoolExecutor(max_workers=16) as ex:
        return list(ex.map(lambda row: O.port().run(row), planar))


@pytest.mark.parametrize("n_ch,n_frames", [(64, 48000), (45, 30000), (96, 25610), (33, 1000)])
def test_fused_chain_equals_oracle(n_ch, n_frames):
    """whole sets and whole stages (64 x 48000), leftover channels (45 = 32 + 13), a ragged end (25610 = 100 stages + 10
    samples), and both at once on a run shorter than a tile"""
    x = synth_host(SynthParams(seed=77 + n_ch, sigma=300.0, rho=0.7), n_ch, n_frames)
    res = run_chain(x, "fused")
    for c, w in enumerate(oracle_all(x)):
        check_channel(res, c, w)
    assert sum(int(v["ok"]) for v in res["counters"]) > 0


def test_fused_chain_edge_signals():
    """rows that stress the FIR's open-output queue and the DPLL: silence (every output exactly 0: each one goes to the
    resolver), +-2 alternation, full-scale noise, DC -- next to ordinary rows, 64 channels x 20480 samples"""
    rng = np.random.default_rng(5)
    n_ch, n = 64, 20480
    x = synth_host(SynthParams(seed=9, sigma=1500.0, rho=0.9), n_ch, n)
    x[3] = 0
    x[17] = np.where(np.arange(n) % 2, 2, -2)
    x[18] = rng.integers(-32768, 32768, n, dtype=np.int64).astype(np.int16)
    x[40] = 1000
    x[41, ::2] = 32767
    x[41, 1::2] = -32768
    res = run_chain(x, "fused")
    for c, w in enumerate(oracle_all(x)):
        check_channel(res, c, w)


def test_fused_chain_streams_and_host_buffers():
    """several runs on one context (ragged chunk sizes: state, history and open frames cross launches) and the host-buffer
    path (a launch per staging tile) give what one device run gives"""
    n_ch, n = 64, 60000
    x = synth_host(SynthParams(seed=31, sigma=300.0, rho=0.8), n_ch, n)
    whole = run_chain(x, "fused")
    parts = run_chain(x, "fused", chunks=[256, 20000, 1, 12345, 27398])
    host = run_chain(x, "fused", host=True, tile_frames=8192)
    want = oracle_all(x)
    for res in (whole, parts, host):
        for c in (0, 1, 31, 32, 63):
            check_channel(res, c, want[c])
        assert [tuple(int(v[k]) for k in ("ok", "crcfail", "sizefail")) for v in res["counters"]] == [w.counters() for w in want]
    assert whole["launches"][0] < host["launches"][0]        # one chain launch per run vs one per staging tile


def test_fused_chain_equals_two_kernel_chain_many_sets():
    """71088 channels = 2221 sets + 16 leftover channels: more sets than one CTA per SM can own (14 x 148 = 2072), so the
    grid does not fit one wave; records, counters and state equal the two-kernel chain's, a sample of channels equals
    the oracle"""
    torch = torch_dev()
    n_ch, n = 2221 * 32 + 16, 6400
    p = SynthParams(seed=404, sigma=300.0, rho=0.6)
    d = torch.empty((n_ch, n), dtype=torch.int16, device="cuda")
    synth_device(p, d, n_ch, n)
    out = {}
    for chain in ("fused", "two_kernel"):
        with BatchReceiver(n_ch, n, chain=chain) as rx:
            rx.run(d)
            out[chain] = (rx.messages(), rx.counters(), rx.state(), rx.timing()["launches"])
    a, b = out["fused"], out["two_kernel"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert len(a[0]) > 1000
    host = d.cpu().numpy()
    for c in (0, 31, 32, 479, 480, 9999, 40000, 71039, 71040, 71071, 71072, 71087):
        w = O.port().run(host[c], want_bits=False)
        cnt, st = a[1][c], a[2][c]
        assert (int(cnt["ok"]), int(cnt["crcfail"]), int(cnt["sizefail"])) == w.counters()
        assert (int(st["pll"]), int(st["fsm_state"]), int(st["seqnr"])) == (w.pll, w.fsm_state, w.seqnr)


def test_default_chain_choice():
    """small batches stay on the two kernels, batches that fill the GPU take the fused kernel (one chain launch per run)"""
    torch = torch_dev()
    n_sms = torch.cuda.get_device_properties(0).multi_processor_count
    n = 4096
    for n_ch, fused in ((1024, False), (5 * 32 * n_sms, True)):
        d = torch.empty((n_ch, n), dtype=torch.int16, device="cuda")
        synth_device(SynthParams(seed=1, sigma=300.0, rho=0.5), d, n_ch, n)
        with BatchReceiver(n_ch, n) as rx, BatchReceiver(n_ch, n, chain="fused") as rf, BatchReceiver(n_ch, n, chain="two_kernel") as r2:
            for r in (rx, rf, r2):
                r.run(d)
            assert rx.timing()["launches"] == (rf if fused else r2).timing()["launches"]
            assert rf.timing()["launches"] < r2.timing()["launches"]


@pytest.mark.parametrize("chain", ["fused", "two_kernel"])
def test_interleaved_batches_take_the_fast_kernels(chain):
    """the reference's own buffer shape (frame-interleaved, src/receiver.c:102) with many channels: 70 channels x 33000
    samples, device and host buffers, several tiles; the tile is re-laid as planar rows on the device and goes through the
    same kernels as planar input -- everything equal to the oracle, and no slower path's launch count (the exact FIR
    kernel would add launches per tile for every channel)"""
    torch = torch_dev()
    n_ch, n = 70, 33000
    planar = synth_host(SynthParams(seed=88, sigma=300.0, rho=0.7), n_ch, n)
    inter = np.ascontiguousarray(planar.T)                      # [n, n_ch]
    want = oracle_all(planar)
    for host in (False, True):
        with BatchReceiver(n_ch, n, layout="interleaved", keep_bits=True, keep_signs=True, chain=chain, tile_frames=8192) as rx:
            rx.run(inter if host else torch.from_numpy(inter).cuda())
            res = dict(msgs=[rx.messages()], nmea_recs=[rx.nmea_records()], counters=rx.counters(), state=rx.state(),
                       bits=rx.bits(), signs=rx.signs(n))
        for c, w in enumerate(want):
            check_channel(res, c, w)


def test_fused_chain_runs_longer_than_one_launch():
    """a run of 2^22 + 5696 samples: the fused kernel takes at most 2^22 samples per launch (sample indices travel in 23 bits), so
    the run is cut in two launches with history, DPLL phase and open frames carried across the cut; 35 channels (one whole set +
    3 left over).  Equal to the two-kernel chain everywhere and to the oracle on a sample of channels."""
    torch = torch_dev()
    n_ch, n = 35, (1 << 22) + 5696
    p = SynthParams(seed=512, sigma=300.0, rho=0.6)
    d = torch.empty((n_ch, n), dtype=torch.int16, device="cuda")
    synth_device(p, d, n_ch, n)
    out = {}
    for chain in ("fused", "two_kernel"):
        with BatchReceiver(n_ch, n, chain=chain) as rx:
            rx.run(d)
            out[chain] = (rx.messages(), rx.counters(), rx.state(), rx.nmea_records())
    a, b = out["fused"], out["two_kernel"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert len(a[0]) > 10000
    res = dict(msgs=[a[0]], nmea_recs=[a[3]], counters=a[1], state=a[2])
    for c in (0, 31, 33):
        w = O.port().run(d[c].cpu().numpy(), want_bits=False, want_signs=False)
        check_channel(res, c, w, bits=False, signs=False)

#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 AIS receive path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU receiver

A step is one pass of the hot path (receiver_run() for every channel of the batch, through
message compaction) over one batch of synthetic GMSK int16 audio that is already resident in
HBM.  Workload at N GPUs: 65536 channels x 480000 samples (10 s at 48 kHz) PER GPU -- BASELINE
config "65536 batched channels ... single B200" at N=1 and "524288 channels sharded across
8xB200" at N=8 (weak scaling; channels are independent, no data-path collective; the only
exchange is the collection of decoded-message buffers on rank 0 -- NVLink peer copies into rank 0's
buffer, counts and completion over NCCL -- inside the timed step).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "Msamples/s demodulated (FIR + DPLL + NRZI + HDLC + CRC-16, AIS msgs/s alongside)"
UNIT = "Msamples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=65536, help="channels per GPU")
    ap.add_argument("--frames", type=int, default=480000, help="samples per channel per step")
    ap.add_argument("--sigma", type=float, default=300.0)
    ap.add_argument("--rho", type=float, default=0.5)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--fir-mode", default="guard", choices=["guard", "exact"])
    ap.add_argument("--tile-frames", type=int, default=0)
    ap.add_argument("--e2e-channels", type=int, default=8192, help="channels per GPU of the host-buffer leg (7.9 GB pinned at 480000 samples)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the cfg2 / cfg3 side measurements")
    ap.add_argument("--no-gather-check", action="store_true")
    ap.add_argument("--check-channels", type=int, default=256, help="channels per GPU re-run through the oracle after the timed region")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-two-kernel", action="store_true", help="skip the side measurement of the two-kernel path (GAIS_FUSED=0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-channels-per-core", type=int, default=24)
    return ap.parse_args()


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(kernel: str, alg_bytes: float):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r1_traffic.json), scaled to this launch's algorithmic bytes; None if no capture."""
    p = ROOT / "profiles" / "r2_traffic.json"
    if not p.exists():
        p = ROOT / "profiles" / "r1_traffic.json"
    try:
        rec = json.loads(p.read_text())[kernel]
        return rec["dram_bytes_per_launch"] * alg_bytes / rec["alg_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark(self):
        """timed region starts now: remember how many samples were taken before it"""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)          # let the last in-region sample land
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = out.splitlines()
        # keep the samples of the timed region: the last ceil(region / 100 ms) + 1 lines
        if getattr(self, "t_mark", None) is not None:
            keep = int((time.time() - self.t_mark) / 0.1) + 1
            lines = lines[-keep:] if keep < len(lines) else lines
        for line in lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(host_planar, cores: int):
    """the reference's CPU receiver (oracle/_ref) -- or the port when it was never built -- on a
    bounded sample of the same workload, one struct receiver per channel, all host threads"""
    import oracle_lib as O

    kind = "reference" if (ROOT / "oracle" / "_ref" / "libgnuais_ref.so").exists() else "port"
    chk = O.ref(tap=False, quiet=False) if kind == "reference" else O.port()
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)          # the reference printf()s one line per message (src/protodec.c:934)
    try:
        secs, ok = chk.bench(host_planar, cores)
    finally:
        os.dup2(saved, 1)
        os.close(saved); os.close(devnull)
    n = host_planar.shape[0] * host_planar.shape[1]
    return {"value": n / secs / 1e6, "unit": UNIT, "cores": cores, "kind": kind, "msgs_per_s": ok / secs,
            "sample": f"{host_planar.shape[0]} channels x {host_planar.shape[1]} samples of the same workload "
                      f"({n / 1e6:.1f} Msamples, {secs:.2f} s wall), receiver_run() in 1020-frame chunks"}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from gnuais_b200 import SynthParams, synth_host

    cores = os.cpu_count() or 1
    n_ch = max(cores, min(cores * args.cpu_channels_per_core, 4096))
    frames = args.frames
    # bounded sample of OUR arm's workload: its first n_ch channels (same seed -> same audio)
    t0 = time.time()
    x = synth_host(SynthParams(seed=args.seed, sigma=args.sigma, rho=args.rho), n_ch, frames)
    gen_s = time.time() - t0
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(x, cores)
        if i >= args.warmup:
            vals.append(last)
    n = n_ch * frames
    secs = [n / (v["value"] * 1e6) for v in vals]
    value = n * len(secs) / sum(secs) / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32/u32", "data": "synthetic",
        "config": {"workload": f"bounded sample: first {n_ch} of the {args.channels} channels/GPU x {frames} samples "
                               f"(48 kHz int16 synthetic GMSK, seed {args.seed}, sigma {args.sigma}, rho {args.rho}) per step",
                   "host_gen_s": round(gen_s, 2)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": last["kind"], "sample": last["sample"]},
        "msgs_per_s": sum(v["msgs_per_s"] for v in vals) / len(vals),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def bind_to_gpu_numa_node(local_rank: int):
    """pin this process (and, by first touch, the page-locked buffers it allocates afterwards) to the CPUs next to its
    GPU: /sys/bus/pci/devices/<bus id>/local_cpulist.  Best effort; returns a note for the bench line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = Path("/sys/bus/pci/devices") / bus
        cpus = set()
        for part in (base / "local_cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no local cpulist inside the process' affinity mask"
        os.sched_setaffinity(0, cpus)
        node = (base / "numa_node").read_text().strip()
        return f"numa node {node}, {len(cpus)} cpus"
    except Exception as e:          # containers without /sys access, one-node hosts ...
        return f"not bound ({type(e).__name__})"


def other_configs(args, dev, local_rank):
    """BASELINE configs 2 and 3, measured outside the timed region (N = 1): they have one lane per channel walking the whole
    time axis in the tracker, so they are latency-bound by construction -- reported, not optimised for."""
    import torch
    from gnuais_b200 import BatchReceiver, SynthParams, synth_device
    out = {}
    for name, n_ch, frames, layout in (("cfg2: 2 channels (AIS1+AIS2) x 2880000 samples, frame-interleaved stereo", 2, 2880000, "interleaved"),
                                       ("cfg3: 1024 channels x 480000 samples, planar", 1024, 480000, "planar")):
        d = torch.empty((n_ch, frames) if layout == "planar" else (frames, n_ch), dtype=torch.int16, device=dev)
        synth_device(SynthParams(seed=args.seed, sigma=args.sigma, rho=args.rho), d, n_ch, frames, layout=layout)
        rx = BatchReceiver(n_ch, frames, layout=layout, device=local_rank, fir_mode=args.fir_mode)
        times = []
        for i in range(4):
            rx.run(d)
            rx.sync()
            if i:
                times.append(rx.timing())
        ms = sum(t["total_ms"] for t in times) / len(times)
        out[name] = {"ms_per_run": ms, "Msamples_per_s": n_ch * frames / ms / 1e3, "fir_ms": sum(t["fir_ms"] for t in times) / len(times),
                     "track_ms": sum(t["track_ms"] for t in times) / len(times), "msgs": rx.message_count(),
                     "note": "latency-bound: one tracker lane per channel walks the whole time axis"}
        rx.close()
        del d
    return out


def gather_parity(args, rx, d, p, first_channel, n_ch, frames, rank, world, peer, dev):
    """once, outside the timed region: a seeded subset of every rank's channels is regenerated on the host and run through
    the oracle; the rank checks its own counters / DPLL+FSM state, rank 0 checks the records it GATHERED (global channel
    numbers, canonical order, NMEA bytes) for every rank's subset, and the frame counters are summed over the ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor
    import oracle_lib as O
    from gnuais_b200 import MSG_DTYPE, nmea_format, synth_host
    from gnuais_b200 import dist as gdist

    k = min(args.check_channels, n_ch)
    rng = np.random.default_rng(1234 + rank)
    subset = np.sort(rng.choice(n_ch, size=k, replace=False))
    rx.reset()
    rx.run(d, stream=torch.cuda.current_stream(dev).cuda_stream)
    rx.sync()
    cnt, st = rx.counters(), rx.state()

    def one(c):
        x = synth_host(p, 1, frames, first_channel=first_channel + int(c))[0]
        return O.port().run(x, want_bits=False)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        want = list(ex.map(one, subset))
    bad = []
    for c, w in zip(subset, want):
        got = (int(cnt[c]["ok"]), int(cnt[c]["crcfail"]), int(cnt[c]["sizefail"]), int(st[c]["pll"]), int(st[c]["fsm_state"]), int(st[c]["seqnr"]))
        if got != (w.ok, w.crcfail, w.sizefail, w.pll, w.fsm_state, w.seqnr):
            bad.append(f"rank {rank} channel {first_channel + int(c)}: counters/state {got} != oracle")
    expect = {first_channel + int(c): w.nmea for c, w in zip(subset, want)}
    totals = rx.totals()
    if world > 1:
        peer.wait_source_free()
        peer.start(gdist.device_records(rx))
        res = peer.finish()
        boxes = [None] * world
        dist.gather_object((expect, bad), boxes if rank == 0 else None, dst=0)
        tot = gdist.reduce_totals(totals, device=dev)
    else:
        res = ([rx.message_count()], [gdist.device_records(rx)])
        boxes = [(expect, bad)]
        tot = totals
    if rank != 0:
        return None
    counts, views = res
    n_checked, n_msgs = 0, 0
    for r in range(world):
        exp, b = boxes[r]
        bad += b
        rec = views[r].cpu().numpy().view(MSG_DTYPE).reshape(-1)
        key = rec["channel"].astype(np.int64) << 32 | rec["end_bit"]
        if len(rec) and not np.all(np.diff(key) > 0):
            bad.append(f"slice {r}: records not in (channel, end_bit) order")
        if len(rec) and not (rec["channel"].min() >= r * n_ch and rec["channel"].max() < (r + 1) * n_ch):
            bad.append(f"slice {r}: channel numbers outside the rank's range")
        for c, text in exp.items():
            sel = rec[rec["channel"] == c]
            n_msgs += len(sel)
            if b"".join(nmea_format(m) for m in sel) != text:
                bad.append(f"slice {r}: NMEA of channel {c} differs from the oracle")
            n_checked += 1
    if sum(counts) != tot[0]:
        bad.append(f"gathered {sum(counts)} records, ranks counted {tot[0]} CRC-ok frames")
    if bad:
        return "FAILED: " + "; ".join(bad[:5])
    return (f"ok ({n_checked} channels of {world} rank(s) bit-exact vs oracle: counters, DPLL/FSM state, {n_msgs} gathered records -> NMEA; "
            f"{sum(counts)} records gathered = sum of the ranks' ok counters)")


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from gnuais_b200 import BatchReceiver, SynthParams, synth_device
    from gnuais_b200 import dist as gdist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gnuais_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_ch, frames = args.channels, args.frames
    note = ""
    try:
        d = torch.empty((n_ch, frames), dtype=torch.int16, device=dev)
    except torch.cuda.OutOfMemoryError:
        frames = 96000
        note = " (480000-sample rows did not fit next to the result buffers: fell back to 96000 samples per step)"
        d = torch.empty((n_ch, frames), dtype=torch.int16, device=dev)
    p = SynthParams(seed=args.seed, sigma=args.sigma, rho=args.rho)
    first_channel = rank * n_ch
    synth_device(p, d, n_ch, frames, first_channel=first_channel)
    torch.cuda.synchronize()

    rx = BatchReceiver(n_ch, frames, device=local_rank, fir_mode=args.fir_mode, tile_frames=args.tile_frames,
                       first_channel=first_channel)
    stream = torch.cuda.current_stream(dev)
    acc = {"fir_ms": 0.0, "track_ms": 0.0, "post_ms": 0.0, "total_ms": 0.0, "nmea_ms": 0.0, "nmea_bytes": 0, "launches": 0, "tiles": 0,
           "msgs": 0, "gather_bytes": 0}

    pending = {"gather": None}
    # how the decoded records reach rank 0: by default each rank copies its slice straight into rank 0's
    # buffer over NVLink (CUDA IPC peer mapping, copy engines); GAIS_GATHER=nccl uses NCCL send/recv instead
    gx = {"peer": None, "kind": "none (one rank)"}
    if world > 1:
        if os.environ.get("GAIS_GATHER", "peer") == "peer":
            try:
                # a slice per rank in rank 0's buffer, sized for one message per AIS slot per channel
                gx["peer"] = gdist.PeerGather(cap_records=n_ch * (frames // 1280 + 2), dst=0)
                gx["kind"] = ("NVLink peer copy of each rank's records into its fixed slice of rank 0's buffer (CUDA IPC, copy engines); "
                              "count in the slice header; nothing on the compute stream waits for another rank")
            except gdist.PeerGatherUnavailable as e:          # raised on every rank together
                gx["kind"] = f"NCCL point-to-point send/recv (peer mapping unavailable: {e})"
        else:
            gx["kind"] = "NCCL point-to-point send/recv of exact-size record arrays, counts over NCCL"

    def step(timed: bool):
        if gx["peer"]:
            gx["peer"].wait_source_free(stream)       # the previous step's records have left the dense array (local event)
        rx.run(d, stream=stream.cuda_stream)
        rx.sync()
        rx.device_nmea()                              # "!AIVDM" text of every message, packed, on the run's stream
        if world > 1:
            # collect the decoded messages on rank 0: records already carry global channel numbers
            recs = gdist.device_records(rx)
            if gx["peer"]:
                gx["peer"].start(recs)
                if timed and rank == 0:
                    acc["gather_bytes"] += int(recs.numel()) * world      # every rank moves about as much (same workload statistics)
            else:
                if pending["gather"] is not None:
                    out = pending["gather"].wait()
                    if timed and out is not None:
                        acc["gather_bytes"] += int(out.numel())
                pending["gather"] = gdist.gather_records_async(recs.clone(), dst=0)
        if timed:
            tm = rx.timing()
            for k in ("fir_ms", "track_ms", "post_ms", "total_ms", "nmea_ms"):
                acc[k] += tm[k]
            acc["launches"] += tm["launches"]
            acc["msgs"] += rx.message_count()
            acc["nmea_bytes"] += rx.device_nmea()[3]

    def drain(timed: bool):
        if gx["peer"]:
            gx["peer"].finish()                       # every rank's last slice is in place on rank 0
        if pending["gather"] is not None:
            out = pending["gather"].wait()
            pending["gather"] = None
            if timed and out is not None:
                acc["gather_bytes"] += int(out.numel())

    # nvidia-smi needs a moment to start: launch it before the warm-up steps so that it is sampling
    # (every 100 ms) by the time the timed region begins; samples are kept only from the timed region
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 0)):
        step(False)
    drain(False)
    if sampler:
        sampler.mark()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(True)
    drain(True)
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    parity = None
    if not args.no_gather_check and (world == 1 or gx["peer"]):
        parity = gather_parity(args, rx, d, p, first_channel, n_ch, frames, rank, world, gx["peer"], dev)
    if gx["peer"]:
        gx["peer"].close()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        m = torch.tensor([acc["msgs"]], dtype=torch.int64, device=dev)
        dist.all_reduce(m, op=dist.ReduceOp.SUM)
        total_msgs = int(m.item())
    else:
        total_msgs = acc["msgs"]

    fused = fused_path(n_ch, args.fir_mode)
    tile_frames = FUSED_MAX_FRAMES if fused else rx_tile_frames(n_ch, args.tile_frames)
    overlap = not fused and os.environ.get("GAIS_OVERLAP", "1") != "0" and frames > tile_frames
    plan = rx_tile_plan(frames, tile_frames, overlap)       # samples per launch of the chain, in order
    n_tiles = len(plan)
    samples_per_step = n_ch * frames * world
    value = samples_per_step * args.steps / (ms * 1e-3) / 1e6
    totals = rx.totals()

    # ---- roofline of the dominant kernel (CUDA events inside the library, on the run's stream) ----
    # the launches of a step are not all the same size (the last tile is shorter): achieved = the
    # algorithmic bytes of ALL launches of the timed region over the sum of their durations
    peak, peak_src = measured_peak()
    fir_step, trk_step = acc["fir_ms"] / args.steps, acc["track_ms"] / args.steps
    msgs_step = acc["msgs"] / args.steps
    if fused:
        # ONE kernel does FIR sign + DPLL + NRZI + HDLC (gais_fused.cuh); the library's "fir" events bracket it, the "track" events
        # bracket the sweep-up kernels of a ragged tile end (none at this shape)
        dom, step_ms_dom = "ais_fused", fir_step
        # algorithmic bytes: 2 B per input sample read + 64 B per frame candidate written
        alg_step = 2.0 * n_ch * frames + 64.0 * msgs_step
    else:
        dom = "fir_sign" if fir_step >= trk_step else "track"
        step_ms_dom = max(fir_step, trk_step)
        # 2 B per input sample for the kernel that reads the audio; 64 B per emitted record belongs to the tracking kernel
        alg_step = 2.0 * n_ch * frames if dom == "fir_sign" else 2.0 * n_ch * frames + 64.0 * msgs_step
    alg_bytes = alg_step / n_tiles
    launch_ms = step_ms_dom / n_tiles
    achieved = alg_step / (step_ms_dom * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(dom, alg_bytes), "peak_source": peak_src, "avg_launch_ms": launch_ms,
                "alg_bytes_per_launch": alg_bytes, "launches_per_step": n_tiles, "tile_plan_frames": plan_summary(plan),
                "timing": "CUDA events on the stream each kernel is launched on, inside the timed region"
                          + ("; FIR of tile t+1 runs CONCURRENTLY with tracking of tile t, so both launch durations "
                             "include the other kernel's share of the SMs (see solo)" if overlap else ""),
                "chain": {"fir_ms_per_step": acc["fir_ms"] / args.steps, "track_ms_per_step": acc["track_ms"] / args.steps,
                          "post_ms_per_step": acc["post_ms"] / args.steps,
                          "whole_chain_GBps": (2.0 * n_ch * frames + 64.0 * msgs_step)
                          / (acc["total_ms"] / args.steps * 1e-3) / 1e9}}
    if fused:
        roofline["chain"] = {"fused_kernel_ms_per_step": fir_step, "ragged_end_kernels_ms_per_step": trk_step,
                             "frame_check_scan_gather_ms_per_step": acc["post_ms"] / args.steps,
                             "whole_chain_GBps": alg_step / (acc["total_ms"] / args.steps * 1e-3) / 1e9,
                             "note": "one launch per step: FIR sign (tcgen05 kind::i8), DPLL, slicer, NRZI and the HDLC bit machine in one "
                                     "persistent kernel, sign words handed over in shared memory; bound by integer instruction issue "
                                     "(ALU pipe), not by HBM: see DESIGN.md"}
        if world == 1 and not args.no_two_kernel:
            # the same workload through the two-kernel path (GAIS_FUSED=0: FIR-sign kernel -> sign words in HBM -> tracking
            # kernel), outside the timed region: 1 warm-up + 2 steps
            os.environ["GAIS_FUSED"] = "0"
            try:
                rx2 = BatchReceiver(n_ch, frames, device=local_rank, fir_mode=args.fir_mode, tile_frames=args.tile_frames)
                two = {"total_ms": 0.0, "fir_ms": 0.0, "track_ms": 0.0}
                for i in range(3):
                    rx2.run(d, stream=stream.cuda_stream)
                    rx2.sync()
                    if i:
                        tm = rx2.timing()
                        for k2 in two:
                            two[k2] += tm[k2] / 2
                two["msgs"] = rx2.message_count()
                rx2.close()
                roofline["two_kernel_path"] = {"ms_per_step": two["total_ms"], "fir_ms_per_step": two["fir_ms"],
                                               "track_ms_per_step": two["track_ms"], "msgs": two["msgs"],
                                               "note": "GAIS_FUSED=0, FIR of tile t+1 overlapped with tracking of tile t"}
            finally:
                del os.environ["GAIS_FUSED"]
    if overlap:
        # the same kernels timed alone (overlap off), outside the timed region: 1 warm-up + 2 steps
        rxs = BatchReceiver(n_ch, frames, device=local_rank, fir_mode=args.fir_mode, tile_frames=args.tile_frames, overlap=False)
        solo = {"fir_ms": 0.0, "track_ms": 0.0}
        for i in range(3):
            rxs.run(d, stream=stream.cuda_stream)
            rxs.sync()
            if i:
                tm = rxs.timing()
                solo["fir_ms"] += tm["fir_ms"]; solo["track_ms"] += tm["track_ms"]
        rxs.close()
        n_plain = (frames + tile_frames - 1) // tile_frames      # overlap off: plain tiling
        f_ms, t_ms = solo["fir_ms"] / (2 * n_plain), solo["track_ms"] / (2 * n_plain)
        solo_gbps = 2.0 * n_ch * frames / (solo["fir_ms"] / 2 * 1e-3) / 1e9
        roofline["solo"] = {"fir_launch_ms": f_ms, "track_launch_ms": t_ms, "launches_per_step": n_plain,
                            "fir_GBps": solo_gbps, "fir_frac": solo_gbps / peak}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "s8 x u8 -> s32 (FIR on tcgen05 kind::i8), u32 (DPLL / HDLC / CRC)", "data": "synthetic",
        "config": {"workload": f"{n_ch} batched channels/GPU x {frames} samples (48 kHz int16, {frames / 48000:.0f} s), planar, "
                               f"synthetic GMSK seed {args.seed} sigma {args.sigma} rho {args.rho}{note}",
                   "channels_per_gpu": n_ch, "frames_per_channel": frames, "fir_mode": args.fir_mode,
                   "tile_frames": min(tile_frames, frames), "chain": "fused (one kernel)" if fused else "two kernels", "parallelism": f"channels sharded x{world}, no data-path collective",
                   "l2": f"input {2 * n_ch * frames / 1e9:.1f} GB per GPU >> 126 MB L2: no flush needed between steps"},
        "msgs_per_s": total_msgs / (ms * 1e-3), "msgs_per_step": total_msgs / args.steps,
        "counters_rank0": {"ok": totals[0], "crcfail": totals[1], "sizefail": totals[2]},
        "roofline": roofline, "gpu_launches": acc["launches"], "clocks": clocks,
        "nmea": {"ms_per_step": acc["nmea_ms"] / args.steps, "bytes_per_step": acc["nmea_bytes"] / args.steps,
                 "note": "packed !AIVDM text of every message of the step, armoured on the GPU inside the timed step "
                         "(3 launches: lengths per block, scan of block totals, one thread per message through shared memory)"},
        "gather_parity": parity,
    }
    if world > 1:
        line["gather_bytes_per_step_rank0"] = acc["gather_bytes"] / args.steps
        line["config"]["gather"] = gx["kind"]

    # ---- end to end through the C-ABI with HOST buffers (H2D + D2H inside the timed region) ----
    if not args.no_e2e:
        e_ch = min(args.e2e_channels, n_ch)
        numa = bind_to_gpu_numa_node(local_rank)        # before the page-locked allocation: first touch decides the node
        host = torch.empty((e_ch, frames), dtype=torch.int16, pin_memory=True)
        host.copy_(d[:e_ch])
        torch.cuda.synchronize()
        rx.close()
        del d
        torch.cuda.empty_cache()
        rxe = BatchReceiver(e_ch, frames, device=local_rank, fir_mode=args.fir_mode, tile_frames=args.tile_frames)
        hv = host.numpy()
        d2h = 0
        for i in range(2):
            rxe.run(hv); rxe.messages(reuse=True)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e_steps = max(3, min(args.steps, 5))
        for _ in range(e_steps):
            rxe.run(hv)
            d2h += rxe.messages(reuse=True).nbytes      # D2H of the records into the receiver's page-locked buffer
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        line["e2e"] = {"value": e_ch * frames * world * e_steps / dt / 1e6, "unit": UNIT,
                       "h2d_bytes_per_step": 2 * e_ch * frames, "d2h_bytes_per_step": d2h // e_steps,
                       "workload": f"{e_ch} channels/GPU x {frames} samples ({2 * e_ch * frames / 1e9:.1f} GB page-locked per rank) from "
                                   f"pinned host memory through gais_run_host(), message records copied back to page-locked host "
                                   f"memory every step; bound by the host link (PCIe H2D), not by the kernels",
                       "steps": e_steps, "numa": numa}
        cpu_src = hv
        rxe.close()
    else:
        cpu_src = None

    if rank == 0 and world == 1 and not args.no_other_configs:
        line["other_configs"] = other_configs(args, dev, local_rank)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = max(cores, min(cores * args.cpu_channels_per_core, n_ch))
        if cpu_src is not None and cpu_src.shape[0] >= n_cpu:
            sample = np.ascontiguousarray(cpu_src[:n_cpu])
        else:
            from gnuais_b200 import synth_host
            sample = synth_host(p, n_cpu, frames, first_channel=first_channel)
        line["cpu_baseline"] = cpu_baseline(sample, cores)

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def rx_tile_plan(frames: int, tile: int, overlap: bool):
    """mirror of the library's time-tile plan (gais_api.cu gais_run_device): plain tiling.  (A plan with a
    short first and last tile, meant to fill and drain the two-kernel pipeline faster, measured 0.6 %
    slower and was dropped: profiles/r1_experiments.txt.)"""
    return [min(tile, frames - f0) for f0 in range(0, frames, tile)]


def plan_summary(plan):
    out, i = [], 0
    while i < len(plan):
        j = i
        while j < len(plan) and plan[j] == plan[i]:
            j += 1
        out.append(f"{j - i}x{plan[i]}" if j - i > 1 else str(plan[i]))
        i = j
    return " + ".join(out)


FUSED_MAX_FRAMES = 1 << 22       # gais_fused.cuh X_MAX_FRAMES: samples per launch of the fused kernel


def fused_path(n_ch: int, fir_mode: str) -> bool:
    """mirror of the library's choice (gais_api.cu fused_part): the one-kernel chain takes planar, aligned input whose
    channel count is a multiple of 32 and gives every SM at least five 32-channel sets, in guard mode, unless GAIS_FUSED=0
    or an older FIR is selected for an A/B run"""
    mode = os.environ.get("GAIS_FUSED", "1")
    n_sms = 148
    try:
        import torch
        n_sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    except Exception:
        pass
    return (fir_mode == "guard" and n_ch % 32 == 0 and (mode == "2" or (mode == "1" and n_ch // 32 >= 5 * n_sms))
            and os.environ.get("GAIS_FIR_IMPL", "") in ("", "tc"))


def rx_tile_frames(n_ch: int, requested: int) -> int:
    """mirror of the library's default time-tile choice (gais_api.cu gais_create)"""
    tile = requested or int(os.environ.get("GAIS_TILE_FRAMES", "0") or 0)
    if tile <= 0:
        tile = 384 * 1024 * 1024 * 8 // n_ch
        tile = max(2048, min(65536, tile))
    return (tile + 1023) // 1024 * 1024


if __name__ == "__main__":
    sys.exit(main())

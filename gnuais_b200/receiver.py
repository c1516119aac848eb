"""Host-side mirror of the gnuais receiver interface over the B200 C-ABI.

``BatchReceiver`` is the batched form of ``struct receiver`` (src/receiver.h:35-51 of the
reference): ``n_channels`` independent receivers advanced together by ``run()`` (=
``receiver_run()`` for every channel, src/receiver.c:87-135).  ``init_receiver`` /
``receiver_run`` / ``free_receiver`` below keep the reference's names, argument meaning and
error behaviour for a single channel so that tests read like calls into the reference.

PyTorch / numpy are used for buffers only; all arithmetic happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L

MSG_DTYPE = np.dtype(
    [("payload", "u1", 53), ("flags", "u1"), ("nbits", "<u2"), ("channel", "<u4"), ("end_bit", "<u4")]
)
COUNTERS_DTYPE = np.dtype([("ok", "<i4"), ("crcfail", "<i4"), ("sizefail", "<i4")])
STATE_DTYPE = np.dtype(
    [("pll", "<u4"), ("prev", "<i4"), ("lastbit", "<i4"), ("fsm_state", "<i4"), ("seqnr", "<i4"), ("n_bits", "<u4")]
)
NMEA_DTYPE = np.dtype([("len", "u1"), ("text", "S175")])
assert MSG_DTYPE.itemsize == 64 and NMEA_DTYPE.itemsize == L.NMEA_STRIDE


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class BatchReceiver:
    """n_channels gnuais receivers on one B200 (see include/gais_b200.h for the C contract)."""

    def __init__(self, n_channels: int, max_frames_per_run: int, layout: str = "planar", device: int = 0,
                 fir_mode: str = "guard", keep_bits: bool = False, keep_signs: bool = False,
                 slot_cap: int = 0, tile_frames: int = 0, overlap=None, keep_peak: bool = False, first_channel: int = 0,
                 chain=None):
        self._lib = L.load()
        self._ctx = C.c_void_p()
        cfg = L.Config()
        cfg.abi_version = L.ABI_VERSION
        cfg.device = device
        cfg.n_channels = n_channels
        cfg.layout = {"planar": L.LAYOUT_PLANAR, "interleaved": L.LAYOUT_INTERLEAVED}[layout]
        cfg.max_frames_per_run = max_frames_per_run
        cfg.fir_mode = {"guard": L.FIR_GUARD, "exact": L.FIR_EXACT}[fir_mode]
        cfg.flags = (L.KEEP_BITS if keep_bits else 0) | (L.KEEP_SIGNS if keep_signs else 0) | (L.KEEP_PEAK if keep_peak else 0)
        cfg.reserved[0] = slot_cap
        cfg.reserved[1] = tile_frames
        cfg.reserved[2] = 0 if overlap is None else (2 if overlap else 1)
        cfg.reserved[3] = first_channel
        # kernel chain: None = library default (fused kernel for batches that fill the GPU), "fused" / "two_kernel" force one
        cfg.reserved[4] = {None: 0, "two_kernel": 1, "fused": 2}[chain]
        L.check(self._lib.gais_create(C.byref(cfg), C.byref(self._ctx)))
        self.n_channels = n_channels
        self.first_channel = first_channel
        self.layout = layout
        self.max_frames_per_run = max_frames_per_run
        self._keepalive = None
        self._pin_ptr, self._pin_cap = C.c_void_p(), 0

    # -- lifecycle -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_pin_ptr", None) is not None and self._pin_ptr.value:
            self._lib.gais_host_free(self._pin_ptr)
            self._pin_ptr, self._pin_cap = C.c_void_p(), 0
        if getattr(self, "_ctx", None):
            self._lib.gais_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self) -> None:
        L.check(self._lib.gais_reset(self._ctx))

    # -- the hot path ----------------------------------------------------------------------
    def run(self, samples, n_frames: Optional[int] = None, stride: Optional[int] = None, stream=None) -> None:
        """Advance every channel by one chunk.

        ``samples``: int16, a CUDA torch tensor (device path, asynchronous on ``stream`` /
        the current torch stream) or a numpy array / CPU torch tensor (host path: H2D copies
        are pipelined inside the library).  planar: shape [n_channels, n_frames];
        interleaved: shape [n_frames, num_ch] with ``num_ch >= n_channels``.
        """
        planar = self.layout == "planar"
        if _is_torch(samples) and samples.is_cuda:
            import torch

            assert samples.dtype == torch.int16 and samples.dim() == 2
            if n_frames is None:
                n_frames = samples.shape[1] if planar else samples.shape[0]
            if stride is None:
                stride = samples.stride(0)
                assert samples.stride(1) == 1 or samples.shape[1] == 1
                if samples.shape[0] == 1:      # the stride of a length-1 axis is arbitrary
                    stride = samples.shape[1]
            if stream is None:
                stream = torch.cuda.current_stream(samples.device).cuda_stream
            self._keepalive = samples
            L.check(self._lib.gais_run_device(self._ctx, C.c_void_p(samples.data_ptr()), n_frames, stride,
                                              C.c_void_p(stream)))
            return
        if _is_torch(samples):
            samples = samples.numpy()
        a = np.asarray(samples)
        assert a.dtype == np.int16 and a.ndim == 2 and (a.strides[1] == 2 or a.shape[1] == 1)
        if n_frames is None:
            n_frames = a.shape[1] if planar else a.shape[0]
        if stride is None:
            stride = a.shape[1] if a.shape[0] == 1 else a.strides[0] // 2
        self._keepalive = a
        L.check(self._lib.gais_run_host(self._ctx, a.ctypes.data_as(C.c_void_p), n_frames, stride))

    def run_bits(self, bits, stream=None) -> None:
        """``protodec_decode()`` for every channel (src/protodec.c:988-1122): ``bits`` is uint8 0/1, shape
        [n_channels, n_bits] -- the NRZI-decoded bits receiver_run() would hand over -- as a numpy array / CPU
        tensor (host path) or a CUDA torch tensor (device path)."""
        if _is_torch(bits) and bits.is_cuda:
            import torch

            assert bits.dtype == torch.uint8 and bits.dim() == 2 and bits.shape[0] == self.n_channels
            if stream is None:
                stream = torch.cuda.current_stream(bits.device).cuda_stream
            self._keepalive = bits
            stride = bits.shape[1] if bits.shape[0] == 1 else bits.stride(0)
            L.check(self._lib.gais_run_bits_device(self._ctx, C.c_void_p(bits.data_ptr()), bits.shape[1], stride, C.c_void_p(stream)))
            return
        if _is_torch(bits):
            bits = bits.numpy()
        a = np.ascontiguousarray(np.asarray(bits, dtype=np.uint8))
        assert a.ndim == 2 and a.shape[0] == self.n_channels
        self._keepalive = a
        L.check(self._lib.gais_run_bits_host(self._ctx, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[1]))

    def sync(self) -> None:
        L.check(self._lib.gais_sync(self._ctx))

    # -- results ---------------------------------------------------------------------------
    def message_count(self) -> int:
        n = C.c_int64()
        L.check(self._lib.gais_message_count(self._ctx, C.byref(n)))
        return n.value

    def messages(self, reuse: bool = False) -> np.ndarray:
        """Records of the last run.  ``reuse=True`` returns a view of a page-locked buffer owned by the
        receiver (valid until the next call): the device-to-host copy then runs at PCIe speed, without the
        allocation, page faults and staging pass of a fresh pageable array."""
        n = self.message_count()
        if reuse:
            if self._pin_cap < n:
                if self._pin_ptr.value:
                    self._lib.gais_host_free(self._pin_ptr)
                    self._pin_ptr, self._pin_cap = C.c_void_p(), 0
                cap = n + n // 8 + 1024
                L.check(self._lib.gais_host_alloc(C.byref(self._pin_ptr), cap * MSG_DTYPE.itemsize))
                self._pin_cap = cap
            buf = (C.c_uint8 * (self._pin_cap * MSG_DTYPE.itemsize)).from_address(self._pin_ptr.value)
            out = np.frombuffer(buf, dtype=MSG_DTYPE, count=n)
        else:
            out = np.empty(n, dtype=MSG_DTYPE)
        got = C.c_int64()
        L.check(self._lib.gais_get_messages(self._ctx, out.ctypes.data_as(C.c_void_p), n, C.byref(got)))
        return out

    def device_messages(self):
        """(device pointer, count) of the dense message array of the last run."""
        p, n = C.c_void_p(), C.c_int64()
        L.check(self._lib.gais_device_messages(self._ctx, C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def nmea_records(self) -> np.ndarray:
        n = self.message_count()
        out = np.zeros(n, dtype=NMEA_DTYPE)
        got = C.c_int64()
        L.check(self._lib.gais_get_nmea(self._ctx, out.ctypes.data_as(C.c_void_p), n, C.byref(got)))
        return out

    def nmea(self) -> bytes:
        """Concatenated ``!AIVDM...\\r\\n`` sentences of the last run, (channel, end_bit) order (packed on the GPU)."""
        n = C.c_int64()
        L.check(self._lib.gais_get_nmea_text(self._ctx, None, 0, C.byref(n)))
        buf = np.empty(max(n.value, 1), dtype=np.uint8)
        L.check(self._lib.gais_get_nmea_text(self._ctx, buf.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return buf[: n.value].tobytes()

    def device_nmea(self):
        """(text pointer, offsets pointer, n_msgs, n_bytes) of the packed NMEA text on the device."""
        t, o, n, nb = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        L.check(self._lib.gais_device_nmea(self._ctx, C.byref(t), C.byref(o), C.byref(n), C.byref(nb)))
        return t.value or 0, o.value or 0, n.value, nb.value

    def counters(self) -> np.ndarray:
        out = np.zeros(self.n_channels, dtype=COUNTERS_DTYPE)
        L.check(self._lib.gais_get_counters(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def state(self) -> np.ndarray:
        out = np.zeros(self.n_channels, dtype=STATE_DTYPE)
        L.check(self._lib.gais_get_state(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def totals(self):
        t = (C.c_int64 * 3)()
        L.check(self._lib.gais_get_totals(self._ctx, C.byref(t)))
        return tuple(int(x) for x in t)

    def bits(self):
        """Per-channel NRZI bit arrays (uint8 0/1) of the last run (needs keep_bits)."""
        rw = C.c_int64()
        L.check(self._lib.gais_bits_row_words(self._ctx, C.byref(rw)))
        words = np.zeros((self.n_channels, rw.value), dtype=np.uint32)
        nbits = np.zeros(self.n_channels, dtype=np.uint32)
        L.check(self._lib.gais_get_bits(self._ctx, words.ctypes.data_as(C.c_void_p), nbits.ctypes.data_as(C.c_void_p)))
        out = []
        for c in range(self.n_channels):
            b = np.unpackbits(words[c].view(np.uint8), bitorder="little")[: nbits[c]]
            out.append(b)
        return out

    def signs(self, n_frames: int) -> np.ndarray:
        """[n_channels, n_frames] uint8 (filtered > 0) of the last run (needs keep_signs)."""
        nw = (n_frames + 31) // 32
        words = np.zeros((nw, self.n_channels), dtype=np.uint32)
        L.check(self._lib.gais_get_signs(self._ctx, words.ctypes.data_as(C.c_void_p), words.size))
        per_ch = np.ascontiguousarray(words.T)
        return np.unpackbits(per_ch.view(np.uint8), axis=1, bitorder="little")[:, :n_frames]

    def peaks(self) -> np.ndarray:
        """Per-channel level of the last run: what filter_run_buf() returns (max of the positive samples,
        src/filter.c:112-119); needs keep_peak."""
        out = np.zeros(self.n_channels, dtype=np.int16)
        L.check(self._lib.gais_get_peaks(self._ctx, out.ctypes.data_as(C.c_void_p)))
        return out

    def timing(self) -> dict:
        t = L.Timing()
        L.check(self._lib.gais_get_timing(self._ctx, C.byref(t)), allow=(L.E_OVERFLOW,))
        return {k: getattr(t, k) for k in ("total_ms", "fir_ms", "track_ms", "post_ms", "launches", "nmea_ms")}


def nmea_format(msg) -> bytes:
    """Host-side armouring of one MSG_DTYPE record (same bytes as the GPU kernel)."""
    lib = L.load()
    m = L.Msg.from_buffer_copy(np.asarray(msg).tobytes())
    buf = C.create_string_buffer(L.NMEA_STRIDE)
    n = lib.gais_nmea_format(C.byref(m), buf)
    return buf.raw[:n]


def text_format(msg, chanid: str = "A") -> bytes:
    """The stdout line gnuais prints for one MSG_DTYPE record (src/protodec.c:931-985)."""
    lib = L.load()
    m = L.Msg.from_buffer_copy(np.asarray(msg).tobytes())
    buf = C.create_string_buffer(1024)
    n = lib.gais_text_format(C.byref(m), chanid.encode()[:1], buf, 1024)
    return buf.raw[:max(n, 0)]


# ---- single-channel mirror of the reference's receiver.h ------------------------------------

class _Decoder:
    """The fields of struct demod_state_t a caller reaches into (src/ais.c:296-310)."""

    def __init__(self, chanid: str):
        self.chanid = chanid
        self.receivedframes = 0
        self.lostframes = 0
        self.lostframes2 = 0
        self.seqnr = 0


class Receiver:
    def __init__(self, name, num_ch, ch_ofs, serial, ipc, batch_frames):
        self.name, self.num_ch, self.ch_ofs = name, num_ch, ch_ofs
        self.serial, self.ipc = serial, ipc
        self.decoder = _Decoder(name)
        self.batch_frames = batch_frames
        self._pending = []
        self._pending_frames = 0
        self._rx = BatchReceiver(1, batch_frames + 4096, layout="planar")

    def _flush(self):
        if not self._pending_frames:
            return
        row = np.concatenate(self._pending)[None, :]
        self._pending, self._pending_frames = [], 0
        self._rx.run(np.ascontiguousarray(row))
        text = self._rx.nmea()
        if self.serial is not None and text:
            self.serial.write(text)                      # "!%s\r\n" per sentence, src/protodec.c:883-885
        if self.ipc is not None and text:
            for line in text.split(b"\r\n"):
                if line:
                    self.ipc.write(line)                 # "!%s", src/protodec.c:886-888
        cnt = self._rx.counters()[0]
        st = self._rx.state()[0]
        d = self.decoder
        d.receivedframes, d.lostframes, d.lostframes2 = int(cnt["ok"]), int(cnt["crcfail"]), int(cnt["sizefail"])
        d.seqnr = int(st["seqnr"])


def init_receiver(name: str, num_ch: int, ch_ofs: int, serial=None, ipc=None, batch_frames: int = 48000) -> Receiver:
    """src/receiver.c:52-74.  ``serial``/``ipc``: objects with ``write(bytes)`` or None."""
    return Receiver(name, num_ch, ch_ofs, serial, ipc, batch_frames)


def receiver_run(rx: Receiver, buf, length: int) -> None:
    """src/receiver.c:87-135.  ``buf``: frame-interleaved int16 (``num_ch`` per frame); this
    receiver reads column ``ch_ofs``.  ``length`` > 4096 aborts in the reference
    (src/receiver.c:104-105); here it raises.  The chunk is copied before returning (the
    caller reuses ``buf``, src/ais.c:216-226) and decoded when ``batch_frames`` have been
    collected or at ``free_receiver``; message order is unchanged."""
    if length > 4096:
        raise RuntimeError("receiver_run: len > FILTERED_LEN (4096): the reference abort()s here")
    a = np.asarray(buf, dtype=np.int16).reshape(-1, rx.num_ch)[:length, rx.ch_ofs]
    rx._pending.append(a.copy())
    rx._pending_frames += length
    if rx._pending_frames >= rx.batch_frames:
        rx._flush()


def free_receiver(rx: Optional[Receiver]) -> None:
    """src/receiver.c:76-82 (NULL-tolerant).  Flushes what is still batched."""
    if rx is None:
        return
    rx._flush()
    rx._rx.close()

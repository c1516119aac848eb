/*
 * gais_track.cuh -- K2+K3, the per-channel sequential stage: zero-crossing DPLL, slicer,
 * NRZI decode (src/receiver.c:109-135) and the HDLC bit FSM (src/protodec.c:988-1122), plus the
 * frame check that follows it (CRC-16, counters, seqnr: src/protodec.c:106-167, :896-929).
 *
 * One lane = one channel, sequential in time, reading the [word][channel] sign words the FIR
 * stage wrote (coalesced: a warp reads 32 consecutive channels of one word row).
 *
 * DPLL, event driven.  The reference does, per sample,
 *     if (cur != prev) pll += (pll < 0x8000) ? +819 : -819;   pll += 13107;
 *     if (pll > 0xffff) { slice; pll &= 0xffff; }
 * Between two sign changes this is a pure translation, so the kernel jumps from crossing to
 * crossing.  The phase lives in a 64-bit register Z = (slices_so_far << 32) | (pll << 16):
 * advancing n samples is ONE 32x32+64 multiply-add  Z += n * (13107 << 16); the carries into the
 * upper word ARE the slices (the reference's "> 0xffff" / "&= 0xffff"), so Z's upper word is the
 * running NRZI bit count.  The +-819 nudge touches the lower word only (it can neither carry nor
 * borrow: +819 is applied below 0x8000, -819 at or above it).
 *
 * NRZI without looking at slice positions.  The decoded bit of a slice is 1 iff the sliced
 * sign equals the previously sliced sign, i.e. iff an EVEN number of sign changes happened
 * between the two slices.  Every crossing therefore toggles bit number <slices so far> of a
 * "difference" accumulator (the bit that belongs to the next slice to come); a slice simply
 * moves on to the next bit, which starts at 0.  NRZI bits = ~difference bits.
 *
 * HDLC.  Difference bits are handed to the bit FSM in chunks of 24..31.  While hunting for a
 * preamble (the common state on noise) a chunk is cleared with a few bit tricks -- the FSM can
 * leave ST_SKURR only after more than 14 alternations ending in a 0 (src/protodec.c:1029-1037),
 * which a run-length test on the chunk rules out exactly; otherwise the chunk goes through the
 * per-bit FSM.  A closed frame is NOT checked here: the stored bits go to the channel's slot
 * list as a 64-byte candidate and crc_kernel / finalize_kernel (massively parallel, no
 * divergence) do CRC, counters, seqnr and the in-place compaction into gais_msg records.  The
 * reference's FSM never looks at the CRC verdict (it resets either way, src/protodec.c:1113),
 * so deferring it changes nothing observable.
 */
#ifndef GAIS_TRACK_CUH
#define GAIS_TRACK_CUH

#include "gais_kernels.cuh"

namespace gais {

#define GAIS_INC64 (GAIS_PLL_INC << 16)      /* 0x33330000: one sample of phase, in Z units */
#define GAIS_NUDGE64 (GAIS_PLL_NUDGE << 16)

struct HdlcRegs {
	uint32_t fsm, stuffed, last, nflag, nones, nalt, pos, cur;
};

__device__ __forceinline__ void hdlc_reset(HdlcRegs &f)       /* src/protodec.c:87-100 */
{
	f.fsm = GAIS_ST_HUNT;
	f.nflag = 0; f.nalt = 0; f.nones = 0; f.last = 0; f.stuffed = 0; f.pos = 0; f.cur = 0;
}

/* candidate layout (64 B, same slot a gais_msg will occupy): words 0..13 stored bits (LSB first),
 * word 14 = bufferpos | stop_bit << 16 (| status << 24 after crc_kernel), word 15 = closing bit index */
__device__ __noinline__ void hdlc_emit(const HdlcRegs &f, uint32_t b, uint32_t bit_index, const ChanState *s, int c,
				       uint32_t &ncand, const TrackOut &out)
{
	if (ncand >= (uint32_t) out.slot_cap) {
		*out.overflow = 1;
		return;
	}
	uint32_t *w = reinterpret_cast<uint32_t *>(&out.slots[(int64_t) c * out.slot_cap + ncand]);
	const uint32_t nw = f.pos >> 5;
#pragma unroll
	for (uint32_t i = 0; i < 14; i++)
		w[i] = (i < nw) ? s->store[i] : (i == nw ? f.cur : 0u);
	w[14] = f.pos | (b << 16);
	w[15] = bit_index;
	ncand++;
}

/* one NRZI bit through the FSM -- src/protodec.c:988-1122, frame check deferred */
__device__ __forceinline__ void hdlc_bit(HdlcRegs &f, uint32_t b, uint32_t bit_index, ChanState *s, int c, uint32_t &ncand,
					 const TrackOut &out)
{
	switch (f.fsm) {
	case GAIS_ST_DATA:
		if (f.stuffed) {
			if (b) f.fsm = GAIS_ST_STOPFLAG;
			f.stuffed = 0;
		} else {
			if (b == f.last && b == 1u) {
				if (++f.nones == 4u) { f.stuffed = 1; f.nones = 0; }
			} else {
				f.nones = 0;
			}
			f.cur |= b << (f.pos & 31u);
			f.pos++;
			if ((f.pos & 31u) == 0u) {
				s->store[(f.pos >> 5) - 1u] = f.cur;
				f.cur = 0;
			}
			if (f.pos >= 449u)
				hdlc_reset(f);
		}
		break;
	case GAIS_ST_HUNT:
		f.nalt = (b != f.last) ? f.nalt + 1u : 0u;
		if (f.nalt > 14u && b == 0u) { f.fsm = GAIS_ST_PREAMBLE; f.nalt = 0; }
		break;
	case GAIS_ST_PREAMBLE:
		if (b != f.last && f.nflag == 0u) {
			/* antallpreamble++ : never read before it is zeroed again */
		} else if (b == 1u) {
			if (f.nflag == 0u) f.nflag = 3;
			else if (f.nflag == 5u) { f.nflag = 6; f.nalt = 0; f.fsm = GAIS_ST_STARTFLAG; }
			else f.nflag++;
		} else {
			if (f.nflag == 0u) f.nflag = 1;
			else hdlc_reset(f);
		}
		break;
	case GAIS_ST_STARTFLAG:
		if (f.nflag >= 7u) {
			if (b == 0u) { f.fsm = GAIS_ST_DATA; f.nflag = 0; f.nones = 0; f.pos = 0; f.cur = 0; }
			else hdlc_reset(f);
		} else if (b == 0u) {
			hdlc_reset(f);
		}
		f.nflag++;                                   /* src/protodec.c:1092: even after a reset */
		break;
	default: /* GAIS_ST_STOPFLAG: src/protodec.c:1095-1115 */
		hdlc_emit(f, b, bit_index, s, c, ncand, out);
		hdlc_reset(f);
		break;
	}
	f.last = b;                                          /* src/protodec.c:1119 */
}

/* n (1..31) NRZI bits, bit 0 of W the oldest; hb = index of that bit in the channel's stream */
__device__ __forceinline__ void hdlc_chunk(HdlcRegs &f, uint32_t W, uint32_t n, uint32_t hb, ChanState *s, int c,
					   uint32_t &ncand, const TrackOut &out)
{
	if (f.fsm == GAIS_ST_HUNT) {
		const uint32_t vm = (1u << n) - 1u;
		const uint32_t A = (W ^ ((W << 1) | f.last)) & vm;      /* bit i: b_i != b_(i-1) */
		const uint32_t L = (uint32_t) __ffs((int) ~A) - 1u;          /* leading alternations (<= n) */
		uint32_t r = A & (A >> 1);
		r &= r >> 2;
		r &= r >> 4;
		r &= r >> 7;                                                 /* a run of >= 15 alternations inside */
		if (f.nalt + L <= 14u && r == 0u) {
			f.nalt = (L == n) ? f.nalt + n : (uint32_t) __clz((int) ~(A << (32u - n)));
			f.last = (W >> (n - 1u)) & 1u;
			return;
		}
	}
	for (uint32_t i = 0; i < n; i++)
		hdlc_bit(f, (W >> i) & 1u, hb + i, s, c, ncand, out);
}

/* OR n bits of W into the run-relative bit record (GAIS_KEEP_BITS); off may be negative for
 * bits that were sliced in the previous run */
__device__ __forceinline__ void bits_or(const TrackOut &out, int c, int64_t off, uint32_t W, uint32_t n)
{
	if (off < 0) {
		if ((int64_t) n <= -off)
			return;
		W >>= (uint32_t) (-off);
		n -= (uint32_t) (-off);
		off = 0;
	}
	W &= (n >= 32u) ? 0xffffffffu : ((1u << n) - 1u);
	uint32_t *row = out.bits + (int64_t) c * out.bits_row_words;
	const uint32_t sh = (uint32_t) off & 31u;
	row[off >> 5] |= W << sh;
	if (sh && (W >> (32u - sh)))
		row[(off >> 5) + 1] |= W >> (32u - sh);
}

__global__ void __launch_bounds__(128)
track_kernel(const uint32_t *__restrict__ signs, ChanState *st, int n_channels, int64_t n_frames, TrackOut out)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_channels)
		return;
	ChanState *s = &st[c];

	uint32_t zlo = s->pll << 16, zhi = s->n_bits;
	uint32_t prevword = s->prev;                 /* bit 0 = sign of the last sample seen */
	uint32_t dlo = s->dacc, nd = s->nd;          /* difference bits not yet given to the FSM */
	uint32_t hb = zhi - nd;                      /* stream index of dlo bit 0 */
	HdlcRegs f;
	f.fsm = s->fsm; f.stuffed = s->stuffed; f.last = s->last; f.nflag = s->nflag; f.nones = s->nones;
	f.nalt = s->nalt; f.pos = s->pos; f.cur = s->cur;
	uint32_t ncand = out.run_count[c];
	const uint32_t zhi_start = zhi;
	const int64_t run_start = (int64_t) zhi - (int64_t) out.run_bits[c];   /* stream index of the run's first bit */

	const int64_t n_words = (n_frames + 31) >> 5;
	for (int64_t w = 0; w < n_words; w++) {
		const uint32_t sw = signs[w * n_channels + c];
		const int64_t left = n_frames - w * 32;
		const uint32_t nb = left < 32 ? (uint32_t) left : 32u;
		/* MSB-first words: bit 31 is the first sample.  x marks samples whose sign differs from
		 * the sample before */
		uint32_t x = sw ^ __funnelshift_r(sw, prevword, 1);
		if (nb < 32u) {
			x &= 0xffffffffu << (32u - nb);
			prevword = sw >> (32u - nb);
		} else {
			prevword = sw;
		}
		uint32_t jp = 0;
		while (x) {
			const uint32_t j = (uint32_t) __clz((int) x);
			x &= ~(0x80000000u >> j);
			/* samples jp .. j-1: no sign change (src/receiver.c:121-134 only) */
			unsigned long long Z = ((unsigned long long) zhi << 32) | zlo;
			Z += (unsigned long long) (j - jp) * GAIS_INC64;
			zlo = (uint32_t) Z; zhi = (uint32_t) (Z >> 32);
			jp = j;
			/* sample j changes sign: the pending NRZI difference bit flips, the phase is nudged
			 * towards the crossing (src/receiver.c:113-119) before this sample's own increment */
			dlo ^= 1u << (zhi - hb);
			zlo += ((int32_t) zlo < 0) ? (0u - GAIS_NUDGE64) : GAIS_NUDGE64;
		}
		{
			unsigned long long Z = ((unsigned long long) zhi << 32) | zlo;
			Z += (unsigned long long) (nb - jp) * GAIS_INC64;
			zlo = (uint32_t) Z; zhi = (uint32_t) (Z >> 32);
		}
		nd = zhi - hb;
		if (nd >= 24u) {
			/* at most 7 slices per 32 samples, so bit 31 is never reached before this flush */
			const uint32_t W = ~dlo;
			if (out.bits)
				bits_or(out, c, (int64_t) hb - run_start, W, nd);
			hdlc_chunk(f, W, nd, hb, s, c, ncand, out);
			dlo >>= nd;
			hb += nd;
			nd = 0;
		}
	}
	if (nd) {
		/* end of the tile: hand the sliced bits over now, so that FSM state, candidates and
		 * counters at a run boundary are exactly the reference's after the same samples */
		const uint32_t W = ~dlo;
		if (out.bits)
			bits_or(out, c, (int64_t) hb - run_start, W, nd);
		hdlc_chunk(f, W, nd, hb, s, c, ncand, out);
		dlo >>= nd;
		hb += nd;
		nd = 0;
	}

	s->pll = zlo >> 16; s->n_bits = zhi; s->prev = (uint8_t) (prevword & 1u);
	s->dacc = dlo; s->nd = (uint8_t) nd;
	s->lastbit = (uint8_t) ((prevword ^ (dlo >> nd)) & 1u);   /* sign at the last slice */
	s->fsm = (uint8_t) f.fsm; s->stuffed = (uint8_t) f.stuffed; s->last = (uint8_t) f.last;
	s->nflag = (uint8_t) f.nflag; s->nones = (uint8_t) f.nones;
	s->nalt = (uint16_t) (f.nalt > 0xffffu ? 0xffffu : f.nalt); s->pos = (uint16_t) f.pos; s->cur = f.cur;
	out.run_count[c] = ncand;
	out.run_bits[c] += zhi - zhi_start;
}

/* ---- frame check, fully parallel: one thread per candidate ------------------------------- */
__global__ void __launch_bounds__(256)
crc_kernel(gais_msg *__restrict__ slots, const uint32_t *__restrict__ run_count, int slot_cap, int n_channels)
{
	__shared__ uint16_t table[256];
	{
		/* CRC-16/X.25 (reflected 0x8408) byte table; same recurrence as src/protodec.c:106-118 */
		uint32_t v = threadIdx.x;
#pragma unroll
		for (int k = 0; k < 8; k++)
			v = (v & 1u) ? (v >> 1) ^ 0x8408u : v >> 1;
		table[threadIdx.x] = (uint16_t) v;
	}
	__syncthreads();
	const int64_t idx = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const int c = (int) (idx / slot_cap), k = (int) (idx % slot_cap);
	if (c >= n_channels || (uint32_t) k >= run_count[c])
		return;
	uint32_t *w = reinterpret_cast<uint32_t *>(&slots[idx]);
	const uint32_t meta = w[14];
	const int pos = (int) (meta & 0xffffu), nbits = pos - 22;          /* src/protodec.c:1096 */
	const uint32_t stopbit = (meta >> 16) & 1u;
	uint32_t status = 2;                                               /* lostframes2 */
	if (stopbit == 0u && nbits > 0) {
		const int nbytes = (nbits >> 3) + 2;                           /* src/protodec.c:133-134 */
		uint32_t crc = 0xffffu;
		for (int j = 0; j < nbytes; j++) {
			const uint32_t byte = (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
			crc = (crc >> 8) ^ table[(crc ^ byte) & 0xffu];
		}
		status = ((~crc & 0xffffu) == 0x0f47u) ? 0u : 1u;              /* src/protodec.c:166 */
	}
	w[14] = meta | (status << 24);
}

/* one thread per channel: counters, type gate + seqnr (src/protodec.c:896-929), and in-place
 * compaction of the CRC-ok candidates into gais_msg records */
__global__ void __launch_bounds__(128)
finalize_kernel(gais_msg *__restrict__ slots, uint32_t *__restrict__ run_count, int slot_cap, ChanState *st, int n_channels)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_channels)
		return;
	const uint32_t ncand = run_count[c];
	ChanState *s = &st[c];
	int32_t ok = s->ok, crcfail = s->crcfail, sizefail = s->sizefail;
	uint32_t seqnr = s->seqnr, nout = 0;
	uint4 *row = reinterpret_cast<uint4 *>(slots + (int64_t) c * slot_cap);
	for (uint32_t k = 0; k < ncand; k++) {
		uint4 q[4];
#pragma unroll
		for (int i = 0; i < 4; i++)
			q[i] = row[4 * k + i];
		const uint32_t meta = q[3].z, status = (meta >> 24) & 3u;
		if (status == 1u) { crcfail++; continue; }
		if (status == 2u) { sizefail++; continue; }
		ok++;
		uint32_t w[16] = { q[0].x, q[0].y, q[0].z, q[0].w, q[1].x, q[1].y, q[1].z, q[1].w,
				   q[2].x, q[2].y, q[2].z, q[2].w, q[3].x, q[3].y, q[3].z, q[3].w };
		const int nbits = (int) (meta & 0xffffu) - 22, nb = nbits >> 3;
		const uint32_t type = (w[0] & 0xffu) >> 2;
		const uint32_t gate = (type >= 1u && type <= 24u) ? 1u : 0u;
		const uint32_t flags = seqnr | (gate << 4);
#pragma unroll
		for (int i = 0; i < 13; i++) {
			const int lo = 32 * i;
			if (8 * nb <= lo) w[i] = 0;
			else if (8 * nb < lo + 32) w[i] &= (1u << (8 * nb - lo)) - 1u;
		}
		w[13] = ((nb > 52) ? (w[13] & 0xffu) : 0u) | (flags << 8) | ((uint32_t) nbits << 16);
		w[14] = (uint32_t) c;
		/* w[15] already holds the closing bit index */
		if (gate)
			seqnr = (seqnr + 1u) % 10u;                           /* src/protodec.c:924-926 */
#pragma unroll
		for (int i = 0; i < 4; i++)
			row[4 * nout + i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
		nout++;
	}
	s->ok = ok; s->crcfail = crcfail; s->sizefail = sizefail; s->seqnr = (uint8_t) seqnr;
	run_count[c] = nout;
}

static inline int track_launch(const uint32_t *signs, ChanState *st, int n_ch, int64_t n_frames, const TrackOut &out,
			       cudaStream_t stream)
{
	track_kernel<<<(n_ch + 127) / 128, 128, 0, stream>>>(signs, st, n_ch, n_frames, out);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

/* after the last tile of a run */
static inline int finalize_launch(ChanState *st, int n_ch, const TrackOut &out, cudaStream_t stream)
{
	const int64_t total = (int64_t) n_ch * out.slot_cap;
	crc_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>>(out.slots, out.run_count, out.slot_cap, n_ch);
	finalize_kernel<<<(n_ch + 127) / 128, 128, 0, stream>>>(out.slots, out.run_count, out.slot_cap, st, n_ch);
	return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

} /* namespace gais */
#endif

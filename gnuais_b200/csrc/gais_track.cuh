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
 * HDLC, branch-free.  With one lane per channel the 32 lanes of a warp sit in different FSM
 * states, so a switch() would execute every state's code for every bit.  Instead the FSM is a
 * transition table in shared memory: 80 states (the reachable combinations of state / nstartsign
 * / antallpreamble / antallenner / bitstuff / last of src/protodec.h:44-71, antallpreamble
 * saturated at 15 because only "> 14" is ever tested) x input bit -> next state + action flags
 * (store the bit, enter DATA, frame closed).  All lanes do the same few instructions per bit.
 * A chunk on which every lane of the warp is hunting (idle channels) is skipped with a
 * run-length test instead.
 *
 * A closed frame is NOT checked here: the stored bits go to the channel's slot list as a
 * 64-byte candidate and frame_check_kernel (one warp per channel, one lane per candidate) does
 * CRC, counters, seqnr and the in-place compaction into gais_msg records.  The reference's FSM
 * never looks at the CRC verdict (it resets either way, src/protodec.c:1113), so deferring it
 * changes nothing observable.
 */
#ifndef GAIS_TRACK_CUH
#define GAIS_TRACK_CUH

#include "gais_kernels.cuh"

namespace gais {

#define GAIS_INC64 (GAIS_PLL_INC << 16)      /* 0x33330000: one sample of phase, in Z units */
#define GAIS_NUDGE64 (GAIS_PLL_NUDGE << 16)

/* ---- HDLC state numbering (ChanState.fsm holds these ids) --------------------------------
 *   0..63   ST_SKURR      id = leak*32 + nalt*2 + last   (nalt = antallpreamble, saturated at 15;
 *                         leak = the nstartsign==1 a failed ST_STARTSIGN leaves behind, :1092)
 *   64,65   ST_PREAMBLE   nstartsign == 0, still alternating; id = 64 + last
 *   66..70  ST_PREAMBLE   nstartsign == 1..5; id = 65 + nstartsign
 *   71,72   ST_STARTSIGN  nstartsign == 6, 7
 *   73      ST_DATA       last == 0
 *   74..77  ST_DATA       last == 1, antallenner == 0..3
 *   78      ST_DATA       bitstuff pending
 *   79      ST_STOPSIGN
 */
constexpr int H_NSTATES = 80;
constexpr uint32_t H_STORE = 0x80u, H_ENTER = 0x100u, H_EMIT = 0x200u;

__host__ __device__ inline uint32_t hdlc_hunt_id(uint32_t leak, uint32_t nalt, uint32_t last)
{
	return leak * 32u + (nalt > 15u ? 15u : nalt) * 2u + last;
}

/* transition of state `id` on bit b: next id | action flags.  Written straight from
 * src/protodec.c:988-1122 (the trailing "d->last = in[i]" is folded into the next id). */
__host__ __device__ inline uint32_t hdlc_transition(uint32_t id, uint32_t b)
{
	if (id < 64u) {                                         /* :1028-1041 */
		const uint32_t leak = id >> 5, nalt = (id >> 1) & 15u, last = id & 1u;
		const uint32_t n2 = (b != last) ? (nalt < 15u ? nalt + 1u : 15u) : 0u;
		if (n2 > 14u && b == 0u)
			return leak ? 66u : 64u;                        /* ST_PREAMBLE, nstartsign = leak, last = 0 */
		return hdlc_hunt_id(leak, n2, b);
	}
	if (id < 66u) {                                         /* :1043-1070, nstartsign == 0 */
		const uint32_t last = id - 64u;
		if (b != last)
			return 64u + b;
		return b ? 68u : 66u;                               /* nstartsign = 3 | 1 */
	}
	if (id < 71u) {                                         /* nstartsign = 1..5 */
		const uint32_t k = id - 65u;
		if (b)
			return k == 5u ? 71u : id + 1u;
		return hdlc_hunt_id(0, 0, 0);                       /* protodec_reset(); last = 0 */
	}
	if (id == 71u)                                          /* :1072-1093, nstartsign == 6 */
		return b ? 72u : hdlc_hunt_id(1, 0, 0);             /* reset, then nstartsign++ -> 1 */
	if (id == 72u)                                          /* nstartsign >= 7 */
		return b ? hdlc_hunt_id(1, 0, 1) : (73u | H_ENTER);
	if (id == 73u)                                          /* :993-1026 */
		return (b ? 74u : 73u) | H_STORE;
	if (id < 78u) {
		const uint32_t n = id - 74u;
		if (!b)
			return 73u | H_STORE;
		return (n == 3u ? 78u : id + 1u) | H_STORE;
	}
	if (id == 78u)
		return b ? 79u : 73u;                               /* sixth one -> ST_STOPSIGN | stuffed 0 dropped */
	return hdlc_hunt_id(0, 0, b) | H_EMIT;                  /* :1095-1115 */
}

/* ST_* value of src/protodec.h:30-34 for an id (gais_chan_state.fsm_state) */
__host__ __device__ inline int hdlc_public_state(uint32_t id)
{
	return id < 64u ? GAIS_ST_HUNT : id < 71u ? GAIS_ST_PREAMBLE : id < 73u ? GAIS_ST_STARTFLAG : id < 79u ? GAIS_ST_DATA : GAIS_ST_STOPFLAG;
}

struct HdlcRegs {
	uint32_t id;        /* state id */
	uint32_t pos;       /* bufferpos */
	uint32_t shi, slo;  /* the last 64 stored bits, newest at bit 31 of shi */
};

/* nibble table entry: what four input bits do to state `id` (bit 0 of v first).  The fields sit where the
 * loop gets each of them with ONE instruction:
 *   bits 0..1   position of the closing bit (EMIT only)
 *   bits 6..12  next id << 6 = byte offset of the next state's row (16 entries x 4 B): the next lookup
 *               address is (entry & 0x1fc0) | (nibble << 2)
 *   bit 13 ENTER (bufferpos = 0 before storing) | bit 14 EMIT (frame closed)
 *   byte 2      k = bits stored (0..4)
 *   byte 3      the stored bits (first at bit 24) */
constexpr uint32_t N_ROW = 0x1fc0u, N_ENTER = 1u << 13, N_EMIT = 1u << 14;

__host__ __device__ inline uint32_t hdlc_nibble_entry(uint32_t id, uint32_t v)
{
	uint32_t k = 0, bits = 0, flags = 0, p = 0;
	for (uint32_t i = 0; i < 4; i++) {
		const uint32_t b = (v >> i) & 1u, e = hdlc_transition(id, b);
		id = e & 0x7fu;
		if (e & H_STORE) { bits |= b << k; k++; }
		if (e & H_ENTER) { flags |= N_ENTER; k = 0; bits = 0; }
		if (e & H_EMIT) { flags |= N_EMIT; p = i; }
	}
	return p | (id << 6) | flags | (k << 16) | (bits << 24);
}

/* candidate layout (64 B, same slot a gais_msg will occupy): words 0..13 stored bits (LSB first),
 * word 14 = bufferpos | stop_bit << 16 , word 15 = closing bit index.  Only frames that reach the CRC get a
 * slot: a closing bit of 1 or bufferpos - 22 <= 0 is lostframes2 (src/protodec.c:1095-1113) whatever the
 * bits are, and is counted by the tracker itself -- a bit stream can close such frames every ~30 bits,
 * CRC candidates need >= 54 */
/* everything by value: taking the address of the per-lane FSM registers would push them to local memory */
__device__ __noinline__ uint32_t hdlc_emit(uint32_t pos, uint32_t shi, uint32_t b, uint32_t bit_index, const ChanState *s, int c,
					   uint32_t ncand, gais_msg *slots, int slot_cap, int32_t *overflow)
{
	const uint32_t nw = pos >> 5, part = (pos & 31u) ? shi >> (32u - (pos & 31u)) : 0u;
	if (ncand >= (uint32_t) slot_cap) {
		/* no slot left (only reachable with far more frames per run than AIS slots allow; the default capacity
		 * is one candidate per 1024 samples).  The frame is still checked, here, bit by bit (src/protodec.c:106-167),
		 * so that the COUNTERS stay the reference's; a CRC-ok frame has nowhere to go, which is what
		 * GAIS_EOVERFLOW reports */
		const int nbytes = (int) ((pos - 22u) >> 3) + 2;
		uint32_t crc = 0xffffu;
		for (int j = 0; j < nbytes * 8; j++) {
			const uint32_t wi = (uint32_t) j >> 5, word = wi < nw ? s->store[wi] : (wi == nw ? part : 0u);
			const uint32_t bit = (word >> (j & 31)) & 1u;
			crc = ((crc ^ bit) & 1u) ? (crc >> 1) ^ 0x8408u : crc >> 1;
		}
		ChanState *sw = const_cast<ChanState *>(s);
		if ((~crc & 0xffffu) == 0x0f47u) {
			sw->ok++;
			*overflow = 1;
		} else
			sw->crcfail++;
		return ncand;
	}
	uint32_t *w = reinterpret_cast<uint32_t *>(&slots[(int64_t) c * slot_cap + ncand]);
#pragma unroll
	for (uint32_t i = 0; i < 14; i++)
		w[i] = (i < nw) ? s->store[i] : (i == nw ? part : 0u);
	w[14] = pos | (b << 16);
	w[15] = bit_index;
	return ncand + 1u;
}

/* bits [i0, i1) of W one at a time (tile tails, and nibbles in which a frame outgrows the buffer) */
struct HdlcSerialRet { HdlcRegs f; uint32_t ncand, nsize; };

__device__ __noinline__ HdlcSerialRet hdlc_bits_serial(HdlcRegs f, const uint16_t *__restrict__ tab, uint32_t W, uint32_t i0, uint32_t i1,
						       uint32_t hb, ChanState *s, int c, uint32_t ncand, gais_msg *slots, int slot_cap,
						       int32_t *overflow)
{
	uint32_t nsize = 0;
	for (uint32_t i = i0; i < i1; i++) {
		const uint32_t b = (W >> i) & 1u;
		const uint32_t e = tab[f.id * 2u + b];
		f.id = e & 0x7fu;
		if (e & H_STORE) {                                            /* src/protodec.c:1016-1024 */
			f.slo = (f.slo >> 1) | (f.shi << 31);
			f.shi = (f.shi >> 1) | (b << 31);
			f.pos++;
			if ((f.pos & 31u) == 0u)
				s->store[(f.pos >> 5) - 1u] = f.shi;
			else if (f.pos >= 449u) {
				f.id = hdlc_hunt_id(0, 0, b);                         /* frame too long: protodec_reset() */
				f.pos = 0;
			}
		}
		if (e & (H_ENTER | H_EMIT)) {
			if (e & H_EMIT) {
				if (b | (uint32_t) (f.pos <= 22u))
					nsize++;
				else
					ncand = hdlc_emit(f.pos, f.shi, b, hb + i, s, c, ncand, slots, slot_cap, overflow);
			}
			f.pos = 0;
		}
	}
	HdlcSerialRet r;
	r.f = f;
	r.ncand = ncand;
	r.nsize = nsize;
	return r;
}

/* n (1..31) NRZI bits, bit 0 of W the oldest; hb = index of that bit in the channel's stream.
 * Whole nibbles go through the nibble table; `tail` says whether the 1..3 left-over bits are
 * consumed too (end of a tile) or left to the caller. Returns the number of bits consumed. */
__device__ __forceinline__ uint32_t hdlc_chunk(HdlcRegs &f, const uint16_t *__restrict__ tab, const uint32_t *__restrict__ ntab,
					       uint32_t W, uint32_t n, uint32_t hb, bool tail, ChanState *s, int c, uint32_t &ncand,
					       uint32_t &nsize, const TrackOut &out, uint32_t wmask)
{
	/* called by all lanes of `wmask` together (the votes below are over exactly that set) */
	const uint32_t nn = n >> 2;              /* whole nibbles */
	const uint32_t used = tail ? n : nn * 4u;
	/* warp-uniform shortcut for idle channels: while hunting, the FSM can only move on after more
	 * than 14 alternations ending in a 0 (src/protodec.c:1029-1037); a run-length test rules that
	 * out exactly */
	bool fast = false;
	uint32_t fast_id = 0;
	/* (one vote first: with a single lane of the warp inside a frame the shortcut is off, and the
	 * run-length test below -- ~50 instructions -- is not worth starting) */
	if (__all_sync(wmask, f.id < 64u || used == 0u)) {
		const uint32_t leak = f.id >> 5, nalt = (f.id >> 1) & 15u, last = f.id & 1u;
		const uint32_t vm = (used >= 32u) ? 0xffffffffu : (1u << used) - 1u;
		const uint32_t A = (W ^ ((W << 1) | last)) & vm;          /* bit i: b_i != b_(i-1) */
		const uint32_t L = (uint32_t) __ffs((int) ~A) - 1u;      /* leading alternations (<= used) */
		uint32_t r = A & (A >> 1);
		r &= r >> 2;
		r &= r >> 4;
		r &= r >> 7;                                             /* a run of >= 15 alternations inside */
		if (used == 0u) {
			fast = true;
			fast_id = f.id;
		} else if (nalt + L <= 14u && r == 0u) {
			fast = true;
			const uint32_t n2 = (L == used) ? nalt + used : (uint32_t) __clz((int) ~(A << (32u - used)));
			fast_id = hdlc_hunt_id(leak, n2, (W >> (used - 1u)) & 1u);
		}
	}
	if (__all_sync(wmask, fast)) {
		f.id = fast_id;
		return used;
	}
	if (used == 0u)
		return 0u;
	/* the state travels through the loop as the byte offset of its table row; W2 = W << 2 puts nibble q at
	 * bits 4q+2 .. 4q+5, where it is the entry index inside the row times 4 (the nibbles cover at most bits 0..27 of W) */
	uint32_t row = f.id << 6;
	const uint32_t W2 = W << 2;
	const char *ntab_b = reinterpret_cast<const char *>(ntab);
	for (uint32_t q = 0; q < nn; q++) {
		const uint32_t e = *reinterpret_cast<const uint32_t *>(ntab_b + (row | ((W2 >> (4u * q)) & 0x3cu)));
		const uint32_t k = __byte_perm(e, 0u, 0x4442u);           /* byte 2 */
		if (f.pos + k >= 449u && !(e & N_ENTER)) {
			f.id = row >> 6;
			const HdlcSerialRet r = hdlc_bits_serial(f, tab, W, 4u * q, 4u * q + 4u, hb, s, c, ncand, out.slots, out.slot_cap,
								 out.overflow);   /* rare: frame outgrows the buffer */
			f = r.f;
			ncand = r.ncand;
			nsize += r.nsize;
			row = f.id << 6;
			continue;
		}
		if (e & N_ENTER)
			f.pos = 0;
		row = e & N_ROW;
		const uint32_t bits = e >> 24, pos2 = f.pos + k;
		f.slo = __funnelshift_r(f.slo, f.shi, k);                 /* append k bits at the top of shi:slo */
		f.shi = __funnelshift_r(f.shi, bits, k);
		if ((f.pos ^ pos2) & 32u)                                 /* a 32-bit word of the frame is complete */
			s->store[(pos2 >> 5) - 1u] = __funnelshift_rc(f.slo, f.shi, 32u - (pos2 & 31u));
		f.pos = pos2;
		if (e & N_EMIT) {
			const uint32_t p = e & 3u, stopbit = (W >> (4u * q + p)) & 1u;
			if (stopbit | (uint32_t) (f.pos <= 22u))
				nsize++;                                      /* lostframes2: nothing to check, nothing to store */
			else
				ncand = hdlc_emit(f.pos, f.shi, stopbit, hb + 4u * q + p, s, c, ncand, out.slots, out.slot_cap, out.overflow);
			f.pos = 0;
		}
	}
	f.id = row >> 6;
	if (tail && used > nn * 4u) {
		const HdlcSerialRet r = hdlc_bits_serial(f, tab, W, nn * 4u, used, hb, s, c, ncand, out.slots, out.slot_cap, out.overflow);
		f = r.f;
		ncand = r.ncand;
		nsize += r.nsize;
	}
	return used;
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

#ifndef TRK_BLOCK
#define TRK_BLOCK 128
#endif
constexpr int TRK_THREADS = TRK_BLOCK;
#ifndef TRK_PREFETCH_DEPTH
#define TRK_PREFETCH_DEPTH 2
#endif
constexpr int TRK_PREFETCH = TRK_PREFETCH_DEPTH;   /* sign words loaded ahead of use: a word takes a warp ~1700 cycles, an L2 hit ~700 */
#ifndef TRK_MIN_BLOCKS
#define TRK_MIN_BLOCKS 1
#endif

/* the sign changes of one word (bit j of x: sample j differs from the sample before it), in time order */
__device__ __forceinline__ void dpll_word(uint32_t x, uint32_t &zb, uint32_t &dlo)
{
	while (x) {
		const uint32_t iso = x & (0u - x);
		const uint32_t j = 31u - (uint32_t) __clz((int) iso);
		x ^= iso;
		/* samples 0 .. j-1 of the word: no sign change since the last one (src/receiver.c:121-134 only),
		 * so the register at sample j is one 32-bit multiply-add away from the word's base */
		const uint32_t zj = j * GAIS_PLL_INC + zb;
		/* sample j changes sign: the pending NRZI difference bit flips, the phase is nudged towards the
		 * crossing (src/receiver.c:113-119) before this sample's own increment.  The nudge goes into the
		 * base: (zb + nudge) + j * 13107 = zj + nudge, which neither carries nor borrows */
		dlo ^= 1u << (zj >> 16);
		zb += (zj & 0x8000u) ? (0u - GAIS_PLL_NUDGE) : GAIS_PLL_NUDGE;
	}
}

__global__ void __launch_bounds__(TRK_THREADS, TRK_MIN_BLOCKS)
track_kernel(const uint32_t *__restrict__ signs, ChanState *st, int c_begin, int c_end, int n_channels, int64_t n_frames, TrackOut out)
{
	/* channels [c_begin, c_end) of a batch of n_channels (= the row length of the sign words) */
	__shared__ uint32_t ntab[H_NSTATES * 16];
	__shared__ uint16_t tab[H_NSTATES * 2];
	for (int i = threadIdx.x; i < H_NSTATES * 16; i += TRK_THREADS)
		ntab[i] = hdlc_nibble_entry((uint32_t) i >> 4, (uint32_t) i & 15u);
	for (int i = threadIdx.x; i < H_NSTATES * 2; i += TRK_THREADS)
		tab[i] = (uint16_t) hdlc_transition((uint32_t) i >> 1, (uint32_t) i & 1u);
	__syncthreads();

	const int c = c_begin + blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= c_end)
		return;
	ChanState *s = &st[c];
	/* the lanes that are left walk the same number of words and take the HDLC hand-over together: every vote in
	 * hdlc_chunk() is over exactly this set, taken outside any per-lane branch */
	const uint32_t wmask = __activemask();

	uint32_t prevword = (uint32_t) s->prev << 31;   /* bit 31 = sign of the last sample seen */
	uint32_t dlo = s->dacc, nd = s->nd;             /* difference bits not yet given to the FSM */
	uint32_t hb = s->n_bits - nd;                   /* stream index of dlo bit 0 */
	/* the phase register at sample 0 of the current word: DPLL phase in the lower half, slices since
	 * the last hand-over (hb) in the upper half.  At sample j of the word it is zb + j * 13107 */
	uint32_t zb = (nd << 16) | (s->pll & 0xffffu);
	HdlcRegs f;
	f.id = s->fsm; f.pos = s->pos; f.shi = s->cur; f.slo = s->cur2;
	uint32_t ncand = out.run_count[c], nsize = 0;
	const uint32_t bits_start = hb + nd;
	const uint32_t run_start = bits_start - out.run_bits[c];   /* stream index of the run's first bit (all of this modulo 2^32) */

	const int n_words = (int) ((n_frames + 31) >> 5);                    /* a tile is far below 2^31 words */
	const uint32_t *sp = signs + c;
	uint32_t q[TRK_PREFETCH];
#pragma unroll
	for (int k = 0; k < TRK_PREFETCH; k++)
		q[k] = (k < n_words) ? sp[(int64_t) k * n_channels] : 0u;
	const uint32_t *pf = sp + (int64_t) TRK_PREFETCH * n_channels;      /* next word to prefetch */

	const int n_full = (int) (n_frames >> 5);          /* whole 32-sample words; a ragged end comes after the loop */
	for (int w = 0; w < n_full; w++, pf += n_channels) {
		const uint32_t sw = q[0];
#pragma unroll
		for (int k = 0; k + 1 < TRK_PREFETCH; k++)
			q[k] = q[k + 1];
		q[TRK_PREFETCH - 1] = (w + TRK_PREFETCH < n_words) ? *pf : 0u;
		/* LSB-first words: bit 0 is the first sample.  x marks samples whose sign differs from the
		 * sample before */
		const uint32_t x = sw ^ __funnelshift_l(prevword, sw, 1);
		prevword = sw;
		dpll_word(x, zb, dlo);
		zb += 32u * GAIS_PLL_INC;                /* base of the next word */
		nd = zb >> 16;
		if (__any_sync(wmask, nd >= 24u)) {
			/* some lane has 24 bits: the whole warp hands its whole nibbles over (at most 7 slices per 32 samples
			 * and at most 3 bits left over from the last hand-over, so bit 31 of dlo is never reached before this
			 * point) */
			const uint32_t W = ~dlo;
			if (out.bits && nd >= 4u)
				bits_or(out, c, (int64_t) (int32_t) (hb - run_start), W, nd & ~3u);
			const uint32_t used = hdlc_chunk(f, tab, ntab, W, nd, hb, false, s, c, ncand, nsize, out, wmask);
			dlo >>= used;
			hb += used;
			nd -= used;
			zb -= used << 16;
		}
	}
	if (n_frames & 31) {
		/* ragged end of the tile: the last word holds 1..31 samples.  What it slices is handed over by the
		 * flush below (at most 23 + 7 bits are pending, dlo has room for them) */
		const uint32_t nb = (uint32_t) (n_frames & 31), sw = q[0];
		const uint32_t x = (sw ^ __funnelshift_l(prevword, sw, 1)) & ((1u << nb) - 1u);
		prevword = sw << (32u - nb);
		dpll_word(x, zb, dlo);
		zb += nb * GAIS_PLL_INC;
		nd = zb >> 16;
	}
	if (__any_sync(wmask, nd != 0u)) {
		/* end of the tile: hand ALL sliced bits over now, so that FSM state, candidates and
		 * counters at a run boundary are exactly the reference's after the same samples */
		const uint32_t W = ~dlo;
		if (out.bits && nd)
			bits_or(out, c, (int64_t) (int32_t) (hb - run_start), W, nd);
		hdlc_chunk(f, tab, ntab, W, nd, hb, true, s, c, ncand, nsize, out, wmask);
		dlo >>= nd;
		hb += nd;
		zb -= nd << 16;
		nd = 0;
	}

	s->pll = zb & 0xffffu; s->n_bits = hb + (zb >> 16); s->prev = (uint8_t) (prevword >> 31);
	s->dacc = dlo; s->nd = (uint8_t) nd;
	s->lastbit = (uint8_t) (((prevword >> 31) ^ (dlo >> nd)) & 1u);   /* sign at the last slice */
	s->fsm = (uint8_t) f.id; s->pos = (uint16_t) f.pos; s->cur = f.shi; s->cur2 = f.slo;
	out.run_count[c] = ncand;
	s->sizefail += (int32_t) nsize;
	out.run_bits[c] += hb + (zb >> 16) - bits_start;
}

/* The HDLC half alone: protodec_decode() (src/protodec.c:988-1122) for every channel of the batch, fed with
 * NRZI-decoded bits, one per byte (0/1), exactly what receiver_run() hands over (src/receiver.c:126-131).
 * bits of channel c at bits[c * stride + i].  Same FSM registers, candidates and counters as track_kernel. */
__global__ void __launch_bounds__(TRK_THREADS)
hdlc_bits_kernel(const uint8_t *__restrict__ bits, int64_t stride, int64_t n_bits, ChanState *st, int n_channels, TrackOut out)
{
	__shared__ uint32_t ntab[H_NSTATES * 16];
	__shared__ uint16_t tab[H_NSTATES * 2];
	for (int i = threadIdx.x; i < H_NSTATES * 16; i += TRK_THREADS)
		ntab[i] = hdlc_nibble_entry((uint32_t) i >> 4, (uint32_t) i & 15u);
	for (int i = threadIdx.x; i < H_NSTATES * 2; i += TRK_THREADS)
		tab[i] = (uint16_t) hdlc_transition((uint32_t) i >> 1, (uint32_t) i & 1u);
	__syncthreads();
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_channels)
		return;
	ChanState *s = &st[c];
	const uint32_t wmask = __activemask();
	HdlcRegs f;
	f.id = s->fsm; f.pos = s->pos; f.shi = s->cur; f.slo = s->cur2;
	uint32_t hb = s->n_bits, ncand = out.run_count[c], nsize = 0;
	const uint8_t *row = bits + (int64_t) c * stride;
	for (int64_t off = 0; off < n_bits; off += 24) {
		const uint32_t n = (uint32_t) (n_bits - off < 24 ? n_bits - off : 24);
		uint32_t W = 0;
		for (uint32_t i = 0; i < n; i++)
			W |= (uint32_t) (row[off + i] & 1u) << i;
		hb += hdlc_chunk(f, tab, ntab, W, n, hb, off + 24 >= n_bits, s, c, ncand, nsize, out, wmask);
	}
	s->n_bits = hb;
	s->fsm = (uint8_t) f.id; s->pos = (uint16_t) f.pos; s->cur = f.shi; s->cur2 = f.slo;
	s->sizefail += (int32_t) nsize;
	out.run_count[c] = ncand;
	out.run_bits[c] += (uint32_t) n_bits;
}

/* ---- frame check: one warp per channel, one lane per candidate ---------------------------------
 * CRC-16 (src/protodec.c:106-167), the counters (src/protodec.c:1095-1115), the type gate and seqnr
 * (src/protodec.c:896-929) and the in-place compaction of the CRC-ok candidates into gais_msg records.
 * 32 candidates are checked at a time; their order-dependent parts -- the output position and the
 * sequence number, which only advances on gated types -- are prefix counts over warp ballots. */
__global__ void __launch_bounds__(256)
frame_check_kernel(gais_msg *__restrict__ slots, uint32_t *__restrict__ run_count, int slot_cap, ChanState *st, int n_channels)
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
	const int c = (int) (((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
	if (c >= n_channels)
		return;
	const uint32_t ncand = run_count[c];
	ChanState *s = &st[c];
	uint32_t seqnr = s->seqnr, nout = 0, n_ok = 0, n_crc = 0, n_size = 0;
	uint4 *row = reinterpret_cast<uint4 *>(slots + (int64_t) c * slot_cap);
	for (uint32_t base = 0; base < ncand; base += 32u) {
		const uint32_t k = base + lane;
		const bool valid = k < ncand;
		uint32_t w[16];
		if (valid) {
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const uint4 q = row[4 * k + i];
				w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
			}
		} else {
#pragma unroll
			for (int i = 0; i < 16; i++)
				w[i] = 0;
		}
		const uint32_t meta = w[14];
		const int nbits = (int) (meta & 0xffffu) - 22;                     /* src/protodec.c:1096 */
		uint32_t status = valid ? 2u : 3u;                                  /* 2 = lostframes2 */
		if (valid && ((meta >> 16) & 1u) == 0u && nbits > 0) {
			const int nbytes = (nbits >> 3) + 2;                           /* src/protodec.c:133-134 */
			uint32_t crc = 0xffffu;
			for (int j = 0; j < nbytes; j++) {
				const uint32_t byte = (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
				crc = (crc >> 8) ^ table[(crc ^ byte) & 0xffu];
			}
			status = ((~crc & 0xffffu) == 0x0f47u) ? 0u : 1u;              /* src/protodec.c:166 */
		}
		const uint32_t type = (w[0] & 0xffu) >> 2;
		const bool gate = status == 0u && type >= 1u && type <= 24u;
		const uint32_t m_ok = __ballot_sync(0xffffffffu, status == 0u), m_crc = __ballot_sync(0xffffffffu, status == 1u);
		const uint32_t m_size = __ballot_sync(0xffffffffu, status == 2u), m_gate = __ballot_sync(0xffffffffu, gate);
		if (status == 0u) {
			const uint32_t my_seq = (seqnr + (uint32_t) __popc(m_gate & lt)) % 10u;   /* src/protodec.c:924-926 */
			const uint32_t my_out = nout + (uint32_t) __popc(m_ok & lt);
			const int nb = nbits >> 3;
			const uint32_t flags = my_seq | ((gate ? 1u : 0u) << 4);
#pragma unroll
			for (int i = 0; i < 13; i++) {
				const int lo = 32 * i;
				if (8 * nb <= lo) w[i] = 0;
				else if (8 * nb < lo + 32) w[i] &= (1u << (8 * nb - lo)) - 1u;
			}
			w[13] = ((nb > 52) ? (w[13] & 0xffu) : 0u) | (flags << 8) | ((uint32_t) nbits << 16);
			w[14] = (uint32_t) c;
			/* w[15] already holds the closing bit index.  my_out <= k, and every candidate of this batch
			 * is already in registers, so compacting in place cannot overwrite anything unread */
#pragma unroll
			for (int i = 0; i < 4; i++)
				row[4 * my_out + i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
		}
		nout += (uint32_t) __popc(m_ok);
		seqnr = (seqnr + (uint32_t) __popc(m_gate)) % 10u;
		n_ok += (uint32_t) __popc(m_ok); n_crc += (uint32_t) __popc(m_crc); n_size += (uint32_t) __popc(m_size);
	}
	if (lane == 0u) {
		s->ok += (int32_t) n_ok; s->crcfail += (int32_t) n_crc; s->sizefail += (int32_t) n_size;
		s->seqnr = (uint8_t) seqnr;
		run_count[c] = nout;
	}
}

static inline int track_launch(const uint32_t *signs, ChanState *st, int c_begin, int c_end, int n_ch, int64_t n_frames, const TrackOut &out,
			       cudaStream_t stream)
{
	track_kernel<<<(c_end - c_begin + TRK_THREADS - 1) / TRK_THREADS, TRK_THREADS, 0, stream>>>(signs, st, c_begin, c_end, n_ch, n_frames, out);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

/* after the last tile of a run */
static inline int finalize_launch(ChanState *st, int n_ch, const TrackOut &out, cudaStream_t stream)
{
	const int64_t threads = (int64_t) n_ch * 32;
	frame_check_kernel<<<(unsigned) ((threads + 255) / 256), 256, 0, stream>>>(out.slots, out.run_count, out.slot_cap, st, n_ch);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

} /* namespace gais */
#endif

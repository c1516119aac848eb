/*
 * synth_core.h -- integer-only synthetic AIS/GMSK discriminator-audio generator, shared
 * verbatim by the host build (gcc) and the device build (nvcc) so both produce the same
 * int16 samples bit-for-bit.  This is the WORKLOAD generator of SURVEY.md 8(d); it has no
 * counterpart in the reference (gnuais ships no test signal).
 *
 * Signal model, per channel:
 *   * time is cut into TDMA slots of 256 symbols = 1280 samples (9600 bit/s at 48 kHz, 5
 *     samples per symbol); slots are handled in PAIRS (2560 samples) so a two-slot message
 *     (AIS type 5, 424 bits) can live inside one independently generated unit;
 *   * a burst is  8 ramp-up bits | 24 training bits 0101.. | flag 0x7E | bit-stuffed
 *     (data + CRC-16/X.25, each byte LSB first) | flag 0x7E | 2 tail bits;
 *   * NRZI (0 = toggle) -> +-1 levels -> Gaussian BT=0.4 symbol pulse (19 non-zero taps at
 *     5 samples/symbol, unit DC gain, Q14) -> * amplitude;
 *   * + Irwin-Hall(4 x u16) noise from a counter hash, scaled by noise_q16;
 *   * TX off (exact 0 before noise) outside bursts; clip to int16.
 *
 * What the decoder is expected to do with it is NOT encoded here: parity is always checked
 * against the oracle, never against the transmitted messages.
 */
#ifndef GAIS_SYNTH_CORE_H
#define GAIS_SYNTH_CORE_H

#include <stdint.h>

#ifdef __CUDACC__
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD static inline
#endif

#define GS_SLOT_SAMPLES 1280
#define GS_PAIR_SAMPLES 2560
#define GS_SYM_WORDS 16          /* up to 512 symbols per burst */
#define GS_MAX_SYM_1SLOT 248     /* 10 + 4 + 5*248 + 11 < 1280 */
#define GS_MAX_SYM_2SLOT 504     /* 10 + 4 + 5*504 + 11 < 2560 */

typedef struct gais_synth_params {
	uint64_t seed;
	int32_t amplitude;   /* peak level in int16 units (default 12000) */
	int32_t noise_q16;   /* round(sigma * 65536 / 37837.2); sigma=300 -> 520, 1500 -> 2598 */
	int32_t rho_q16;     /* P(slot carries a burst) * 65536 (default 32768) */
	int32_t jitter;      /* 1: burst start jitters 0..4 samples */
} gais_synth_params;

typedef struct gs_burst {
	uint32_t lv[GS_SYM_WORDS]; /* NRZ level of symbol k = bit (k&31) of lv[k>>5] */
	int32_t nsym;              /* 0: no burst */
	int32_t start;             /* first sample of symbol 0, relative to the pair start */
} gs_burst;

GS_HD uint32_t gs_mix(uint32_t x)
{
	x ^= x >> 16; x *= 0x7feb352dU;
	x ^= x >> 15; x *= 0x846ca68bU;
	x ^= x >> 16;
	return x;
}

GS_HD uint32_t gs_channel_key(uint64_t seed, uint32_t channel)
{
	return gs_mix((uint32_t) seed ^ gs_mix(channel * 0x9E3779B9U + (uint32_t) (seed >> 32) + 0x632BE5ABU));
}

/* Q14 response of one 5-sample symbol through the Gaussian BT=0.4 filter, taps -7..+11
 * around the first sample of the symbol (tools/make_pulse.py). */
GS_HD int32_t gs_pulse(int d)
{
	switch (d) {
	case 0: case 18: return 1;
	case 1: case 17: return 6;
	case 2: case 16: return 48;
	case 3: case 15: return 261;
	case 4: case 14: return 1026;
	case 5: case 13: return 2930;
	case 6: case 12: return 6213;
	case 7: case 11: return 10118;
	case 8: case 10: return 13193;
	case 9: return 14331;
	default: return 0;
	}
}

/* bit writer over the symbol-level array: applies NRZI as bits are appended */
typedef struct gs_writer {
	uint32_t *lv;
	int32_t n, cap;
	uint32_t level;
} gs_writer;

GS_HD void gs_put(gs_writer *w, uint32_t airbit)
{
	if (!airbit)
		w->level ^= 1U;
	if (w->n < w->cap) {
		if (w->level)
			w->lv[w->n >> 5] |= 1U << (w->n & 31);
		w->n++;
	}
}

GS_HD void gs_put_flag(gs_writer *w)
{
	gs_put(w, 0);
	for (int i = 0; i < 6; i++)
		gs_put(w, 1);
	gs_put(w, 0);
}

/* data bit i of the (pre-stuffing) on-air payload+FCS stream */
GS_HD uint32_t gs_data_bit(uint32_t hm, uint32_t byte0, int i)
{
	if (i < 8)
		return (byte0 >> i) & 1U;
	return (gs_mix(hm ^ ((uint32_t) (i >> 5) * 0x9E3779B1U + 0x51ED270BU)) >> (i & 31)) & 1U;
}

/*
 * Build one burst.  kind selects the message family, hm is the per-burst hash.
 *   two_slot = 0: kinds  0..79 168-bit types 1/2/3/18          80..91 other types, 48..168 bits
 *                        92..93 types 0/25/26/27/40/63 (168)    94..95 odd bit count
 *                        96..97 168-bit with one flipped bit    98 abort (>= 7 ones)   99 tiny frame, no FCS
 *   two_slot = 1: kinds  0..87 type 5, 424 bits  88..91 type 5 flipped bit  92..95 over-long
 *                        96..99 type 8, 240..392 bits
 */
GS_HD void gs_build_burst(gs_burst *b, uint32_t hm, int kind, int two_slot, int jitter)
{
	gs_writer w;
	int nbits, flip = -1, abort_at = -1, no_fcs = 0, i, ones = 0;
	uint32_t type, crc = 0xffffU, h2 = gs_mix(hm + 0x3C6EF372U);

	for (i = 0; i < GS_SYM_WORDS; i++)
		b->lv[i] = 0;
	w.lv = b->lv;
	w.n = 0;
	w.cap = two_slot ? GS_MAX_SYM_2SLOT : GS_MAX_SYM_1SLOT;
	w.level = 0;

	if (two_slot) {
		if (kind < 88) { type = 5; nbits = 424; }
		else if (kind < 92) { type = 5; nbits = 424; flip = (int) (h2 % 440U); }
		else if (kind < 96) { type = 5; nbits = 440; }
		else { type = 8; nbits = 8 * (30 + (int) (h2 % 20U)); }
	} else {
		if (kind < 80) {
			const uint32_t t4[4] = { 1, 2, 3, 18 };
			type = t4[h2 & 3U]; nbits = 168;
		} else if (kind < 92) {
			const uint32_t t8[8] = { 4, 6, 8, 12, 14, 21, 24, 9 };
			type = t8[h2 & 7U]; nbits = 8 * (6 + (int) ((h2 >> 8) % 16U));
		} else if (kind < 94) {
			const uint32_t t6[8] = { 0, 25, 26, 27, 40, 63, 25, 27 };
			type = t6[h2 & 7U]; nbits = 168;
		} else if (kind < 96) {
			type = 1; nbits = 96 + (int) (h2 % 64U);
		} else if (kind < 98) {
			type = 1; nbits = 168; flip = (int) (h2 % 184U);
		} else if (kind < 99) {
			type = 1; nbits = 168; abort_at = 40 + (int) (h2 % 100U);
		} else {
			type = 1; nbits = 10; no_fcs = 1;   /* flag, 10 bits, flag: bufferpos - 22 < 0 */
		}
	}

	for (i = 0; i < 8; i++)
		gs_put(&w, 0);
	for (i = 0; i < 24; i++)
		gs_put(&w, (uint32_t) (i & 1));
	gs_put_flag(&w);

	/* payload + FCS, bit-stuffed; FCS = ~crc sent LSB first (CRC-16/X.25, poly 0x8408) */
	for (i = 0; i < nbits + (no_fcs ? 0 : 16); i++) {
		uint32_t bit;
		if (i < nbits) {
			bit = gs_data_bit(hm, ((type << 2) | (h2 >> 30)) & 0xffU, i);
			crc = ((crc ^ bit) & 1U) ? (crc >> 1) ^ 0x8408U : crc >> 1;
		} else {
			bit = ((~crc) >> (i - nbits)) & 1U;
		}
		if (i == flip)
			bit ^= 1U;
		if (i == abort_at)
			break;
		gs_put(&w, bit);
		if (bit) {
			if (++ones == 5) {
				gs_put(&w, 0);
				ones = 0;
			}
		} else {
			ones = 0;
		}
	}
	if (abort_at >= 0) {
		for (i = 0; i < 9; i++)
			gs_put(&w, 1);
		gs_put(&w, 0);
	} else {
		gs_put_flag(&w);
	}
	gs_put(&w, 0);
	gs_put(&w, 0);

	b->nsym = w.n;
	b->start = 10 + (jitter ? (int32_t) ((h2 >> 12) % 5U) : 0);
}

/* Decide and build the (up to two) bursts of slot pair `pair` of the channel keyed ck. */
GS_HD void gs_build_pair(gs_burst bursts[2], uint32_t ck, uint32_t pair, const gais_synth_params *p)
{
	uint32_t hp = gs_mix(ck ^ gs_mix(pair * 0x9E3779B1U + 0x01234567U));
	uint32_t rho = (uint32_t) p->rho_q16;

	bursts[0].nsym = 0; bursts[0].start = 0;
	bursts[1].nsym = 0; bursts[1].start = 0;
	if ((hp & 0xffffU) < ((rho * 5243U) >> 16)) {   /* 8 % of bursts are two-slot */
		gs_build_burst(&bursts[0], gs_mix(hp ^ 0xA511E9B3U), (int) ((hp >> 16) % 100U), 1, p->jitter);
		return;
	}
	for (uint32_t s = 0; s < 2; s++) {
		uint32_t hs = gs_mix(hp + 0x68BC21EBU * (s + 1U));
		if ((hs & 0xffffU) < rho) {
			gs_build_burst(&bursts[s], gs_mix(hs ^ 0x2545F491U), (int) ((hs >> 16) % 100U), 0, p->jitter);
			bursts[s].start += (int32_t) s * GS_SLOT_SAMPLES;
		}
	}
}

/* noiseless signal of one burst at sample m (relative to the pair start), Q0 int16 units */
GS_HD int32_t gs_burst_signal(const gs_burst *b, int32_t m, int32_t amplitude)
{
	int32_t r, klo, khi, acc = 0;
	if (b->nsym == 0)
		return 0;
	r = m - b->start;                         /* sample relative to symbol 0 */
	if (r < -7 || r > 5 * (b->nsym - 1) + 11)
		return 0;
	/* symbol k contributes pulse tap d = r - 5k + 7, 0 <= d <= 18 */
	khi = (r + 7) / 5;
	klo = (r + 7 - 18 + 4 + 500) / 5 - 100;   /* ceil((r-11)/5) without negative division */
	if (klo < 0) klo = 0;
	if (khi > b->nsym - 1) khi = b->nsym - 1;
	for (int32_t k = klo; k <= khi; k++) {
		int32_t pv = gs_pulse(r - 5 * k + 7);
		acc += ((b->lv[k >> 5] >> (k & 31)) & 1U) ? pv : -pv;
	}
	/* amplitude * acc / 2^14, rounding toward zero on magnitudes (no signed shifts) */
	if (acc >= 0)
		return (int32_t) (((int64_t) amplitude * acc) >> 14);
	return -(int32_t) (((int64_t) amplitude * (-acc)) >> 14);
}

GS_HD int32_t gs_noise(uint32_t ck, uint32_t n, int32_t noise_q16)
{
	uint32_t h1 = gs_mix(ck ^ (n * 0x9E3779B1U + 0x7F4A7C15U));
	uint32_t h2 = gs_mix(h1 ^ 0x85EBCA6BU);
	int32_t ih = (int32_t) ((h1 & 0xffffU) + (h1 >> 16) + (h2 & 0xffffU) + (h2 >> 16)) - 131070;
	int64_t v = (int64_t) ih * noise_q16;
	return (v >= 0) ? (int32_t) (v >> 16) : -(int32_t) ((-v) >> 16);
}

GS_HD int16_t gs_sample(const gs_burst bursts[2], uint32_t ck, uint32_t n, int32_t m, const gais_synth_params *p)
{
	int32_t v = gs_burst_signal(&bursts[0], m, p->amplitude) + gs_burst_signal(&bursts[1], m, p->amplitude)
		  + gs_noise(ck, n, p->noise_q16);
	if (v > 32767) v = 32767;
	if (v < -32768) v = -32768;
	return (int16_t) v;
}

#endif /* GAIS_SYNTH_CORE_H */

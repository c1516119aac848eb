/*
 * oracle/gais_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See gais_oracle.h.
 *
 * A from-scratch CPU restatement of the gnuais receive chain, written from the behaviour of
 * the reference (every function cites the reference lines it follows) and pinned against the
 * unmodified reference objects by tests/test_oracle_vs_ref.py.
 *
 * Must be compiled WITHOUT fp contraction / fast-math (oracle/Makefile passes
 * -ffp-contract=off -fno-fast-math): the FIR sum is float32, multiply then add, strictly in
 * tap order, denormals honoured (src/filter.h:40-49; no -ffast-math in CMakeLists.txt).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#include <pthread.h>

#include <xmmintrin.h>

#include "gais_oracle.h"

#define NTAPS 36

/* The reference runs with the x86 default MXCSR (denormals honoured: taps 2/33 ARE
 * denormal).  A host process may have had FTZ/DAZ switched on by some fast-math library;
 * clear them for the duration of an oracle call. */
static unsigned mxcsr_enter(void)
{
	unsigned old = _mm_getcsr();
	_mm_setcsr(old & ~0x8040u);
	return old;
}
static void mxcsr_leave(unsigned old) { _mm_setcsr(old); }

/* src/receiver.c:39-49 -- the double literals as gcc rounds them to float32 (RN).  Taps 0,1
 * (and 34,35) underflow to +0.0f, taps 2/33 are the denormal 0x00000069. */
static const uint32_t tap_bits_half[18] = {
	0x00000000u, 0x00000000u, 0x00000069u, 0x0130bd6du, 0x0982c347u, 0x112a6907u,
	0x18439833u, 0x1ec5b74eu, 0x24b00698u, 0x2a0a0629u, 0x2ebea222u, 0x32e7e4d5u,
	0x36786fe0u, 0x396a68bfu, 0x3bc2cc99u, 0x3d8e92d5u, 0x3eb7cd8au, 0x3f50b242u,
};
static uint32_t tap_bits_full[NTAPS];
static float taps[NTAPS];
static pthread_once_t taps_once = PTHREAD_ONCE_INIT;

static void taps_init(void)
{
	for (int i = 0; i < NTAPS; i++) {
		tap_bits_full[i] = tap_bits_half[i < 18 ? i : 35 - i];
		memcpy(&taps[i], &tap_bits_full[i], 4);
	}
}

const uint32_t *goracle_tap_bits(void)
{
	pthread_once(&taps_once, taps_init);
	return tap_bits_full;
}

/* ---- HDLC / protocol state (src/protodec.h:30-34,44-71; src/protodec.c:87-100) --------- */

enum { FSM_HUNT = 1, FSM_PREAMBLE = 2, FSM_STARTFLAG = 3, FSM_DATA = 4, FSM_STOPFLAG = 5 };

#define STORE_BITS 450   /* DEMOD_BUFFER_LEN */

typedef struct {
	/* DSP: src/filter.h:57-62 (only the last 36 inputs matter), src/receiver.h:35-46 */
	float win[2 * NTAPS];    /* doubled ring: win[wp..wp+35] = x[n-36..n-1] */
	int wp;
	uint32_t pll;
	int prev, lastbit;
	/* FSM */
	int state, nflag, nalt, nones, stuffed, pos;
	int last;
	uint8_t store[STORE_BITS];
	uint8_t seqnr;
	int32_t ok, crcfail, sizefail;
	/* taps for the tests */
	uint64_t bit_index;
	uint8_t *bits; int64_t bits_cap, n_bits;
	uint8_t *signs; int64_t signs_cap, n_signs;
	char *nmea; int64_t nmea_cap, nmea_len;
	goracle_frame *frames; int64_t frames_cap, n_frames;
} chan_t;

/* src/protodec.c:87-100 */
static void fsm_reset(chan_t *c)
{
	c->state = FSM_HUNT;
	c->nflag = 0;
	c->nalt = 0;
	c->nones = 0;
	c->last = 0;
	c->stuffed = 0;
	c->pos = 0;
}

static void chan_init(chan_t *c)
{
	memset(c, 0, sizeof(*c));
	fsm_reset(c);   /* src/protodec.c:54-76, src/receiver.c:52-74: everything else starts at 0 */
}

/* src/protodec.c:106-118 */
uint16_t goracle_crc16(const uint8_t *data, unsigned len)
{
	uint16_t crc = 0xffff;
	for (unsigned j = 0; j < len; j++) {
		unsigned byte = data[j];
		for (int i = 0; i < 8; i++) {
			unsigned bit = (byte >> i) & 1u;
			crc = ((crc ^ bit) & 1u) ? (uint16_t) ((crc >> 1) ^ 0x8408u) : (uint16_t) (crc >> 1);
		}
	}
	return (uint16_t) ~crc;
}

/* MSB-first bit-field read over payload bytes where payload bit k = (bytes[k/8] >> (7-k%8)) & 1
 * for k < 8*nbytes and 0 beyond: the content of rbuffer after src/protodec.c:150-162 and the
 * fill-bit zeroing of src/protodec.c:909-914; field read is src/protodec.c:205-214. */
static unsigned payload_field(const uint8_t *bytes, int nbytes, int from, int size)
{
	unsigned v = 0;
	for (int i = 0; i < size; i++) {
		int k = from + i;
		unsigned bit = (k < 8 * nbytes) ? (bytes[k >> 3] >> (7 - (k & 7))) & 1u : 0u;
		v = (v << 1) | bit;
	}
	return v;
}

/*
 * src/protodec.c:896-929 (type gate, fill bits, seqnr) + src/protodec.c:780-894 (armouring).
 * Quirks kept: single-sentence messages always say channel 'A' and fill 0; multi-sentence
 * ones have an empty channel field and carry the fill count only on the last sentence; the
 * type gate returns BEFORE the seqnr bump.
 */
int goracle_nmea(const uint8_t *payload, int nbits, uint8_t *seqnr, char *out)
{
	int nbytes = nbits / 8, fill, total, nsent, pos = 0, w = 0;
	unsigned type = payload_field(payload, nbytes, 0, 6);

	if (type < 1 || type > 24)
		return 0;
	fill = (nbits % 6) ? 6 - nbits % 6 : 0;
	total = nbits + fill;
	nsent = (total <= 366) ? 1 : (total + 365) / 366;

	for (int s = 1; s <= nsent; s++) {
		char line[128];
		int k = 0;
		unsigned char cs = 0;
		k += sprintf(line + k, "AIVDM,%c,%c,", (char) ('0' + nsent), (char) ('0' + s));
		if (nsent > 1)
			k += sprintf(line + k, "%c,,", (char) ('0' + *seqnr));
		else
			k += sprintf(line + k, ",A,");
		for (int n = 0; n < 61 && pos < total; n++, pos += 6) {
			unsigned v = payload_field(payload, nbytes, pos, 6);
			line[k++] = (char) (v < 40 ? v + 48 : v + 56);
		}
		line[k++] = ',';
		line[k++] = (char) ('0' + ((nsent > 1 && s == nsent) ? fill : 0));
		for (int i = 0; i < k; i++)
			cs ^= (unsigned char) line[i];
		line[k] = 0;
		w += sprintf(out + w, "!%s*%02X\r\n", line, cs);
	}
	*seqnr = (uint8_t) ((*seqnr + 1) % 10);
	return w;
}

/* frame closed by the bit after the sixth one: src/protodec.c:1095-1115 + :120-167 */
static void fsm_frame_end(chan_t *c, int b)
{
	int nbits = c->pos - 22;
	goracle_frame fr;
	memset(&fr, 0, sizeof(fr));
	fr.end_bit = (uint32_t) c->bit_index;
	fr.nbits = (int16_t) nbits;

	if (b == 0 && nbits > 0) {
		uint8_t bytes[STORE_BITS / 8 + 2];
		int nb = nbits / 8;
		for (int j = 0; j < nb + 2; j++) {
			unsigned v = 0;
			for (int i = 0; i < 8; i++)
				v |= (unsigned) c->store[8 * j + i] << i;
			bytes[j] = (uint8_t) v;
		}
		if (goracle_crc16(bytes, (unsigned) nb + 2) == 0x0f47) {
			char text[256];
			int n;
			c->ok++;
			fr.status = 0;
			fr.nbytes = (uint8_t) nb;
			memcpy(fr.payload, bytes, (size_t) nb);
			n = goracle_nmea(bytes, nbits, &c->seqnr, text);
			if (c->nmea && c->nmea_len + n <= c->nmea_cap)
				memcpy(c->nmea + c->nmea_len, text, (size_t) n);
			c->nmea_len += n;
		} else {
			c->crcfail++;
			fr.status = 1;
		}
	} else {
		c->sizefail++;
		fr.status = 2;
	}
	if (c->frames && c->n_frames < c->frames_cap)
		c->frames[c->n_frames] = fr;
	c->n_frames++;
	fsm_reset(c);
}

/* one NRZI-decoded bit through the HDLC state machine: src/protodec.c:988-1122 */
static void fsm_bit(chan_t *c, int b)
{
	switch (c->state) {
	case FSM_DATA:                                           /* :993-1026 */
		if (c->stuffed) {
			if (b)
				c->state = FSM_STOPFLAG;
			else
				c->last = b;
			c->stuffed = 0;
		} else {
			if (b == c->last && b == 1) {
				if (++c->nones == 4) {
					c->stuffed = 1;
					c->nones = 0;
				}
			} else {
				c->nones = 0;
			}
			c->store[c->pos++] = (uint8_t) b;
			if (c->pos >= 449)
				fsm_reset(c);
		}
		break;
	case FSM_HUNT:                                           /* :1028-1041 */
		c->nalt = (b != c->last) ? c->nalt + 1 : 0;
		c->last = b;
		if (c->nalt > 14 && b == 0) {
			c->state = FSM_PREAMBLE;
			c->nalt = 0;
		}
		break;
	case FSM_PREAMBLE:                                       /* :1043-1070 */
		if (b != c->last && c->nflag == 0) {
			c->nalt++;
		} else if (b == 1) {
			if (c->nflag == 0) {
				c->nflag = 3;
				c->last = b;
			} else if (c->nflag == 5) {
				c->nflag = 6;
				c->nalt = 0;
				c->state = FSM_STARTFLAG;
			} else {
				c->nflag++;
			}
		} else {
			if (c->nflag == 0)
				c->nflag = 1;
			else
				fsm_reset(c);
		}
		break;
	case FSM_STARTFLAG:                                      /* :1072-1093 */
		if (c->nflag >= 7) {
			if (b == 0) {
				c->state = FSM_DATA;
				c->nflag = 0;
				c->nones = 0;
				memset(c->store, 0, sizeof(c->store));
				c->pos = 0;
			} else {
				fsm_reset(c);
			}
		} else if (b == 0) {
			fsm_reset(c);
		}
		c->nflag++;          /* :1092 -- runs after either branch, even after a reset */
		break;
	case FSM_STOPFLAG:                                       /* :1095-1115 */
		fsm_frame_end(c, b);
		break;
	}
	c->last = b;             /* :1119 -- always */
	c->bit_index++;
}

/* one input sample: src/filter.c:106-143 (window = the 36 samples BEFORE this one),
 * src/filter.h:40-49 (sequential float32 mul+add), src/receiver.c:109-135 (DPLL/slicer/NRZI) */
static void chan_sample(chan_t *c, int16_t x)
{
	const float *w = &c->win[c->wp];
	float sum = 0.0f;
	int cur;

	/* x86-64 SSE scalar float: every * and + rounds to float32; gcc neither reassociates
	 * nor contracts this loop without -ffast-math / -mfma */
	for (int i = 0; i < NTAPS; i++)
		sum += w[i] * taps[i];
	c->win[c->wp] = c->win[c->wp + NTAPS] = (float) x;
	c->wp = (c->wp + 1 == NTAPS) ? 0 : c->wp + 1;

	cur = (sum > 0);
	if (c->signs) {
		if (c->n_signs < c->signs_cap)
			c->signs[c->n_signs] = (uint8_t) cur;
		c->n_signs++;
	}
	if (cur != c->prev)
		c->pll += (c->pll < 0x8000u) ? 819u : (uint32_t) -819;
	c->prev = cur;
	c->pll += 13107u;
	if (c->pll > 0xffffu) {
		int b = (cur == c->lastbit) ? 1 : 0;
		c->lastbit = cur;
		c->pll &= 0xffffu;
		if (c->bits) {
			if (c->n_bits < c->bits_cap)
				c->bits[c->n_bits] = (uint8_t) b;
		}
		c->n_bits++;
		fsm_bit(c, b);
	}
}

int goracle_run(const int16_t *buf, int64_t n_frames, int num_ch, int ch_ofs, int chunk,
		uint8_t *bits, int64_t bits_cap, int64_t *n_bits,
		uint8_t *signs,
		char *nmea, int64_t nmea_cap, int64_t *nmea_len,
		int32_t *stats,
		goracle_frame *frames, int64_t frames_cap, int64_t *n_frames_out)
{
	chan_t *c;
	unsigned csr;
	if (chunk < 1 || chunk > 4096)     /* src/receiver.c:104-105 would abort() */
		return -1;
	pthread_once(&taps_once, taps_init);
	csr = mxcsr_enter();
	c = (chan_t *) malloc(sizeof(*c));
	if (!c)
		return -2;
	chan_init(c);
	c->bits = bits; c->bits_cap = bits_cap;
	c->signs = signs; c->signs_cap = n_frames;
	c->nmea = nmea; c->nmea_cap = nmea_cap;
	c->frames = frames; c->frames_cap = frames_cap;

	/* the result is chunk-size invariant (all state is carried); chunk is honoured only to
	 * mirror the driving loop of src/ais.c:214-248 */
	for (int64_t off = 0; off < n_frames; off += chunk) {
		int64_t len = (n_frames - off < chunk) ? n_frames - off : chunk;
		for (int64_t i = 0; i < len; i++)
			chan_sample(c, buf[(off + i) * num_ch + ch_ofs]);
	}

	if (stats) {
		stats[0] = c->ok; stats[1] = c->crcfail; stats[2] = c->sizefail;
		stats[3] = (int32_t) c->pll; stats[4] = c->prev; stats[5] = c->lastbit;
		stats[6] = c->state; stats[7] = c->seqnr;
	}
	if (n_bits) *n_bits = c->n_bits;
	if (nmea_len) *nmea_len = c->nmea_len;
	if (n_frames_out) *n_frames_out = c->n_frames;
	free(c);
	mxcsr_leave(csr);
	return 0;
}

/* ---- multi-threaded timing of the port (bench.py cpu_baseline, kind "port") ------------- */

struct bench_arg {
	const int16_t *buf;
	int64_t n_channels, n_samples;
	int chunk, tid, n_threads;
	int64_t ok;
};

/* test hook: the HDLC bit machine alone -- fsm_bit() fed with NRZI-decoded bits (one per byte) from the
 * reset state.  stats = {ok, crcfail, sizefail}; frames as in goracle_run().  Lets the tests compare the GPU
 * path's state tables with this FSM without going through audio. */
int goracle_fsm_bits(const uint8_t *bits, int64_t n_bits, int32_t stats[3], goracle_frame *frames, int64_t frames_cap,
		     int64_t *n_frames_out)
{
	chan_t *c = (chan_t *) malloc(sizeof(chan_t));
	if (!c)
		return -1;
	chan_init(c);
	c->frames = frames;
	c->frames_cap = frames_cap;
	for (int64_t i = 0; i < n_bits; i++)
		fsm_bit(c, bits[i] & 1);
	stats[0] = c->ok; stats[1] = c->crcfail; stats[2] = c->sizefail;
	if (n_frames_out)
		*n_frames_out = c->n_frames;
	free(c);
	return 0;
}

static void *bench_thread(void *p)
{
	struct bench_arg *a = (struct bench_arg *) p;
	chan_t *c = (chan_t *) malloc(sizeof(*c));
	unsigned csr = mxcsr_enter();
	for (int64_t ch = a->tid; ch < a->n_channels; ch += a->n_threads) {
		const int16_t *row = a->buf + ch * a->n_samples;
		chan_init(c);
		for (int64_t i = 0; i < a->n_samples; i++)
			chan_sample(c, row[i]);
		a->ok += c->ok;
	}
	free(c);
	mxcsr_leave(csr);
	return NULL;
}

double goracle_bench(const int16_t *buf, int64_t n_channels, int64_t n_samples, int n_threads, int chunk,
		     int64_t *ok_total)
{
	struct timespec t0, t1;
	pthread_t th[256];
	struct bench_arg args[256];
	int64_t ok = 0;

	pthread_once(&taps_once, taps_init);
	if (n_threads < 1) n_threads = 1;
	if (n_threads > 256) n_threads = 256;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int t = 0; t < n_threads; t++) {
		args[t] = (struct bench_arg) { buf, n_channels, n_samples, chunk, t, n_threads, 0 };
		pthread_create(&th[t], NULL, bench_thread, &args[t]);
	}
	for (int t = 0; t < n_threads; t++) {
		pthread_join(th[t], NULL);
		ok += args[t].ok;
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (ok_total) *ok_total = ok;
	return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Drives the UNMODIFIED gnuais reference objects (compiled where they lie under
 * /root/reference/src by oracle/Makefile, outputs only into oracle/_ref/) through the
 * reference's own entry points -- init_receiver()/receiver_run()/free_receiver()
 * (src/receiver.h:48-51) -- and taps what the parity tests compare against:
 *
 *   * NMEA sentences: a serial_state_t{fd} (src/serial.h:24-26) pointing at a memfd; the
 *     reference writes "!AIVDM...\r\n" to it (src/protodec.c:883-885).
 *   * NRZI-decoded bits fed to protodec_decode(): linked with -Wl,--wrap=protodec_decode in
 *     the *_tap.so variant only (src/receiver.c:126-130).
 *   * FIR output signs (filtered[i] > 0): -Wl,--wrap=filter_run_buf (src/receiver.c:107-111).
 *   * counters receivedframes/lostframes/lostframes2 and the final DPLL state, read from
 *     struct receiver / struct demod_state_t (src/receiver.h:35-46, src/protodec.h:44-71).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the libraries built from this file.  No reference source is copied here: the
 * reference headers are included from /root/reference/src at build time.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <time.h>
#include <pthread.h>
#include <sys/mman.h>
#include <xmmintrin.h>

#include "receiver.h"
#include "filter.h"
#include "protodec.h"
#include "cfg.h"
#include "hlog.h"
extern void protodec_deinit(struct demod_state_t *d);
#include <syslog.h>

/* ---- taps (only live in the *_tap.so link; harmless otherwise) ------------------------- */

struct tap_ctx {
	uint8_t *bits;
	int64_t bits_cap, n_bits;
	uint8_t *signs;
	int64_t signs_cap, n_signs;
	int peak;                /* largest value filter_run_buf() has returned (its maxval, src/filter.c:112-119) */
};
static __thread struct tap_ctx *tls_tap;

#ifdef GREF_WITH_TAPS
void __real_protodec_decode(char *in, int count, struct demod_state_t *d);
void __wrap_protodec_decode(char *in, int count, struct demod_state_t *d)
{
	struct tap_ctx *t = tls_tap;
	if (t && t->bits) {
		for (int i = 0; i < count; i++)
			if (t->n_bits < t->bits_cap)
				t->bits[t->n_bits++] = (uint8_t) in[i];
			else
				t->n_bits++;
	}
	__real_protodec_decode(in, count, d);
}

short __real_filter_run_buf(struct filter *f, short *in, float *out, int step, int len);
short __wrap_filter_run_buf(struct filter *f, short *in, float *out, int step, int len)
{
	short r = __real_filter_run_buf(f, in, out, step, len);
	struct tap_ctx *t = tls_tap;
	if (t && r > t->peak)
		t->peak = r;
	if (t && t->signs) {
		for (int i = 0; i < len; i++)
			if (t->n_signs < t->signs_cap)
				t->signs[t->n_signs++] = out[i] > 0;
			else
				t->n_signs++;
	}
	return r;
}
int gref_has_taps(void) { return 1; }
#else
int gref_has_taps(void) { return 0; }
#endif

/* silence the per-message stdout line of protodec_getdata (src/protodec.c:931-934): the
 * skip_type[] knob gates only printf + field decoders, never NMEA/seqnr/counters. */
void gref_set_quiet(int quiet)
{
	for (int i = 0; i <= MAX_AIS_PACKET_TYPE; i++)
		skip_type[i] = quiet ? 1 : 0;
	/* the "Level on ch A too high" notice of src/receiver.c:137-147 is wall-clock gated
	 * diagnostics on stderr, not a parity output: raise the hlog threshold above it */
	log_level = quiet ? LOG_ERR : LOG_INFO;
}

/* capture what the reference printf()s (the per-message line of src/protodec.c:934-985) */
static int cap_saved = -1, cap_fd = -1;
int gref_stdout_begin(void)
{
	fflush(stdout);
	cap_fd = memfd_create("gref_stdout", 0);
	cap_saved = dup(1);
	if (cap_fd < 0 || cap_saved < 0)
		return -1;
	dup2(cap_fd, 1);
	return 0;
}
int64_t gref_stdout_end(char *out, int64_t cap)
{
	int64_t got = 0;
	fflush(stdout);
	dup2(cap_saved, 1);
	close(cap_saved);
	off_t sz = lseek(cap_fd, 0, SEEK_END);
	lseek(cap_fd, 0, SEEK_SET);
	while (got < sz && got < cap) {
		ssize_t n = read(cap_fd, out + got, (size_t) ((sz < cap ? sz : cap) - got));
		if (n <= 0)
			break;
		got += n;
	}
	close(cap_fd);
	cap_fd = cap_saved = -1;
	return (int64_t) sz;
}

/*
 * The reference's per-message text for one CRC-ok frame: fills a demod_state_t the way
 * protodec_calculate_crc() leaves it (src/protodec.c:150-162: rbuffer[k] = payload bit k for the
 * nbits/8 whole bytes, 0 beyond) and calls protodec_getdata() (src/protodec.c:896-986) with stdout
 * captured.  Returns the number of bytes the reference printed.
 */
int64_t gref_getdata_text(const uint8_t *payload, int nbits, int seqnr, char chanid, char *out, int64_t cap)
{
	struct demod_state_t d;
	int nb = nbits / 8;
	int64_t n;
	protodec_initialize(&d, NULL, NULL, chanid);
	d.seqnr = (unsigned char) seqnr;
	memset(d.rbuffer, 0, DEMOD_BUFFER_LEN);
	for (int k = 0; k < 8 * nb && k < DEMOD_BUFFER_LEN; k++)
		d.rbuffer[k] = (payload[k >> 3] >> (7 - (k & 7))) & 1;
	if (gref_stdout_begin() != 0)
		return -1;
	protodec_getdata(nbits, &d);
	n = gref_stdout_end(out, cap);
	protodec_deinit(&d);
	return n;
}

/*
 * The reference's HDLC bit machine alone: protodec_initialize() + protodec_decode() (src/protodec.c:988-1122)
 * over NRZI-decoded bits, one per byte, as receiver_run() hands them over (src/receiver.c:126-131).
 * stats[3] = { receivedframes, lostframes, lostframes2 }.  The NMEA the decoder emits goes to `nmea_fd`
 * (a serial_state_t on that descriptor), -1 for none.
 */
int gref_fsm_bits(const uint8_t *bits, int64_t n_bits, int32_t stats[3], int nmea_fd)
{
	struct demod_state_t d;
	struct serial_state_t ser;
	char chunk[4096];
	memset(&ser, 0, sizeof(ser));
	ser.fd = nmea_fd;
	protodec_initialize(&d, nmea_fd >= 0 ? &ser : NULL, NULL, 'A');
	for (int64_t off = 0; off < n_bits; off += (int64_t) sizeof(chunk)) {
		int n = (int) ((n_bits - off < (int64_t) sizeof(chunk)) ? n_bits - off : (int64_t) sizeof(chunk));
		for (int i = 0; i < n; i++)
			chunk[i] = (char) (bits[off + i] & 1);
		protodec_decode(chunk, n, &d);
	}
	stats[0] = (int32_t) d.receivedframes;
	stats[1] = (int32_t) d.lostframes;
	stats[2] = (int32_t) d.lostframes2;
	protodec_deinit(&d);
	return 0;
}

/*
 * Run one reference receiver over frame-interleaved int16 audio exactly as main() does
 * (src/ais.c:214-248): receiver_run() per `chunk` frames.
 *
 * stats[8] = { receivedframes, lostframes, lostframes2, pll, prev, lastbit, fsm state, seqnr }
 * Returns 0 on success.
 */
int gref_run(const int16_t *buf, int64_t n_frames, int num_ch, int ch_ofs, int chunk,
	     uint8_t *bits, int64_t bits_cap, int64_t *n_bits,
	     uint8_t *signs,
	     char *nmea, int64_t nmea_cap, int64_t *nmea_len,
	     int32_t *stats)
{
	struct serial_state_t ser;
	struct tap_ctx tap;
	struct receiver *rx;
	int64_t off;

	if (chunk < 1 || chunk > 4096)
		return -1;
	/* reference semantics = x86 default MXCSR: denormals honoured (taps 2/33 are denormal) */
	unsigned csr = _mm_getcsr();
	_mm_setcsr(csr & ~0x8040u);

	ser.fd = memfd_create("gref_nmea", 0);
	if (ser.fd < 0)
		return -2;

	memset(&tap, 0, sizeof(tap));
	tap.bits = bits;
	tap.bits_cap = bits_cap;
	tap.signs = signs;
	tap.signs_cap = n_frames;
	tls_tap = &tap;

	rx = init_receiver('A', num_ch, ch_ofs, &ser, NULL);
	for (off = 0; off < n_frames; off += chunk) {
		int len = (n_frames - off < chunk) ? (int) (n_frames - off) : chunk;
		/* receiver_run() takes a non-const pointer but only reads (src/receiver.c:102-107) */
		receiver_run(rx, (short *) (buf + off * num_ch), len);
	}
	fflush(stdout);
	tls_tap = NULL;

	if (stats) {
		struct demod_state_t *d = rx->decoder;
		stats[0] = d->receivedframes;
		stats[1] = d->lostframes;
		stats[2] = d->lostframes2;
		stats[3] = (int32_t) rx->pll;
		stats[4] = rx->prev;
		stats[5] = rx->lastbit;
		stats[6] = d->state;
		stats[7] = d->seqnr;
	}
	if (n_bits)
		*n_bits = tap.n_bits;

	if (nmea_len) {
		off_t sz = lseek(ser.fd, 0, SEEK_END);
		*nmea_len = (int64_t) sz;
		if (nmea && sz > 0) {
			int64_t want = sz < nmea_cap ? sz : nmea_cap, got = 0;
			lseek(ser.fd, 0, SEEK_SET);
			while (got < want) {
				ssize_t n = read(ser.fd, nmea + got, (size_t) (want - got));
				if (n <= 0)
					break;
				got += n;
			}
		}
	}
	close(ser.fd);
	/* free_receiver() leaks rx->decoder by design (src/receiver.c:76-82); leave it so. */
	free_receiver(rx);
	_mm_setcsr(csr);
	return 0;
}

/*
 * The level of a run: the largest value filter_run_buf() returned over the run's receiver_run() calls
 * (src/receiver.c:107 keeps it as maxval; src/filter.c:112-119 starts each call at 0, so negative samples
 * never count).  Tap library only; returns -1 without taps.
 */
int gref_peak(const int16_t *buf, int64_t n_frames, int num_ch, int ch_ofs, int chunk)
{
#ifdef GREF_WITH_TAPS
	struct tap_ctx tap;
	struct receiver *rx;
	memset(&tap, 0, sizeof(tap));
	tls_tap = &tap;
	rx = init_receiver('A', num_ch, ch_ofs, NULL, NULL);
	for (int64_t off = 0; off < n_frames; off += chunk) {
		int len = (n_frames - off < chunk) ? (int) (n_frames - off) : chunk;
		receiver_run(rx, (short *) (buf + off * num_ch), len);
	}
	fflush(stdout);
	tls_tap = NULL;
	free_receiver(rx);
	return tap.peak;
#else
	(void) buf; (void) n_frames; (void) num_ch; (void) ch_ofs; (void) chunk;
	return -1;
#endif
}

/* ---- multi-threaded CPU baseline ------------------------------------------------------- */

struct bench_arg {
	const int16_t *buf;
	int64_t n_channels, n_samples;
	int chunk, tid, n_threads;
	int64_t ok, samples;
};

static void *bench_thread(void *p)
{
	struct bench_arg *a = (struct bench_arg *) p;
	_mm_setcsr(_mm_getcsr() & ~0x8040u);
	for (int64_t c = a->tid; c < a->n_channels; c += a->n_threads) {
		const int16_t *row = a->buf + c * a->n_samples;
		struct receiver *rx = init_receiver('A', 1, 0, NULL, NULL);
		for (int64_t off = 0; off < a->n_samples; off += a->chunk) {
			int len = (a->n_samples - off < a->chunk) ? (int) (a->n_samples - off) : a->chunk;
			receiver_run(rx, (short *) (row + off), len);
		}
		a->ok += rx->decoder->receivedframes;
		a->samples += a->n_samples;
		free_receiver(rx);
	}
	return NULL;
}

/*
 * Time the reference chain on planar [n_channels][n_samples] int16, one struct receiver per
 * channel, channels dealt round-robin to n_threads pthreads; 1020-frame chunks match
 * src/ais.c:179-181.  serial = ipc = NULL.  The caller is expected to have pointed fd 1 at
 * /dev/null (the per-message printf of src/protodec.c:934 is part of the reference's work).
 * Returns wall seconds of the decode loop only (CLOCK_MONOTONIC).
 */
double gref_bench(const int16_t *buf, int64_t n_channels, int64_t n_samples, int n_threads, int chunk,
		  int64_t *ok_total)
{
	struct timespec t0, t1;
	pthread_t th[256];
	struct bench_arg args[256];
	int64_t ok = 0;

	if (n_threads < 1)
		n_threads = 1;
	if (n_threads > 256)
		n_threads = 256;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int t = 0; t < n_threads; t++) {
		args[t] = (struct bench_arg) { buf, n_channels, n_samples, chunk, t, n_threads, 0, 0 };
		pthread_create(&th[t], NULL, bench_thread, &args[t]);
	}
	for (int t = 0; t < n_threads; t++) {
		pthread_join(th[t], NULL);
		ok += args[t].ok;
	}
	fflush(stdout);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (ok_total)
		*ok_total = ok;
	return (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
}

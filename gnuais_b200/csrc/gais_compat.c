/*
 * gais_compat.c -- init_receiver()/receiver_run()/free_receiver() over the batched C-ABI.
 * See include/gais_compat.h for what is replaced and why decoding is deferred.
 *
 * Sinks: the reference pushes every sentence to serial_write() ("!%s\r\n",
 * src/protodec.c:883-885) and ipc_write() ("!%s", src/protodec.c:886-888) on the caller's
 * thread.  Those functions belong to the host program (src/serial.c, src/ipc.c); they are
 * bound weakly here and called in the same order when present.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gais_b200.h"
#include "gais_compat.h"

extern int serial_write(struct serial_state_t *state, char *s, int len) __attribute__((weak));
extern int ipc_write(struct ipc_state_t *ipc, char *buffer, int buflength) __attribute__((weak));

struct shim {
	gais_ctx *ctx;
	int16_t *queue;      /* mono samples of this receiver's channel */
	int64_t queued, batch;
	gais_msg *msgs;
	int64_t msgs_cap;
	int print_stdout;
};

static void die(const char *what)
{
	/* the reference has no error channel on this path: hlog(LOG_CRIT)+abort */
	fprintf(stderr, "gnuais-b200: %s: %s\n", what, gais_last_error());
	abort();
}

struct receiver *init_receiver(char name, int num_ch, int ch_ofs, struct serial_state_t *serial, struct ipc_state_t *ipc)
{
	struct receiver *rx = (struct receiver *) calloc(1, sizeof(*rx));
	struct demod_state_t *d = (struct demod_state_t *) calloc(1, sizeof(*d));
	struct shim *s = (struct shim *) calloc(1, sizeof(*s));
	const char *e = getenv("GAIS_SHIM_BATCH_FRAMES");
	gais_config cfg;

	if (!rx || !d || !s)
		exit(1);                              /* hmalloc() exits on OOM, src/hmalloc.c:40-55 */
	s->batch = (e && atoll(e) > 0) ? atoll(e) : 48000;
	s->print_stdout = !(getenv("GAIS_SHIM_STDOUT") && atoi(getenv("GAIS_SHIM_STDOUT")) == 0);
	s->queue = (int16_t *) malloc(sizeof(int16_t) * (size_t) (s->batch + 4096));
	if (!s->queue)
		exit(1);
	memset(&cfg, 0, sizeof(cfg));
	cfg.abi_version = GAIS_ABI_VERSION;
	cfg.device = getenv("GAIS_SHIM_DEVICE") ? atoi(getenv("GAIS_SHIM_DEVICE")) : 0;
	cfg.n_channels = 1;
	cfg.layout = GAIS_LAYOUT_PLANAR;
	cfg.max_frames_per_run = s->batch + 4096;
	cfg.fir_mode = GAIS_FIR_GUARD;
	if (gais_create(&cfg, &s->ctx) != 0)
		die("init_receiver");

	d->chanid = name;
	d->state = 1;                                 /* ST_SKURR, src/protodec.c:89 */
	d->serial = serial;
	d->ipc = ipc;
	rx->filter = (struct filter *) s;
	rx->decoder = d;
	rx->name = name;
	rx->num_ch = num_ch;
	rx->ch_ofs = ch_ofs;
	rx->pllinc = 0x10000 / 5;
	return rx;
}

void gais_compat_flush(struct receiver *rx)
{
	struct shim *s;
	struct demod_state_t *d;
	int64_t n = 0;
	gais_counters cnt;
	gais_chan_state st;

	if (!rx)
		return;
	s = (struct shim *) rx->filter;
	d = rx->decoder;
	if (s->queued == 0)
		return;
	if (gais_run_host(s->ctx, s->queue, s->queued, s->queued) != 0)
		die("receiver_run");
	s->queued = 0;
	if (gais_message_count(s->ctx, &n) != 0)
		die("receiver_run");
	if (n > s->msgs_cap) {
		free(s->msgs);
		s->msgs = (gais_msg *) malloc(sizeof(gais_msg) * (size_t) n);
		if (!s->msgs)
			exit(1);
		s->msgs_cap = n;
	}
	if (gais_get_messages(s->ctx, s->msgs, n, &n) != 0)
		die("receiver_run");
	for (int64_t i = 0; i < n; i++) {
		char text[GAIS_NMEA_STRIDE + 1];
		char line[1024];
		int len = gais_nmea_format(&s->msgs[i], text), pos = 0;
		while (pos < len) {                   /* one or two "!AIVDM...\r\n" sentences */
			int end = pos;
			while (end < len && text[end] != '\n')
				end++;
			end++;
			if (d->serial && serial_write)
				serial_write(d->serial, text + pos, end - pos);
			if (d->ipc && ipc_write)
				ipc_write(d->ipc, text + pos, end - pos - 2);
			pos = end;
		}
		/* the per-message stdout line of protodec_getdata() (src/protodec.c:934-985; skip_type[] is
		 * the host program's configuration and is not visible here: GAIS_SHIM_STDOUT=0 silences it) */
		if (s->print_stdout && gais_text_format(&s->msgs[i], d->chanid, line, (int) sizeof(line)) > 0) {
			fputs(line, stdout);
			fflush(stdout);
		}
	}
	if (gais_get_counters(s->ctx, &cnt) != 0 || gais_get_state(s->ctx, &st) != 0)
		die("receiver_run");
	d->receivedframes = cnt.ok;
	d->lostframes = cnt.crcfail;
	d->lostframes2 = cnt.sizefail;
	d->state = st.fsm_state;
	d->seqnr = (unsigned char) st.seqnr;
	rx->pll = st.pll;
	rx->prev = st.prev;
	rx->lastbit = st.lastbit;
}

void receiver_run(struct receiver *rx, short *buf, int len)
{
	struct shim *s = (struct shim *) rx->filter;

	if (len > 4096)
		abort();                              /* src/receiver.c:104-105 */
	buf += rx->ch_ofs;                            /* src/receiver.c:102 */
	for (int i = 0; i < len; i++)
		s->queue[s->queued + i] = buf[(int64_t) i * rx->num_ch];
	s->queued += len;
	if (s->queued >= s->batch)
		gais_compat_flush(rx);
}

void free_receiver(struct receiver *rx)
{
	struct shim *s;
	if (!rx)
		return;
	gais_compat_flush(rx);
	s = (struct shim *) rx->filter;
	gais_destroy(s->ctx);
	free(s->queue);
	free(s->msgs);
	free(s);
	/* the reference leaks rx->decoder (src/receiver.c:76-82 never calls protodec_deinit);
	 * callers read it after free_receiver() is NOT a pattern in ais.c, so release it */
	free(rx->decoder);
	free(rx);
}

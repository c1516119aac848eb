/*
 * gais_compat.c -- the reference's own entry points over the batched C-ABI (libgnuais_rx_b200.so):
 *   init_receiver() / receiver_run() / free_receiver()                     src/receiver.h:48-51
 *   protodec_initialize() / protodec_reset() / protodec_decode() / protodec_getdata()   src/protodec.h:73-76
 * See include/gais_compat.h for what is replaced and why decoding is deferred.
 *
 * Sinks: the reference pushes every sentence to serial_write() ("!%s\r\n", src/protodec.c:883-885) and
 * ipc_write() ("!%s", src/protodec.c:886-888) on the caller's thread, then prints the per-message line unless
 * skip_type[type] is set (src/protodec.c:931-985).  serial_write / ipc_write / skip_type belong to the host program
 * (src/serial.c, src/ipc.c, src/cfg.c:86); they are bound weakly here and used in the same order when present.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gais_b200.h"
#include "gais_compat.h"

extern int serial_write(struct serial_state_t *state, char *s, int len) __attribute__((weak));
extern int ipc_write(struct ipc_state_t *ipc, char *buffer, int buflength) __attribute__((weak));
extern int skip_type[] __attribute__((weak));          /* src/cfg.c:86, MAX_AIS_PACKET_TYPE + 1 = 25 entries */

#define DEMOD_BUFFER_LEN 450       /* src/protodec.h:41 */
#define SERBUFFER_LEN 100          /* src/protodec.c:52 */
#define IPCBUFFER_LEN 255          /* src/ipc.h:26 */

static void die(const char *what)
{
	/* the reference has no error channel on this path: hlog(LOG_CRIT)+abort */
	fprintf(stderr, "gnuais-b200: %s: %s\n", what, gais_last_error());
	abort();
}

static int stdout_enabled(void)
{
	const char *e = getenv("GAIS_SHIM_STDOUT");
	return !(e && atoi(e) == 0);
}

/* what protodec_getdata() does with one CRC-ok frame after the type gate (src/protodec.c:917-985): the sentences
 * to the sinks, then the stdout line */
static void emit_message(struct demod_state_t *d, const gais_msg *m, int print_stdout)
{
	char text[GAIS_NMEA_STRIDE + 1];
	char line[1024];
	int len = gais_nmea_format(m, text), pos = 0;
	while (pos < len) {                       /* one or two "!AIVDM...\r\n" sentences */
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
	if (len > 0 && print_stdout) {
		const int type = m->payload[0] >> 2;
		if (skip_type && type <= 24 && skip_type[type])
			return;                               /* ignored by configuration, src/protodec.c:931-932 */
		if (gais_text_format(m, d->chanid, line, (int) sizeof(line)) > 0) {
			fputs(line, stdout);
			fflush(stdout);
		}
	}
}

/* messages, counters and FSM state of the context's last run -> sinks and the caller-visible struct */
static void deliver(gais_ctx *ctx, struct demod_state_t *d, gais_msg **msgs, int64_t *msgs_cap, int print_stdout, gais_chan_state *st_out)
{
	int64_t n = 0;
	gais_counters cnt;
	gais_chan_state st;
	if (gais_message_count(ctx, &n) != 0)
		die("decode");
	if (n > *msgs_cap) {
		free(*msgs);
		*msgs = (gais_msg *) malloc(sizeof(gais_msg) * (size_t) n);
		if (!*msgs)
			exit(1);
		*msgs_cap = n;
	}
	if (gais_get_messages(ctx, *msgs, n, &n) != 0)
		die("decode");
	for (int64_t i = 0; i < n; i++)
		emit_message(d, &(*msgs)[i], print_stdout);
	if (gais_get_counters(ctx, &cnt) != 0 || gais_get_state(ctx, &st) != 0)
		die("decode");
	d->receivedframes = cnt.ok;
	d->lostframes = cnt.crcfail;
	d->lostframes2 = cnt.sizefail;
	d->state = st.fsm_state;
	d->seqnr = (unsigned char) st.seqnr;
	if (st_out)
		*st_out = st;
}

/* ---- receiver.h ------------------------------------------------------------------------------------------- */

struct shim {
	gais_ctx *ctx;
	int16_t *queue;      /* mono samples of this receiver's channel */
	int64_t queued, batch;
	gais_msg *msgs;
	int64_t msgs_cap;
	int print_stdout;
};

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
	s->print_stdout = stdout_enabled();
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
	/* one slot per frame that can possibly reach the CRC: >= 54 bits each (15 alternations, flag, 23 stored bits,
	 * flag), <= 13926 / 65536 bit per sample -- no input can overflow it, so the shim never aborts on a signal */
	cfg.reserved[0] = (int32_t) (cfg.max_frames_per_run / 254 + 2);
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
	gais_chan_state st;

	if (!rx)
		return;
	s = (struct shim *) rx->filter;
	if (s->queued == 0)
		return;
	if (gais_run_host(s->ctx, s->queue, s->queued, s->queued) != 0)
		die("receiver_run");
	s->queued = 0;
	deliver(s->ctx, rx->decoder, &s->msgs, &s->msgs_cap, s->print_stdout, &st);
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

/* ---- protodec.h: the bits -> frames -> NMEA half on its own ----------------------------------------------------
 * protodec_decode() queues the caller's bits (the reference's receiver hands them over one at a time,
 * src/receiver.c:126-131) and runs the GPU bit machine when GAIS_SHIM_BATCH_BITS (default 9600 = 1 s) are queued,
 * at protodec_reset(), and at gais_compat_flush_decoder().  The private state hangs off d->tbuffer, a field the
 * reference allocates nothing for and never reads (src/protodec.h:52). */

struct pshim {
	gais_ctx *ctx;
	uint8_t *queue;
	int64_t queued, batch;
	gais_msg *msgs;
	int64_t msgs_cap;
	int print_stdout;
};

void protodec_reset(struct demod_state_t *d)
{
	struct pshim *p = (struct pshim *) d->tbuffer;
	if (p) {
		gais_compat_flush_decoder(d);
		if (gais_reset_fsm(p->ctx) != 0)
			die("protodec_reset");
	}
	d->state = 1;                                 /* ST_SKURR; src/protodec.c:87-100 */
	d->nskurr = d->ndata = d->npreamble = d->nstartsign = d->nstopsign = 0;
	d->antallpreamble = d->antallenner = 0;
	d->last = 0;
	d->bitstuff = 0;
	d->bufferpos = 0;
}

void protodec_initialize(struct demod_state_t *d, struct serial_state_t *serial, struct ipc_state_t *ipc, char chanid)
{
	struct pshim *p = (struct pshim *) calloc(1, sizeof(*p));
	const char *e = getenv("GAIS_SHIM_BATCH_BITS");
	gais_config cfg;

	memset(d, 0, sizeof(*d));                     /* src/protodec.c:54-76 */
	d->chanid = chanid;
	d->serial = serial;
	d->ipc = ipc;
	d->state = 1;
	d->buffer = (unsigned char *) calloc(1, DEMOD_BUFFER_LEN);
	d->rbuffer = (unsigned char *) calloc(1, DEMOD_BUFFER_LEN);
	d->serbuffer = (char *) calloc(1, SERBUFFER_LEN);
	d->ipcbuffer = (char *) calloc(1, IPCBUFFER_LEN);
	d->nmea = (char *) calloc(1, SERBUFFER_LEN);
	if (!p || !d->buffer || !d->rbuffer || !d->serbuffer || !d->ipcbuffer || !d->nmea)
		exit(1);
	p->batch = (e && atoll(e) > 0) ? atoll(e) : 9600;
	p->print_stdout = stdout_enabled();
	p->queue = (uint8_t *) malloc((size_t) p->batch);
	if (!p->queue)
		exit(1);
	memset(&cfg, 0, sizeof(cfg));
	cfg.abi_version = GAIS_ABI_VERSION;
	cfg.device = getenv("GAIS_SHIM_DEVICE") ? atoi(getenv("GAIS_SHIM_DEVICE")) : 0;
	cfg.n_channels = 1;
	cfg.layout = GAIS_LAYOUT_PLANAR;
	cfg.max_frames_per_run = p->batch;
	cfg.fir_mode = GAIS_FIR_GUARD;
	cfg.reserved[0] = (int32_t) (p->batch / 54 + 2);      /* every frame that can reach the CRC gets a slot */
	if (gais_create(&cfg, &p->ctx) != 0)
		die("protodec_initialize");
	d->tbuffer = (char *) p;
}

void gais_compat_flush_decoder(struct demod_state_t *d)
{
	struct pshim *p = d ? (struct pshim *) d->tbuffer : NULL;
	if (!p || p->queued == 0)
		return;
	if (gais_run_bits_host(p->ctx, p->queue, p->queued, p->queued) != 0)
		die("protodec_decode");
	p->queued = 0;
	deliver(p->ctx, d, &p->msgs, &p->msgs_cap, p->print_stdout, NULL);
}

void protodec_decode(char *in, int count, struct demod_state_t *d)
{
	struct pshim *p = (struct pshim *) d->tbuffer;
	int i = 0;
	while (i < count) {
		int64_t room = p->batch - p->queued, n = count - i < room ? count - i : room;
		for (int64_t k = 0; k < n; k++)
			p->queue[p->queued + k] = (uint8_t) (in[i + k] & 1);
		p->queued += n;
		i += (int) n;
		if (p->queued >= p->batch)
			gais_compat_flush_decoder(d);
	}
	if (count > 0)
		d->last = in[count - 1];              /* src/protodec.c:1119 */
}

void gais_compat_free_decoder(struct demod_state_t *d)
{
	struct pshim *p = d ? (struct pshim *) d->tbuffer : NULL;
	if (!p)
		return;
	gais_compat_flush_decoder(d);
	gais_destroy(p->ctx);
	free(p->queue);
	free(p->msgs);
	free(p);
	d->tbuffer = NULL;
	free(d->buffer); free(d->rbuffer); free(d->serbuffer); free(d->ipcbuffer); free(d->nmea);   /* protodec_deinit, src/protodec.c:78-85 */
	d->buffer = d->rbuffer = NULL;
	d->serbuffer = d->ipcbuffer = d->nmea = NULL;
}

/* One CRC-ok frame whose payload bits the caller has put into d->rbuffer, one per byte (src/protodec.c:896-986):
 * type gate 1..24, fill bits, NMEA to the sinks with the decoder's current sequence number, seqnr advance, stdout
 * line.  Host arithmetic only -- this is the per-message tail of the path, not its data-parallel part. */
void protodec_getdata(int bufferlen, struct demod_state_t *d)
{
	gais_msg m;
	int type;
	memset(&m, 0, sizeof(m));
	if (bufferlen < 0 || bufferlen > 53 * 8)
		return;
	for (int k = 0; k < bufferlen / 8 * 8; k++)
		m.payload[k >> 3] |= (uint8_t) ((d->rbuffer[k] & 1) << (7 - (k & 7)));
	type = m.payload[0] >> 2;
	if (type < 1 || type > 24)
		return;                                   /* src/protodec.c:899-900: no NMEA, no seqnr advance */
	m.nbits = (uint16_t) bufferlen;
	m.flags = (uint8_t) ((d->seqnr % 10) | 16);
	emit_message(d, &m, stdout_enabled());
	d->seqnr++;                                   /* src/protodec.c:924-926 */
	if (d->seqnr > 9)
		d->seqnr = 0;
}

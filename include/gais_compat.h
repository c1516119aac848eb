/*
 * gais_compat.h -- link-compatible legacy entry points (libgnuais_rx_b200.so).
 *
 * gnuais's main() (src/ais.c:139-149, :236-248, :296-313) calls exactly three functions of the
 * receive path and reaches into two structs; this header declares them -- and the four protodec_*
 * functions of src/protodec.h:73-76 -- with the reference's own names and layouts so ais.c can be
 * relinked against the B200 path unchanged:
 *
 *   init_receiver()   replaces src/receiver.c:52-74   (declared src/receiver.h:48)
 *   free_receiver()   replaces src/receiver.c:76-82   (declared src/receiver.h:49)
 *   receiver_run()    replaces src/receiver.c:87-148  (declared src/receiver.h:51)
 *
 * struct receiver / struct demod_state_t keep the field order and types of
 * src/receiver.h:35-46 and src/protodec.h:44-71 because the caller reads
 * rx->decoder->{receivedframes,lostframes,lostframes2,chanid,best_range} directly
 * (src/ais.c:296-310, src/range.c:47-53).  Fields the B200 path does not use stay zero.
 *
 * Behavioural differences, all forced by batching: receiver_run() copies the
 * chunk and returns; decoding happens when GAIS_SHIM_BATCH_FRAMES (default 48000 = 1 s) frames
 * are queued and at free_receiver().  Messages reach serial_write()/ipc_write() later than in
 * the reference but in the same order with the same bytes.
 */
#ifndef GAIS_COMPAT_H
#define GAIS_COMPAT_H

#include <time.h>

#ifdef __cplusplus
extern "C" {
#endif

struct serial_state_t;
struct ipc_state_t;
struct filter;   /* opaque here: the shim keeps its private state behind this pointer */

struct demod_state_t {
	char chanid;
	int state;
	unsigned int offset;
	int nskurr, npreamble, nstartsign, ndata, nstopsign;
	int antallenner;
	unsigned char *buffer;
	unsigned char *rbuffer;
	char *tbuffer;
	int bufferpos;
	char last;
	int antallpreamble;
	int bitstuff;
	int receivedframes;
	int lostframes;
	int lostframes2;
	unsigned char seqnr;
	float best_range;
	struct serial_state_t *serial;
	struct ipc_state_t *ipc;
	char *serbuffer;
	char *ipcbuffer;
	char *nmea;
};

struct receiver {
	struct filter *filter;
	char name;
	int lastbit;
	int num_ch;
	int ch_ofs;
	unsigned int pll;
	unsigned int pllinc;
	struct demod_state_t *decoder;
	int prev;
	time_t last_levellog;
};

struct receiver *init_receiver(char name, int num_ch, int ch_ofs, struct serial_state_t *serial, struct ipc_state_t *ipc);
void free_receiver(struct receiver *rx);
void receiver_run(struct receiver *rx, short *buf, int len);

/* extra (not in the reference): decode whatever is queued now */
void gais_compat_flush(struct receiver *rx);

/*
 * The protocol decoder on its own, src/protodec.h:73-76 -- the bits -> HDLC frames -> CRC -> NMEA half of the path
 * for callers that bring their own demodulator:
 *
 *   protodec_initialize()  replaces src/protodec.c:54-76    (allocates the same caller-visible buffers)
 *   protodec_reset()       replaces src/protodec.c:87-100
 *   protodec_decode()      replaces src/protodec.c:988-1122 (in: `count` NRZI-decoded bits, one per byte)
 *   protodec_getdata()     replaces src/protodec.c:896-986  (one CRC-ok frame in d->rbuffer -> NMEA, sinks, seqnr)
 *
 * protodec_decode() queues bits and runs the GPU bit machine every GAIS_SHIM_BATCH_BITS (default 9600) bits, at
 * protodec_reset() and at gais_compat_flush_decoder(); counters, d->state and d->seqnr are those of the reference
 * after the same bits once the queue is flushed.  d->buffer / d->bufferpos and the per-bit fields (nstartsign,
 * antallenner, ...) are NOT maintained: the partially received frame lives on the device.
 */
void protodec_initialize(struct demod_state_t *d, struct serial_state_t *serial, struct ipc_state_t *ipc, char chanid);
void protodec_reset(struct demod_state_t *d);
void protodec_getdata(int bufferlengde, struct demod_state_t *d);
void protodec_decode(char *in, int count, struct demod_state_t *d);

/* extras (not in the reference): decode the queued bits now; release what protodec_initialize() allocated (the
 * reference's protodec_deinit(), src/protodec.c:78-85, is not declared in protodec.h) */
void gais_compat_flush_decoder(struct demod_state_t *d);
void gais_compat_free_decoder(struct demod_state_t *d);

#ifdef __cplusplus
}
#endif
#endif

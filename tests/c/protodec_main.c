/*
 * tests/c/protodec_main.c -- drives the protodec_* entry points of libgnuais_rx_b200.so (src/protodec.h:73-76) the way
 * a demodulator would: protodec_initialize(), then protodec_decode() with the NRZI bits of a file (one per byte) in
 * calls of 1, 7 and 4096 bits, one protodec_reset() in the middle when asked, and a serial AND an ipc sink that
 * append to files ("!%s\r\n" and "!%s", src/protodec.c:883-888).  Used by tests/test_shim_gpu.py.
 *
 *   protodec_main <bits file> <out prefix> <reset at bit, or -1>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>

#include "gais_compat.h"

struct serial_state_t { int fd; };
struct ipc_state_t { int fd; };
int serial_write(struct serial_state_t *state, char *s, int len) { return (int) write(state->fd, s, (size_t) len); }
/* src/ipc.c:121-134 broadcasts the buffer to every connected gnuaisgui client; here: one per line */
int ipc_write(struct ipc_state_t *ipc, char *buffer, int buflength)
{
	if (write(ipc->fd, buffer, (size_t) buflength) != buflength) return -1;
	return (int) write(ipc->fd, "\n", 1);
}
int skip_type[25];      /* the host program's configuration (src/cfg.c:86): type 5 is switched off below */

int main(int argc, char **argv)
{
	if (argc != 4) return 2;
	FILE *in = fopen(argv[1], "rb");
	long reset_at = atol(argv[3]);
	char path[512];
	struct serial_state_t ser;
	struct ipc_state_t ipc;
	struct demod_state_t d;
	static char buf[4096];
	long pos = 0;
	int sizes[3] = { 1, 7, 4096 }, k = 0;
	if (!in) return 2;
	skip_type[5] = 1;
	snprintf(path, sizeof(path), "%s.serial", argv[2]);
	ser.fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
	snprintf(path, sizeof(path), "%s.ipc", argv[2]);
	ipc.fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
	protodec_initialize(&d, &ser, &ipc, 'B');
	for (;;) {
		int want = sizes[k++ % 3];
		if (reset_at >= 0 && pos < reset_at && pos + want > reset_at)
			want = (int) (reset_at - pos);
		int n = (int) fread(buf, 1, (size_t) want, in);
		if (n <= 0) break;
		protodec_decode(buf, n, &d);
		pos += n;
		if (pos == reset_at)
			protodec_reset(&d);
	}
	gais_compat_flush_decoder(&d);
	printf("Received correctly: %d packets, wrong CRC: %d packets, wrong size: %d packets, seqnr %d, state %d\n", d.receivedframes,
	       d.lostframes, d.lostframes2, d.seqnr, d.state);
	gais_compat_free_decoder(&d);
	close(ser.fd);
	close(ipc.fd);
	fclose(in);
	return 0;
}

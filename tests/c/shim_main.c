/*
 * tests/c/shim_main.c -- drives libgnuais_rx_b200.so exactly the way gnuais' main() drives its
 * receivers (src/ais.c:139-149 init, :214-248 read loop with 1020-frame chunks, :296-313 stats +
 * free), with a serial sink that appends to a file.  Used by tests/test_shim_gpu.py.
 *
 *   shim_main <raw int16 file> <channels 1|2> <out prefix>
 * writes <prefix>.A.nmea [, <prefix>.B.nmea] and prints the frame counters.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>

#include "gais_compat.h"

/* the host program's sink (src/serial.c:110-122 has the same signature) */
struct serial_state_t { int fd; };
int serial_write(struct serial_state_t *state, char *s, int len) { return (int) write(state->fd, s, (size_t) len); }

int main(int argc, char **argv)
{
	if (argc != 4) return 2;
	FILE *in = fopen(argv[1], "rb");
	int channels = atoi(argv[2]);
	char path[512];
	struct serial_state_t ser[2];
	struct receiver *rx[2] = { NULL, NULL };
	if (!in || channels < 1 || channels > 2) return 2;
	for (int c = 0; c < channels; c++) {
		snprintf(path, sizeof(path), "%s.%c.nmea", argv[3], 'A' + c);
		ser[c].fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
		rx[c] = init_receiver((char) ('A' + c), channels, c, &ser[c], NULL);
	}
	int buffer_l = 1024 - 1024 % 5;                       /* src/ais.c:179-181 */
	short *buffer = malloc(sizeof(short) * (size_t) buffer_l * (size_t) channels);
	for (;;) {
		int n = (int) fread(buffer, sizeof(short) * (size_t) channels, (size_t) buffer_l, in);
		if (n <= 0) break;
		for (int c = 0; c < channels; c++)
			receiver_run(rx[c], buffer, n);
	}
	for (int c = 0; c < channels; c++) {
		gais_compat_flush(rx[c]);
		printf("%c: Received correctly: %d packets, wrong CRC: %d packets, wrong size: %d packets\n", 'A' + c,
		       rx[c]->decoder->receivedframes, rx[c]->decoder->lostframes, rx[c]->decoder->lostframes2);
		free_receiver(rx[c]);
		close(ser[c].fd);
	}
	free(buffer);
	fclose(in);
	return 0;
}

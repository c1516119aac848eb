/*
 * gais_nmea.h -- !AIVDM armouring of one CRC-ok frame record, shared by the host formatter
 * (gais_nmea_format) and the GPU kernel (nmea_kernel) so both give the same bytes.
 *
 * Behaviour follows protodec_getdata() (src/protodec.c:896-929: type gate 1..24, zero fill
 * bits up to a multiple of 6) and protodec_generate_nmea() (src/protodec.c:780-894: at most
 * 61 six-bit characters per sentence, "c<40 ? c+48 : c+56", single sentences always carry
 * channel 'A' and fill 0, multi-part sentences an empty channel field, the seqnr, and the
 * fill count on the last part only; checksum = XOR of everything between '!' and '*').
 * The serial form "!%s\r\n" of src/protodec.c:883 is what is produced.
 */
#ifndef GAIS_NMEA_H
#define GAIS_NMEA_H

#include <stdint.h>

#ifdef __CUDACC__
#define GN_HD __host__ __device__ __forceinline__
#else
#define GN_HD static inline
#endif

/* six payload bits starting at bit `from` (MSB-first; bits at or beyond 8*nbytes read 0,
 * src/protodec.c:150-162 + :909-914) */
GN_HD unsigned gn_sixbit(const uint8_t *payload, int nbytes, int from)
{
	unsigned v = 0;
	for (int i = 0; i < 6; i++) {
		int k = from + i;
		unsigned bit = (k < 8 * nbytes) ? (payload[k >> 3] >> (7 - (k & 7))) & 1u : 0u;
		v = (v << 1) | bit;
	}
	return v;
}

GN_HD char gn_hex(unsigned v) { return (char) (v < 10 ? '0' + v : 'A' + (v - 10)); }

/* returns the number of bytes written to out (<= 175); 0 when the type gate drops the frame */
GN_HD int gn_format(const uint8_t *payload, int nbits, int seqnr, char *out)
{
	int nbytes = nbits >> 3;
	unsigned type = gn_sixbit(payload, nbytes, 0);
	int fill, total, nsent, pos = 0, w = 0;

	if (type < 1 || type > 24)
		return 0;
	fill = (nbits % 6) ? 6 - nbits % 6 : 0;
	total = nbits + fill;
	nsent = (total <= 366) ? 1 : (total + 365) / 366;

	for (int s = 1; s <= nsent; s++) {
		unsigned cs = 0;
		int body = w + 1;
		out[w++] = '!';
		out[w++] = 'A'; out[w++] = 'I'; out[w++] = 'V'; out[w++] = 'D'; out[w++] = 'M'; out[w++] = ',';
		out[w++] = (char) ('0' + nsent); out[w++] = ',';
		out[w++] = (char) ('0' + s); out[w++] = ',';
		if (nsent > 1) {
			out[w++] = (char) ('0' + seqnr); out[w++] = ','; out[w++] = ',';
		} else {
			out[w++] = ','; out[w++] = 'A'; out[w++] = ',';
		}
		for (int n = 0; n < 61 && pos < total; n++, pos += 6) {
			unsigned v = gn_sixbit(payload, nbytes, pos);
			out[w++] = (char) (v < 40 ? v + 48 : v + 56);
		}
		out[w++] = ',';
		out[w++] = (char) ('0' + ((nsent > 1 && s == nsent) ? fill : 0));
		for (int i = body; i < w; i++)
			cs ^= (unsigned char) out[i];
		out[w++] = '*';
		out[w++] = gn_hex((cs >> 4) & 15u);
		out[w++] = gn_hex(cs & 15u);
		out[w++] = '\r';
		out[w++] = '\n';
	}
	return w;
}

#endif

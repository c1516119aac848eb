/*
 * oracle/gais_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement ("port") of the gnuais receive chain, used ONLY as the checker for the
 * CUDA path (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).  Nothing under
 * gnuais_b200/ may include, link or call it.
 *
 * Parity is PINNED: tests/test_oracle_vs_ref.py runs this port and the unmodified
 * reference objects (oracle/_ref, built by oracle/Makefile) on the same inputs and
 * requires identical FIR signs, NRZI bitstreams, NMEA bytes, counters and final DPLL
 * state; tests/golden/ holds reference-generated fixtures for the GPU box where
 * /root/reference does not exist.
 */
#ifndef GAIS_ORACLE_H
#define GAIS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* one HDLC frame-end event, in the order the reference would have hit them */
typedef struct goracle_frame {
	uint32_t end_bit;      /* index (0-based) of the NRZI-decoded bit that closed the frame */
	int16_t  nbits;        /* bufferpos - 22 (may be <= 0 for size failures) */
	uint8_t  status;       /* 0 = CRC ok, 1 = CRC fail, 2 = size/stop-bit fail */
	uint8_t  nbytes;       /* nbits / 8 when status == 0 else 0 */
	uint8_t  payload[56];  /* bytes[0..nbytes): LSB-first packing of the stored bits */
} goracle_frame;

/*
 * Same contract as gref_run() in ref_harness.c, plus an optional frame-event list.
 * stats[8] = { ok, crcfail, sizefail, pll, prev, lastbit, fsm state (1..5), seqnr }
 */
int goracle_run(const int16_t *buf, int64_t n_frames, int num_ch, int ch_ofs, int chunk,
		uint8_t *bits, int64_t bits_cap, int64_t *n_bits,
		uint8_t *signs,
		char *nmea, int64_t nmea_cap, int64_t *nmea_len,
		int32_t *stats,
		goracle_frame *frames, int64_t frames_cap, int64_t *n_frames_out);

/* planar [n_channels][n_samples]; returns seconds of the decode loop; see gref_bench() */
double goracle_bench(const int16_t *buf, int64_t n_channels, int64_t n_samples, int n_threads, int chunk,
		     int64_t *ok_total);

/* NMEA armouring of one CRC-ok frame (protodec_getdata + protodec_generate_nmea).
 * Writes 0, 1 or 2 "!AIVDM...\r\n" lines to out (cap >= 200), returns bytes written and
 * advances *seqnr exactly as the reference does. */
int goracle_nmea(const uint8_t *payload, int nbits, uint8_t *seqnr, char *out);

/* CRC-16/X.25 as protodec_sdlc_crc(): returns ~crc; a good frame+FCS gives 0x0f47 */
uint16_t goracle_crc16(const uint8_t *data, unsigned len);
/* the HDLC bit machine alone, fed with NRZI-decoded bits (one per byte) from the reset state */
int goracle_fsm_bits(const uint8_t *bits, int64_t n_bits, int32_t stats[3], goracle_frame *frames, int64_t frames_cap,
		     int64_t *n_frames_out);

/* float32 bit patterns of the 36 taps as the reference's compiler rounds them */
const uint32_t *goracle_tap_bits(void);

#ifdef __cplusplus
}
#endif
#endif

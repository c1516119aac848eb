/*
 * gais_state.h -- device-resident per-channel carried state and shared constants.
 *
 * One ChanState per channel holds everything struct receiver / struct filter /
 * struct demod_state_t carry between receiver_run() calls in the reference
 * (src/receiver.h:35-46, src/filter.h:57-62, src/protodec.h:44-71): the last 36 input
 * samples, the DPLL phase and slicer memory, the HDLC FSM registers, the partially
 * received frame, the message sequence number and the frame counters.
 */
#ifndef GAIS_STATE_H
#define GAIS_STATE_H

#include <stdint.h>

#define GAIS_NTAPS 36
#define GAIS_PLL_INC 13107u      /* 0x10000 / 5        src/receiver.c:69  */
#define GAIS_PLL_NUDGE 819u      /* pllinc / INC(16)   src/receiver.c:84,115-118 */
#define GAIS_STORE_WORDS 15      /* 449 stored bits max (src/protodec.c:1022) */

enum { GAIS_ST_HUNT = 1, GAIS_ST_PREAMBLE = 2, GAIS_ST_STARTFLAG = 3, GAIS_ST_DATA = 4, GAIS_ST_STOPFLAG = 5 };

struct ChanState {
	/* DPLL / slicer (src/receiver.h:35-46) */
	uint32_t pll;
	uint8_t prev, lastbit;
	/* HDLC FSM (src/protodec.h:44-71): state id of gais_track.cuh (it folds state, nstartsign, antallpreamble,
	 * antallenner, bitstuff and last), bufferpos, seqnr */
	uint8_t fsm, seqnr;
	uint16_t pos;
	uint32_t n_bits;                     /* NRZI bits produced since create/reset, modulo 2^32 (5.2 days of audio) */
	uint32_t dacc;                       /* NRZI difference bits sliced but not yet given to the FSM */
	uint32_t cur, cur2;                  /* the last 64 stored frame bits (newest at bit 31 of cur) */
	uint8_t nd, pad_[3];                 /* number of valid bits in dacc (< 24) */
	uint32_t store[GAIS_STORE_WORDS];    /* stored frame bits, LSB-first */
	int32_t ok, crcfail, sizefail;
	/* FIR history: x[n-36..n-1] of the NEXT run, double-buffered by run parity */
	int16_t hist[2][GAIS_NTAPS];
};

/* float32 bit patterns of the taps, i = 0..17, taps[35-i] = taps[i] (src/receiver.c:39-49 as
 * the compiler rounds the double literals; taps 0/1 are +0.0f, tap 2 is a denormal) */
#define GAIS_TAP_BITS_HALF { \
	0x00000000u, 0x00000000u, 0x00000069u, 0x0130bd6du, 0x0982c347u, 0x112a6907u, \
	0x18439833u, 0x1ec5b74eu, 0x24b00698u, 0x2a0a0629u, 0x2ebea222u, 0x32e7e4d5u, \
	0x36786fe0u, 0x396a68bfu, 0x3bc2cc99u, 0x3d8e92d5u, 0x3eb7cd8au, 0x3f50b242u }

#endif

/*
 * tests/c/layout_check.c -- include/gais_compat.h against the reference's OWN headers: the structs the host program
 * reaches into (src/ais.c:296-310, src/range.c:47-53) must have the same size and field offsets, and the seven entry
 * points the same prototypes (a mismatch is a compile error: both declarations of a function are visible).  The
 * reference's struct tags are renamed while its headers are read; the functions are not.  Compiled, never run, by
 * tests/test_abi.py when /root/reference is present.
 */
#include <stddef.h>

#define receiver ref_receiver
#define demod_state_t ref_demod_state_t
#include "receiver.h"          /* /root/reference/src, pulls in protodec.h, serial.h, ipc.h */
#undef receiver
#undef demod_state_t

/* the reference's prototypes, restated over the RENAMED tags, now meet the shim's over the real ones: C allows that
 * only for compatible types, and struct types with different tags are never compatible -- so the function
 * prototypes are checked through their parameter lists below instead */
#define init_receiver shim_init_receiver
#define free_receiver shim_free_receiver
#define receiver_run shim_receiver_run
#define protodec_initialize shim_protodec_initialize
#define protodec_reset shim_protodec_reset
#define protodec_getdata shim_protodec_getdata
#define protodec_decode shim_protodec_decode
#include "gais_compat.h"
#undef init_receiver
#undef free_receiver
#undef receiver_run
#undef protodec_initialize
#undef protodec_reset
#undef protodec_getdata
#undef protodec_decode

#define SAME(type_a, type_b, field) \
	_Static_assert(offsetof(struct type_a, field) == offsetof(struct type_b, field) && \
		       sizeof(((struct type_a *) 0)->field) == sizeof(((struct type_b *) 0)->field), "layout differs: " #field)

_Static_assert(sizeof(struct ref_receiver) == sizeof(struct receiver), "struct receiver size");
SAME(ref_receiver, receiver, filter);
SAME(ref_receiver, receiver, name);
SAME(ref_receiver, receiver, lastbit);
SAME(ref_receiver, receiver, num_ch);
SAME(ref_receiver, receiver, ch_ofs);
SAME(ref_receiver, receiver, pll);
SAME(ref_receiver, receiver, pllinc);
SAME(ref_receiver, receiver, decoder);
SAME(ref_receiver, receiver, prev);
SAME(ref_receiver, receiver, last_levellog);

_Static_assert(sizeof(struct ref_demod_state_t) == sizeof(struct demod_state_t), "struct demod_state_t size");
SAME(ref_demod_state_t, demod_state_t, chanid);
SAME(ref_demod_state_t, demod_state_t, state);
SAME(ref_demod_state_t, demod_state_t, offset);
SAME(ref_demod_state_t, demod_state_t, nskurr);
SAME(ref_demod_state_t, demod_state_t, npreamble);
SAME(ref_demod_state_t, demod_state_t, nstartsign);
SAME(ref_demod_state_t, demod_state_t, ndata);
SAME(ref_demod_state_t, demod_state_t, nstopsign);
SAME(ref_demod_state_t, demod_state_t, antallenner);
SAME(ref_demod_state_t, demod_state_t, buffer);
SAME(ref_demod_state_t, demod_state_t, rbuffer);
SAME(ref_demod_state_t, demod_state_t, tbuffer);
SAME(ref_demod_state_t, demod_state_t, bufferpos);
SAME(ref_demod_state_t, demod_state_t, last);
SAME(ref_demod_state_t, demod_state_t, antallpreamble);
SAME(ref_demod_state_t, demod_state_t, bitstuff);
SAME(ref_demod_state_t, demod_state_t, receivedframes);
SAME(ref_demod_state_t, demod_state_t, lostframes);
SAME(ref_demod_state_t, demod_state_t, lostframes2);
SAME(ref_demod_state_t, demod_state_t, seqnr);
SAME(ref_demod_state_t, demod_state_t, best_range);
SAME(ref_demod_state_t, demod_state_t, serial);
SAME(ref_demod_state_t, demod_state_t, ipc);
SAME(ref_demod_state_t, demod_state_t, serbuffer);
SAME(ref_demod_state_t, demod_state_t, ipcbuffer);
SAME(ref_demod_state_t, demod_state_t, nmea);

/* prototypes: a pointer to each shim function converts to a pointer to the reference's function type once the struct
 * tags are mapped -- spelled out with the reference's parameter lists (src/receiver.h:48-51, src/protodec.h:73-76) */
static struct receiver *(*const p1)(char, int, int, struct serial_state_t *, struct ipc_state_t *) = shim_init_receiver;
static void (*const p2)(struct receiver *) = shim_free_receiver;
static void (*const p3)(struct receiver *, short *, int) = shim_receiver_run;
static void (*const p4)(struct demod_state_t *, struct serial_state_t *, struct ipc_state_t *, char) = shim_protodec_initialize;
static void (*const p5)(struct demod_state_t *) = shim_protodec_reset;
static void (*const p6)(int, struct demod_state_t *) = shim_protodec_getdata;
static void (*const p7)(char *, int, struct demod_state_t *) = shim_protodec_decode;
const void *const layout_check_refs[] = { (const void *) &p1, (const void *) &p2, (const void *) &p3, (const void *) &p4, (const void *) &p5,
					   (const void *) &p6, (const void *) &p7 };

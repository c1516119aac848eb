/*
 * gais_b200.h -- C-ABI of libgaisb200.so: the B200-native batched AIS receive path.
 *
 * Plain C, plain pointers and sizes; no torch / CUDA types in any signature (streams are
 * passed as void*, i.e. a cudaStream_t / CUstream value, NULL = default stream).
 *
 * What it replaces in rubund/gnuais (citations are /root/reference paths):
 *
 *   gais_create / gais_destroy / gais_reset
 *        init_receiver() / free_receiver()            src/receiver.c:52-82, src/receiver.h:48-49
 *        filter_init(), protodec_initialize()         src/filter.c:57-71, src/protodec.c:54-76
 *        -- but for n_channels independent receivers at once.
 *   gais_run_device / gais_run_host
 *        receiver_run()                               src/receiver.c:87-135, src/receiver.h:51
 *        -> filter_run_buf()                          src/filter.c:106-143
 *        -> protodec_decode()                         src/protodec.c:988-1122
 *        -> protodec_calculate_crc()/protodec_sdlc_crc()  src/protodec.c:106-167
 *        One call = one audio chunk for EVERY channel; all DSP/FSM state carries over to the
 *        next call exactly like struct receiver / struct demod_state_t do.
 *   gais_get_messages / gais_msg
 *        the CRC-ok frames the reference hands to protodec_getdata()   src/protodec.c:1100-1104
 *   gais_get_counters
 *        demod_state_t.receivedframes / lostframes / lostframes2       src/protodec.h:58-60
 *   gais_nmea_format / gais_get_nmea
 *        protodec_getdata() (type gate, fill bits, seqnr) + protodec_generate_nmea()
 *                                                     src/protodec.c:896-929, :780-894
 *
 * The link-compatible legacy entry points (init_receiver/receiver_run/free_receiver with the
 * reference's own struct layouts) live in gais_compat.h / libgnuais_rx_b200.so.
 *
 * All functions return 0 on success and a negative GAIS_E* code on failure;
 * gais_last_error() gives a thread-local human-readable message.  There is NO CPU fallback:
 * without a usable CUDA device gais_create() fails with GAIS_ENODEV.
 */
#ifndef GAIS_B200_H
#define GAIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAIS_ABI_VERSION 1

enum {
	GAIS_OK = 0,
	GAIS_EINVAL = -1,    /* bad argument */
	GAIS_ENODEV = -2,    /* no CUDA device / wrong architecture */
	GAIS_ECUDA = -3,     /* a CUDA call failed (see gais_last_error) */
	GAIS_ENOMEM = -4,
	GAIS_EOVERFLOW = -5, /* a per-run capacity (messages/bits per channel) was exceeded */
};

enum {
	GAIS_LAYOUT_PLANAR = 0,      /* sample (c, n) at base[c * stride + n]  ([channel][time] rows) */
	GAIS_LAYOUT_INTERLEAVED = 1, /* sample (c, n) at base[n * stride + c]  (receiver_run()'s buf:
	                                frame-interleaved, stride = num_ch, c = ch_ofs; src/receiver.c:102) */
};

enum {
	GAIS_FIR_GUARD = 0, /* guard-banded fast FIR, exact fallback where the sign is in doubt (default) */
	GAIS_FIR_EXACT = 1, /* always the sequential 36-tap float32 mul+add chain of src/filter.h:40-49 */
};

enum {
	GAIS_KEEP_BITS = 1u << 0, /* materialise the NRZI-decoded bitstream fed to the HDLC stage */
	GAIS_KEEP_SIGNS = 1u << 1,/* keep the FIR sign words of the last run (parity artefact) */
	GAIS_KEEP_PEAK = 1u << 2, /* per-channel level of the last run: filter_run_buf()'s return value (src/filter.c:112-119) */
};

typedef struct gais_config {
	int32_t abi_version;        /* GAIS_ABI_VERSION */
	int32_t device;             /* CUDA device ordinal */
	int32_t n_channels;         /* independent receivers in the batch */
	int32_t layout;             /* GAIS_LAYOUT_* */
	int64_t max_frames_per_run; /* upper bound on n_frames of one gais_run_* call */
	int32_t fir_mode;           /* GAIS_FIR_* */
	uint32_t flags;             /* GAIS_KEEP_* */
	int32_t reserved[8];        /* tuning knobs, 0 = library default:
	                               [0] message slots per channel per run, [1] time-tile length in frames,
	                               [2] FIR/tracking overlap across tiles (1 = off, 2 = on),
	                               [3] number of this context's channel 0 in a batch sharded over several contexts /
	                                   GPUs: gais_msg.channel = reserved[3] + local index;
	                               [4] kernel chain: 1 = FIR-sign kernel + tracking kernel, 2 = the fused kernel wherever the
	                                   input allows it (planar, 16-byte aligned rows, GAIS_FIR_GUARD), 0 = the fused kernel
	                                   for batches of at least five 32-channel sets per SM (or what GAIS_FUSED says);
	                               [5..7] must be 0 */
} gais_config;

/* One CRC-ok HDLC frame, 64 bytes.  payload[j] is byte j of the frame as the reference packs
 * it for the CRC (stored bit 8j+i -> bit i, src/protodec.c:138-143); AIS payload bit k is
 * (payload[k/8] >> (7 - k%8)) & 1 (src/protodec.c:151-161). */
typedef struct gais_msg {
	uint8_t  payload[53]; /* nbits/8 bytes valid, rest 0 (word-aligned first so the GPU stores it as words) */
	uint8_t  flags;       /* bits 0-3: seqnr the reference holds when it formats this frame (0..9);
	                         bit 4: type gate passed (1 <= type <= 24 -> NMEA is emitted) */
	uint16_t nbits;       /* bufferpos - 22 of src/protodec.c:1096 (payload bits incl. a ragged tail) */
	uint32_t channel;     /* channel index inside the batch (+ gais_config.reserved[3] when the batch is sharded) */
	uint32_t end_bit;     /* index of the NRZI bit that closed the frame, counted per channel since create/reset, modulo 2^32
	                         (about 5.2 days of audio at 9600 bit/s: monotonic within a run, compare across runs with wrap-safe arithmetic) */
} gais_msg;

/* per-channel frame counters: receivedframes, lostframes (CRC), lostframes2 (size/stop bit) */
typedef struct gais_counters {
	int32_t ok, crcfail, sizefail;
} gais_counters;

/* per-channel carried DSP/FSM state, exposed for parity tests (src/receiver.h:35-46,
 * src/protodec.h:44-71) */
typedef struct gais_chan_state {
	uint32_t pll;
	int32_t prev, lastbit;
	int32_t fsm_state;   /* ST_SKURR=1 .. ST_STOPSIGN=5 (src/protodec.h:30-34) */
	int32_t seqnr;
	uint32_t n_bits;     /* NRZI bits produced since create/reset, modulo 2^32 */
} gais_chan_state;

/* fixed-stride NMEA text produced on the GPU: up to two "!AIVDM...\r\n" sentences per message */
#define GAIS_NMEA_STRIDE 176
typedef struct gais_nmea_rec {
	uint8_t len;                       /* bytes used in text (0 if the type gate dropped it) */
	char text[GAIS_NMEA_STRIDE - 1];
} gais_nmea_rec;

/* timing of the last gais_run_* call, measured with CUDA events on the run's stream(s) */
typedef struct gais_timing {
	float total_ms;      /* first kernel start -> last kernel end (device time) */
	float fir_ms;        /* FIR-sign kernel(s), summed over time tiles */
	float track_ms;      /* DPLL/slicer/NRZI/HDLC/CRC kernel(s), summed */
	float post_ms;       /* frame check + compaction */
	int32_t launches;    /* kernels launched by the call */
	float nmea_ms;       /* the last gais_device_nmea() / gais_get_nmea_text() armouring pass (0 if none) */
} gais_timing;

typedef struct gais_ctx gais_ctx;

const char *gais_last_error(void);
int gais_abi_version(void);
/* number of usable sm_100 devices; 0 when there is none (never falls back to the CPU) */
int gais_device_count(void);

int gais_create(const gais_config *cfg, gais_ctx **out);
void gais_destroy(gais_ctx *ctx);
/* back to the state right after init_receiver(): zero history, pll = 0, FSM reset, counters 0 */
int gais_reset(gais_ctx *ctx);
/* protodec_reset() for every channel (src/protodec.c:87-100): the HDLC bit machine goes back to ST_SKURR with an
 * empty frame buffer; counters, seqnr, DPLL and filter history stay */
int gais_reset_fsm(gais_ctx *ctx);

/*
 * One chunk for all channels, samples already in device memory.  Asynchronous on `stream`;
 * results of this run are readable after gais_sync() or any gais_get_*() (which sync).
 * n_frames = samples per channel in this chunk (<= max_frames_per_run; no 4096 limit --
 * src/receiver.c:104-105 aborts above FILTERED_LEN, the result is chunk-size invariant).
 */
int gais_run_device(gais_ctx *ctx, const int16_t *d_samples, int64_t n_frames, int64_t stride, void *stream);
/* Same from HOST memory (pinned or pageable): H2D copies are pipelined with the kernels. */
int gais_run_host(gais_ctx *ctx, const int16_t *h_samples, int64_t n_frames, int64_t stride);
int gais_sync(gais_ctx *ctx);

/*
 * The protocol half alone -- protodec_decode() (src/protodec.c:988-1122, declared src/protodec.h:76) for every
 * channel: n_bits NRZI-decoded bits per channel, ONE PER BYTE (0/1) as receiver_run() hands them over
 * (src/receiver.c:126-131), channel c at bits[c * stride + i].  HDLC flag hunt, de-stuffing, CRC-16, counters,
 * message records and NMEA exactly as after gais_run_*(); the FSM state carries over between calls and may be
 * mixed with gais_run_*() calls on the same context (the DPLL never holds bits back across a call).
 * n_bits <= max_frames_per_run.
 */
int gais_run_bits_device(gais_ctx *ctx, const uint8_t *d_bits, int64_t n_bits, int64_t stride, void *stream);
int gais_run_bits_host(gais_ctx *ctx, const uint8_t *h_bits, int64_t n_bits, int64_t stride);

/* messages of the LAST run, dense, ordered by (channel, end_bit) */
int gais_message_count(gais_ctx *ctx, int64_t *n_msgs);
int gais_get_messages(gais_ctx *ctx, gais_msg *h_out, int64_t cap, int64_t *n_msgs);
/* Page-locked host memory for the buffers handed to gais_run_host() / gais_get_messages(): copies to
 * and from it run at PCIe speed without a staging pass (the reference reads its audio into a plain
 * malloc'd buffer, src/ais.c:176-182; this is the batched equivalent of that allocation). */
int gais_host_alloc(void **h_ptr, size_t bytes);
void gais_host_free(void *h_ptr);
/* device-resident view of the same array (for NCCL gathers / zero-copy consumers) */
int gais_device_messages(gais_ctx *ctx, const gais_msg **d_msgs, int64_t *n_msgs);
/* NMEA text of the last run's messages, armoured on the GPU; record i belongs to message i */
int gais_get_nmea(gais_ctx *ctx, gais_nmea_rec *h_out, int64_t cap, int64_t *n_msgs);

/* The same text packed back to back, armoured by a warp per message on the run's stream: message i is
 * d_text[d_offsets[i] .. d_offsets[i+1]) (nothing for a message the type gate dropped), n_bytes = d_offsets[n_msgs].
 * The buffers belong to the context and are reused between runs (they grow only when a run produces more messages
 * or text than any before it).  Bytes are those of protodec_generate_nmea() in its serial form "!%s\r\n"
 * (src/protodec.c:780-894, :883). */
int gais_device_nmea(gais_ctx *ctx, const char **d_text, const uint64_t **d_offsets, int64_t *n_msgs, int64_t *n_bytes);
/* ... and copied to the host: all sentences of the last run in (channel, end_bit) order */
int gais_get_nmea_text(gais_ctx *ctx, char *h_text, int64_t cap, int64_t *n_bytes);

/* cumulative since create/reset, h_out[n_channels] */
int gais_get_counters(gais_ctx *ctx, gais_counters *h_out);
int gais_get_state(gais_ctx *ctx, gais_chan_state *h_out);
/* sum over channels, without copying the per-channel arrays */
int gais_get_totals(gais_ctx *ctx, int64_t totals[3]);

/* GAIS_KEEP_BITS: bits of the LAST run, LSB-first packed; row c starts at h_words + c*words_per_row;
 * h_nbits[c] = bits produced by channel c in the last run */
int gais_bits_row_words(gais_ctx *ctx, int64_t *words_per_row);
int gais_get_bits(gais_ctx *ctx, uint32_t *h_words, uint32_t *h_nbits);
/* GAIS_KEEP_SIGNS: FIR signs of the LAST run, bit j of word w of channel c = (filtered[32w+j] > 0);
 * layout [word][channel]: h_words[w * n_channels + c] */
int gais_get_signs(gais_ctx *ctx, uint32_t *h_words, int64_t cap_words);

/* GAIS_KEEP_PEAK: h_out[n_channels] = the level filter_run_buf() would have returned for the LAST run had it been
 * one call -- max(0, largest sample): positive samples only, as the reference (src/filter.c:112-119; for a run
 * fed in several receiver_run() chunks it is the maximum of the per-chunk values).  receiver_run() logs it as
 * maxval / 32768 * 100 percent (src/receiver.c:137-147). */
int gais_get_peaks(gais_ctx *ctx, int16_t *h_out);

int gais_get_timing(gais_ctx *ctx, gais_timing *out);

/* Host-side NMEA armouring of one record (same bytes as the GPU path; used by the legacy
 * shim).  out must hold GAIS_NMEA_STRIDE bytes; returns the text length (0 if gated). */
int gais_nmea_format(const gais_msg *msg, char *out);

/* The text line gnuais prints to stdout for a message:
 *   "ch %c type %d mmsi %09ld:<fields> (!AIVDM,...)\n"
 * protodec_getdata() src/protodec.c:931-985 + the field decoders src/protodec.c:216-776 (stdout text
 * only; the MySQL/cache/range sinks are not fed).  Host-side, per message.  Returns the length
 * written (0 when the type gate drops the message), out needs >= 512 bytes. */
int gais_text_format(const gais_msg *msg, char chanid, char *out, int cap);

/* ---- synthetic workload; integer-only, host and device agree bit-for-bit -- */

typedef struct gais_synth {
	uint64_t seed;
	int32_t amplitude;  /* 12000 */
	int32_t noise_q16;  /* round(sigma * 65536 / 37837.2): sigma 300 -> 520, 1500 -> 2598 */
	int32_t rho_q16;    /* P(slot carries a burst) * 65536 */
	int32_t jitter;     /* 1: burst start jitters 0..4 samples */
} gais_synth;

/* channels [first_channel, first_channel + n_channels), samples [0, n_frames) each, written
 * with the given layout/stride (element (c, n) as in GAIS_LAYOUT_*; c relative to first_channel) */
int gais_synth_host(const gais_synth *p, uint32_t first_channel, int32_t n_channels, int64_t n_frames,
		    int16_t *h_out, int32_t layout, int64_t stride);
int gais_synth_device(const gais_synth *p, uint32_t first_channel, int32_t n_channels, int64_t n_frames,
		      int16_t *d_out, int32_t layout, int64_t stride, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GAIS_B200_H */

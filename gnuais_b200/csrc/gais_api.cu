/*
 * gais_api.cu -- the C-ABI of include/gais_b200.h over the kernels in gais_kernels.cuh.
 *
 * A run (one gais_run_* call = one receiver_run() chunk for every channel, src/receiver.c:87)
 * is cut into TIME TILES; all DSP/FSM state is carried in ChanState between tiles exactly as
 * it is carried between runs, so the tiling is invisible in the results (the reference's
 * output is chunk-size invariant).  Per tile: FIR-sign kernel -> sign words in
 * a (normally L2-resident) [word][channel] buffer -> tracking kernel (DPLL .. CRC) appending
 * 64-byte records to per-channel slots.  After the last tile: scan + gather into the dense
 * (channel, end_bit)-ordered message array.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <new>

#include "gais_b200.h"
#include "gais_kernels.cuh"
#include "gais_fir.cuh"
#include "gais_fir_umma.cuh"
#include "gais_fir_tc.cuh"
#include "gais_track.cuh"
#include "gais_fused.cuh"

using namespace gais;

static thread_local char tl_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(tl_err, sizeof(tl_err), fmt, ap);
	va_end(ap);
	return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return fail(GAIS_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" const char *gais_last_error(void) { return tl_err; }
extern "C" int gais_abi_version(void) { return GAIS_ABI_VERSION; }

extern "C" int gais_device_count(void)
{
	int n = 0, ok = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	for (int i = 0; i < n; i++) {
		cudaDeviceProp p;
		if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10)
			ok++;
	}
	return ok;
}

enum { EV_START = 0, EV_END, EV_POST0, EV_POST1, EV_COUNT };

struct gais_ctx {
	gais_config cfg;
	int n_ch;
	int64_t tile_frames;        /* multiple of 32 */
	int hist_sel;
	ChanState *d_state;
	uint32_t *d_signs[2];       /* [tile words][n_ch] double buffer, or one run-sized buffer (KEEP_SIGNS) */
	int64_t sign_words;         /* words per channel in each buffer */
	gais_msg *d_slots;
	int slot_cap;
	uint32_t *d_run_count, *d_run_bits;
	uint64_t *d_offsets;
	gais_msg *d_dense;
	int64_t dense_cap;
	gais_nmea_rec *d_nmea;
	int64_t nmea_cap;
	/* packed NMEA text (gais_device_nmea) */
	char *d_text;
	int64_t text_cap;
	uint64_t *d_text_off, *d_blk_off;
	uint32_t *d_blk_tot;
	int64_t text_msgs_cap;
	int64_t text_bytes;
	int text_valid;
	cudaEvent_t ev_nmea[2];
	uint32_t *d_bits;
	int bits_row_words;
	int16_t *d_peak;            /* GAIS_KEEP_PEAK */
	int32_t *d_overflow;
	unsigned long long *d_totals;
	int16_t *d_stage[2];        /* gais_run_host staging tiles */
	int16_t *d_planar;          /* planar copy of an interleaved tile (batches of >= 32 interleaved channels) */
	int64_t planar_elems, planar_stride;
	int64_t stage_elems;
	cudaStream_t s_copy, s_own, s_fir, s_trk;
	cudaEvent_t ev_fir_done[2], ev_trk_done[2], ev_join;
	int overlap;                /* GAIS_OVERLAP=1: FIR of tile t+1 concurrent with tracking of tile t (two-kernel path only) */
	int fused;                  /* 0: two kernels; 1 (default): the one-kernel chain of gais_fused.cuh for batches that fill the GPU; 2: wherever the input allows it */
	int n_sms;
	cudaEvent_t ev[EV_COUNT], ev_copy[2], ev_free[2];
	cudaEvent_t *ev_tile;       /* 4 per tile: fir start, fir end, track start, track end */
	int ev_tile_cap;
	cudaStream_t last_stream;
	int pending;                /* a run has been enqueued and not finished */
	int finished;               /* dense array of the last run is valid */
	int n_tiles_last;
	int64_t last_frames;
	int64_t n_msgs;
	int launches;
	gais_timing timing;
};

static int64_t env_i64(const char *name, int64_t dflt)
{
	const char *s = getenv(name);
	if (!s || !*s)
		return dflt;
	return strtoll(s, NULL, 10);
}

static int upload_taps(void)
{
	static const uint32_t half[18] = GAIS_TAP_BITS_HALF;
	uint32_t full[GAIS_NTAPS];
	for (int i = 0; i < GAIS_NTAPS; i++)
		full[i] = half[i < 18 ? i : 35 - i];
	CK(cudaMemcpyToSymbol(c_taps, full, sizeof(full)));
	return 0;
}

extern "C" int gais_reset(gais_ctx *ctx)
{
	if (!ctx)
		return fail(GAIS_EINVAL, "null ctx");
	CK(cudaSetDevice(ctx->cfg.device));
	CK(cudaDeviceSynchronize());
	ChanState init;
	memset(&init, 0, sizeof(init));
	init.fsm = (uint8_t) hdlc_hunt_id(0, 0, 0);   /* protodec_reset(); everything else is 0 (src/protodec.c:54-100, src/receiver.c:52-74) */
	/* replicate by doubling copies */
	CK(cudaMemcpy(ctx->d_state, &init, sizeof(init), cudaMemcpyHostToDevice));
	for (int64_t have = 1; have < ctx->n_ch; have *= 2) {
		int64_t n = (have * 2 <= ctx->n_ch) ? have : ctx->n_ch - have;
		CK(cudaMemcpy(ctx->d_state + have, ctx->d_state, (size_t) n * sizeof(ChanState), cudaMemcpyDeviceToDevice));
	}
	CK(cudaMemset(ctx->d_run_count, 0, sizeof(uint32_t) * ctx->n_ch));
	CK(cudaMemset(ctx->d_run_bits, 0, sizeof(uint32_t) * ctx->n_ch));
	CK(cudaMemset(ctx->d_overflow, 0, sizeof(int32_t)));
	ctx->hist_sel = 0;
	ctx->pending = 0;
	ctx->finished = 0;
	ctx->n_msgs = 0;
	return 0;
}

__global__ void reset_fsm_kernel(ChanState *st, int n)
{
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c < n) {
		st[c].fsm = (uint8_t) hdlc_hunt_id(0, 0, 0);      /* src/protodec.c:87-100 */
		st[c].pos = 0;
	}
}

static int finish(gais_ctx *ctx);

extern "C" int gais_reset_fsm(gais_ctx *ctx)
{
	if (!ctx)
		return fail(GAIS_EINVAL, "null ctx");
	int rc = finish(ctx);
	if (rc && rc != GAIS_EOVERFLOW)
		return rc;
	reset_fsm_kernel<<<(ctx->n_ch + 127) / 128, 128>>>(ctx->d_state, ctx->n_ch);
	CK(cudaGetLastError());
	CK(cudaDeviceSynchronize());
	return 0;
}

extern "C" void gais_destroy(gais_ctx *ctx)
{
	if (!ctx)
		return;
	cudaSetDevice(ctx->cfg.device);
	cudaDeviceSynchronize();
	cudaFree(ctx->d_state);
	cudaFree(ctx->d_signs[0]);
	cudaFree(ctx->d_signs[1]);
	cudaFree(ctx->d_slots);
	cudaFree(ctx->d_run_count);
	cudaFree(ctx->d_run_bits);
	cudaFree(ctx->d_offsets);
	cudaFree(ctx->d_dense);
	cudaFree(ctx->d_nmea);
	cudaFree(ctx->d_text); cudaFree(ctx->d_text_off); cudaFree(ctx->d_blk_off); cudaFree(ctx->d_blk_tot);
	for (int i = 0; i < 2; i++)
		if (ctx->ev_nmea[i]) cudaEventDestroy(ctx->ev_nmea[i]);
	cudaFree(ctx->d_bits);
	cudaFree(ctx->d_peak);
	cudaFree(ctx->d_overflow);
	cudaFree(ctx->d_totals);
	cudaFree(ctx->d_stage[0]);
	cudaFree(ctx->d_stage[1]);
	cudaFree(ctx->d_planar);
	if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
	if (ctx->s_own) cudaStreamDestroy(ctx->s_own);
	if (ctx->s_fir) cudaStreamDestroy(ctx->s_fir);
	if (ctx->s_trk) cudaStreamDestroy(ctx->s_trk);
	for (int i = 0; i < 2; i++) {
		if (ctx->ev_fir_done[i]) cudaEventDestroy(ctx->ev_fir_done[i]);
		if (ctx->ev_trk_done[i]) cudaEventDestroy(ctx->ev_trk_done[i]);
	}
	if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
	for (int i = 0; i < EV_COUNT; i++)
		if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
	for (int i = 0; i < 2; i++) {
		if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
		if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
	}
	for (int i = 0; i < ctx->ev_tile_cap; i++)
		cudaEventDestroy(ctx->ev_tile[i]);
	free(ctx->ev_tile);
	delete ctx;
}

extern "C" int gais_create(const gais_config *cfg, gais_ctx **out)
{
	if (!cfg || !out)
		return fail(GAIS_EINVAL, "null argument");
	if (cfg->abi_version != GAIS_ABI_VERSION)
		return fail(GAIS_EINVAL, "abi_version %d != %d", cfg->abi_version, GAIS_ABI_VERSION);
	if (cfg->n_channels < 1 || cfg->max_frames_per_run < 1)
		return fail(GAIS_EINVAL, "n_channels and max_frames_per_run must be >= 1");
	if (cfg->layout != GAIS_LAYOUT_PLANAR && cfg->layout != GAIS_LAYOUT_INTERLEAVED)
		return fail(GAIS_EINVAL, "unknown layout %d", cfg->layout);
	if (cfg->fir_mode != GAIS_FIR_GUARD && cfg->fir_mode != GAIS_FIR_EXACT)
		return fail(GAIS_EINVAL, "unknown fir_mode %d", cfg->fir_mode);

	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
		cudaGetLastError();
		return fail(GAIS_ENODEV, "no CUDA device visible: libgaisb200 has no CPU fallback");
	}
	if (cfg->device < 0 || cfg->device >= ndev)
		return fail(GAIS_ENODEV, "device %d out of range (0..%d)", cfg->device, ndev - 1);
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, cfg->device));
	if (prop.major != 10)
		return fail(GAIS_ENODEV, "device %d is sm_%d%d; this library carries sm_100a code only", cfg->device,
			    prop.major, prop.minor);
	CK(cudaSetDevice(cfg->device));

	gais_ctx *ctx = new (std::nothrow) gais_ctx();
	if (!ctx)
		return fail(GAIS_ENOMEM, "out of host memory");
	memset(ctx, 0, sizeof(*ctx));
	ctx->cfg = *cfg;
	ctx->n_ch = cfg->n_channels;

	int64_t tile = env_i64("GAIS_TILE_FRAMES", 0);
	if (cfg->reserved[1] > 0)
		tile = cfg->reserved[1];
	if (tile <= 0) {
		/* default: 384 MB of sign words per tile (n_ch * tile / 8 bytes).  Measured on B200
		 * (profiles/r1_sweep_tiles.txt, profiles/r1_experiments.txt): fewer, longer launches beat keeping
		 * the sign words inside the 126 MB L2 -- the extra 1/16 byte per sample of HBM traffic is cheaper
		 * than the per-launch ramps; with the final kernels 49152-sample tiles at 65536 channels are
		 * 1.8 % faster than 32768 and than 61440 */
		tile = (int64_t) 384 * 1024 * 1024 * 8 / ctx->n_ch;
		if (tile > 65536) tile = 65536;
		if (tile < 2048) tile = 2048;
	}
	tile = (tile + 1023) / 1024 * 1024;
	ctx->tile_frames = tile;

	int rc = 0;
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	rc = fail(e_ == cudaErrorMemoryAllocation ? GAIS_ENOMEM : GAIS_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); goto bad; } } while (0)

	{
		int64_t run_frames = cfg->max_frames_per_run;
		/* which chain: reserved[4] 1 = the two kernels, 2 = the fused kernel wherever the input allows it, 0 = GAIS_FUSED
		 * (same values; default 1... "auto": fused when the batch gives every SM at least five channel sets) */
		ctx->fused = cfg->reserved[4] == 1 ? 0 : cfg->reserved[4] == 2 ? 2 : (int) env_i64("GAIS_FUSED", 1);
		if (fir_impl() != 2)
			ctx->fused = 0;
		ctx->n_sms = prop.multiProcessorCount;
		if (cfg->flags & GAIS_KEEP_SIGNS) {
			ctx->sign_words = ((run_frames + tile - 1) / tile) * (tile / 32);
			CKC(cudaMalloc(&ctx->d_signs[0], (size_t) ctx->sign_words * ctx->n_ch * 4));
		}
		/* otherwise the sign-word buffers of the two-kernel path are allocated by the first run that needs them
		 * (ensure_signs): the fused kernel keeps the sign words in shared memory */
		ctx->slot_cap = cfg->reserved[0] > 0 ? cfg->reserved[0] : (int) (run_frames / 1280 + run_frames / 5120 + 8);
		CKC(cudaMalloc(&ctx->d_state, (size_t) ctx->n_ch * sizeof(ChanState)));
		CKC(cudaMalloc(&ctx->d_slots, (size_t) ctx->n_ch * ctx->slot_cap * sizeof(gais_msg)));
		CKC(cudaMalloc(&ctx->d_run_count, (size_t) ctx->n_ch * 4));
		CKC(cudaMalloc(&ctx->d_run_bits, (size_t) ctx->n_ch * 4));
		CKC(cudaMalloc(&ctx->d_offsets, (size_t) (ctx->n_ch + 1) * 8));
		CKC(cudaMalloc(&ctx->d_overflow, 4));
		CKC(cudaMalloc(&ctx->d_totals, 3 * 8));
		if (cfg->flags & GAIS_KEEP_BITS) {
			/* the DPLL advances at most (13107 + 819) / 65536 bit per sample */
			int64_t max_bits = run_frames * 13926 / 65536 + 2;
			ctx->bits_row_words = (int) ((max_bits + 31) / 32);
			CKC(cudaMalloc(&ctx->d_bits, (size_t) ctx->n_ch * ctx->bits_row_words * 4));
		}
		if (cfg->flags & GAIS_KEEP_PEAK)
			CKC(cudaMalloc(&ctx->d_peak, (size_t) ctx->n_ch * 2));
		CKC(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
		CKC(cudaStreamCreateWithFlags(&ctx->s_own, cudaStreamNonBlocking));
		{
			int lo = 0, hi = 0;
			CKC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
			CKC(cudaStreamCreateWithPriority(&ctx->s_fir, cudaStreamNonBlocking, lo));
			CKC(cudaStreamCreateWithPriority(&ctx->s_trk, cudaStreamNonBlocking, hi));
			for (int i = 0; i < 2; i++) {
				CKC(cudaEventCreateWithFlags(&ctx->ev_fir_done[i], cudaEventDisableTiming));
				CKC(cudaEventCreateWithFlags(&ctx->ev_trk_done[i], cudaEventDisableTiming));
			}
			CKC(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
			/* reserved[2]: 0 = default (on unless GAIS_OVERLAP=0), 1 = off, 2 = on */
			ctx->overlap = cfg->reserved[2] == 1 ? 0 : cfg->reserved[2] == 2 ? 1 : (int) env_i64("GAIS_OVERLAP", 1);
		}
		for (int i = 0; i < EV_COUNT; i++)
			CKC(cudaEventCreate(&ctx->ev[i]));
		for (int i = 0; i < 2; i++) {
			CKC(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
			CKC(cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming));
		}
	}
	if ((rc = upload_taps()) != 0)
		goto bad;
	if ((rc = fir_setup()) != 0 || (rc = fir_umma_setup()) != 0 || (rc = fir_tc_setup(cfg->device)) != 0 || (rc = fused_setup()) != 0) {
		rc = fail(GAIS_ECUDA, "fir_setup failed: %s", cudaGetErrorString(cudaGetLastError()));
		goto bad;
	}
	if ((rc = gais_reset(ctx)) != 0)
		goto bad;
	*out = ctx;
	return 0;
bad:
	gais_destroy(ctx);
	return rc;
#undef CKC
}

static int ensure_tile_events(gais_ctx *ctx, int n_tiles)
{
	int need = n_tiles * 4;
	if (need <= ctx->ev_tile_cap)
		return 0;
	cudaEvent_t *ne = (cudaEvent_t *) realloc(ctx->ev_tile, sizeof(cudaEvent_t) * need);
	if (!ne)
		return fail(GAIS_ENOMEM, "out of host memory");
	ctx->ev_tile = ne;
	for (int i = ctx->ev_tile_cap; i < need; i++) {
		CK(cudaEventCreate(&ctx->ev_tile[i]));
		ctx->ev_tile_cap = i + 1;
	}
	return 0;
}

static TrackOut make_out(gais_ctx *ctx)
{
	TrackOut out;
	out.slots = ctx->d_slots;
	out.run_count = ctx->d_run_count;
	out.bits = ctx->d_bits;
	out.run_bits = ctx->d_run_bits;
	out.slot_cap = ctx->slot_cap;
	out.bits_row_words = ctx->bits_row_words;
	out.overflow = ctx->d_overflow;
	return out;
}

/* after the last tile: frame check (CRC, counters, seqnr, compaction) and the count scan */
static int enqueue_post(gais_ctx *ctx, cudaStream_t st)
{
	int nl = finalize_launch(ctx->d_state, ctx->n_ch, make_out(ctx), st);
	if (nl < 0)
		return fail(GAIS_ECUDA, "frame-check launch failed: %s", cudaGetErrorString(cudaGetLastError()));
	ctx->launches += nl;
	scan_counts_kernel<<<1, 1024, 0, st>>>(ctx->d_run_count, ctx->n_ch, ctx->d_offsets);
	ctx->launches++;
	return 0;
}

/* interleaved batches of at least 32 channels are copied into planar rows per tile (deinterleave_kernel: one more pass over
 * the tile in HBM, 4 B per sample, against the 64 dependent FP32 operations per sample of the exact FIR kernel on strided loads);
 * the reference's own stereo shape (2 channels) stays on the generic kernels */
static bool relay_planar(const gais_ctx *ctx, const SampleView &view)
{
	return ctx->cfg.layout == GAIS_LAYOUT_INTERLEAVED && view.ch_stride == 1 && ctx->n_ch >= 32 && ctx->cfg.fir_mode == GAIS_FIR_GUARD;
}

static int ensure_planar(gais_ctx *ctx, int64_t tile_frames)
{
	const int64_t stride = (tile_frames + 7) / 8 * 8, need = stride * ctx->n_ch;
	if (need <= ctx->planar_elems) {
		return 0;
	}
	CK(cudaDeviceSynchronize());
	cudaFree(ctx->d_planar);
	ctx->d_planar = NULL;
	ctx->planar_elems = 0;
	cudaError_t e = cudaMalloc(&ctx->d_planar, (size_t) need * 2);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(e == cudaErrorMemoryAllocation ? GAIS_ENOMEM : GAIS_ECUDA, "planar copy of an interleaved tile (%lld samples x %d channels): %s",
			    (long long) stride, ctx->n_ch, cudaGetErrorString(e));
	}
	ctx->planar_elems = need;
	ctx->planar_stride = stride;
	return 0;
}

/* which part of a tile the fused kernel takes: channels [0, ch) x samples [0, frames) (0, 0: none) */
struct FusedPart { int ch; int64_t frames; };

static FusedPart fused_part(const gais_ctx *ctx, const SampleView &view, int64_t n_frames)
{
	FusedPart p = { 0, 0 };
	/* planar rows, or an interleaved batch that enqueue_tile() will re-lay as planar rows first */
	const bool aligned = (view.t_stride == 1 && (view.ch_stride % 8) == 0 && ((uintptr_t) view.base % 16) == 0) || relay_planar(ctx, view);
	/* small batches stay on the two kernels: with few channel sets per SM the chain is paced by tracker warps walking their
	 * blocks nearly alone either way, and the stand-alone tracker (79 registers, no ring hand-over) walks them faster
	 * (1024 channels x 480000: 10.4 vs 13.0 ms; 16384 channels: 4.00 vs 4.13 ms per 131072 samples; 32768: 5.23 vs 4.65 --
	 * the fused kernel wins from about 4.5 sets per SM on, profiles/r2_fused_experiments.txt) */
	const bool want = ctx->fused == 2 || (ctx->fused == 1 && ctx->n_ch / 32 >= 5 * ctx->n_sms);
	if (want && ctx->cfg.fir_mode == GAIS_FIR_GUARD && aligned && n_frames <= X_MAX_FRAMES) {
		p.ch = ctx->n_ch / 32 * 32;
		p.frames = n_frames / P_T * P_T;
		if (p.ch == 0 || p.frames == 0)
			p.ch = 0, p.frames = 0;
	}
	return p;
}

/* sign words per channel a tile needs in global memory: what the fused kernel does not take */
static int64_t tile_sign_words(const gais_ctx *ctx, const FusedPart &fp, int64_t n_frames)
{
	if (fp.ch == ctx->n_ch)
		return (n_frames - fp.frames + 31) / 32;
	return (n_frames + 31) / 32;
}

/* the two sign-word buffers of the two-kernel path, [words][n_ch] each, grown on demand (not with GAIS_KEEP_SIGNS:
 * that buffer is run-sized from the start) */
static int ensure_signs(gais_ctx *ctx, int64_t words)
{
	if ((ctx->cfg.flags & GAIS_KEEP_SIGNS) || words <= ctx->sign_words)
		return 0;
	CK(cudaDeviceSynchronize());
	cudaFree(ctx->d_signs[0]);
	cudaFree(ctx->d_signs[1]);
	ctx->d_signs[0] = ctx->d_signs[1] = NULL;
	ctx->sign_words = 0;
	for (int i = 0; i < 2; i++) {
		cudaError_t e = cudaMalloc(&ctx->d_signs[i], (size_t) words * ctx->n_ch * 4);
		if (e != cudaSuccess) {
			cudaGetLastError();
			return fail(e == cudaErrorMemoryAllocation ? GAIS_ENOMEM : GAIS_ECUDA, "sign-word buffer (%lld words x %d channels): %s",
				    (long long) words, ctx->n_ch, cudaGetErrorString(e));
		}
	}
	ctx->sign_words = words;
	return 0;
}

/* enqueue the chain for one time tile whose samples are at `view` (n = 0 is the first sample of the tile).
 * The fused kernel (gais_fused.cuh) takes whole channel sets and whole 256-sample stages of aligned planar
 * input; the FIR-sign and tracking kernels sweep up what is left (ragged end of the tile, channels beyond a
 * multiple of 32) and everything else (interleaved or unaligned input, GAIS_FIR_EXACT, small batches, GAIS_FUSED=0). */
static int enqueue_tile(gais_ctx *ctx, SampleView view, int64_t n_frames, int tile_idx, int64_t word_ofs, cudaStream_t st,
			cudaStream_t st_trk, bool timed)
{
	/* st carries the FIR stage, st_trk the tracking stage.  When they differ (overlap mode of the two-kernel
	 * path) the FIR of tile t+1 runs while tile t is being tracked; the two sign buffers are handed back and
	 * forth with events. */
	const bool overlap = st != st_trk;
	const bool keep = (ctx->cfg.flags & GAIS_KEEP_SIGNS) != 0;
	int layout = ctx->cfg.layout;
	if (relay_planar(ctx, view) && ctx->d_planar && n_frames <= ctx->planar_stride) {
		dim3 grid((unsigned) ((ctx->n_ch + 31) / 32), (unsigned) ((n_frames + 63) / 64));
		/* (every reader of the previous tile's planar copy -- FIR, history, peak -- is ahead of this launch on the same stream;
		 * the tracking stage reads sign words only) */
		deinterleave_kernel<<<grid, 256, 0, st>>>(view.base, view.t_stride, ctx->n_ch, n_frames, ctx->d_planar, ctx->planar_stride);
		ctx->launches++;
		view.base = ctx->d_planar;
		view.ch_stride = ctx->planar_stride;
		view.t_stride = 1;
		layout = GAIS_LAYOUT_PLANAR;
	}
	const FusedPart fp = overlap ? FusedPart{ 0, 0 } : fused_part(ctx, view, n_frames);
	uint32_t *signs;
	if (keep)
		signs = ctx->d_signs[0] + word_ofs * ctx->n_ch;
	else {
		/* when the fused kernel takes every channel, only the words after its last stage exist in global memory */
		signs = ctx->d_signs[tile_idx & 1];
		if (fp.ch == ctx->n_ch && signs)
			signs -= (fp.frames / 32) * ctx->n_ch;
	}
	TrackOut out = make_out(ctx);

	if (overlap && !keep && tile_idx >= 2)
		CK(cudaStreamWaitEvent(st, ctx->ev_trk_done[tile_idx & 1], 0));   /* buffer still being read by tile t-2 */
	if (timed) CK(cudaEventRecord(ctx->ev_tile[4 * tile_idx + 0], st));
	int hist_saved = 0, nl;
	if (fp.ch) {
		nl = fused_launch(view, ctx->d_state, ctx->hist_sel, ctx->n_ch, fp.ch, fp.frames, fp.frames == n_frames ? 1 : 0,
				  keep ? signs : NULL, out, ctx->n_sms, st);
		if (nl < 0)
			return fail(GAIS_ECUDA, "fused launch failed: %s", cudaGetErrorString(cudaGetLastError()));
		ctx->launches += nl;
		hist_saved = fp.frames == n_frames ? fp.ch : 0;
		/* FIR signs of what is left: the ragged end of the fused channels, and the other channels */
		if (fp.frames < n_frames) {
			dim3 grid((unsigned) ((fp.ch + K1_CH - 1) / K1_CH), (unsigned) ((n_frames - fp.frames + K1_TILE - 1) / K1_TILE));
			fir_sign_exact_kernel<<<grid, K1_CH * 32, 0, st>>>(view, ctx->d_state, ctx->hist_sel, 0, fp.ch, fp.frames, n_frames, ctx->n_ch, signs);
			ctx->launches++;
		}
		if (fp.ch < ctx->n_ch) {
			dim3 grid((unsigned) ((ctx->n_ch - fp.ch + K1_CH - 1) / K1_CH), (unsigned) ((n_frames + K1_TILE - 1) / K1_TILE));
			fir_sign_exact_kernel<<<grid, K1_CH * 32, 0, st>>>(view, ctx->d_state, ctx->hist_sel, fp.ch, ctx->n_ch, 0, n_frames, ctx->n_ch, signs);
			ctx->launches++;
		}
	} else {
		nl = fir_launch(ctx->cfg.fir_mode, layout, view, ctx->d_state, ctx->hist_sel, ctx->n_ch, n_frames, signs, st, &hist_saved);
		if (nl < 0)
			return fail(GAIS_ECUDA, "FIR launch failed: %s", cudaGetErrorString(cudaGetLastError()));
		ctx->launches += nl;
	}
	if (hist_saved < ctx->n_ch) {
		save_hist_kernel<<<(ctx->n_ch - hist_saved + 127) / 128, 128, 0, st>>>(view, ctx->d_state, ctx->hist_sel, hist_saved, ctx->n_ch,
										    n_frames);
		ctx->launches++;
	}
	ctx->hist_sel ^= 1;
	if (ctx->d_peak) {
		peak_kernel<<<(unsigned) (((int64_t) ctx->n_ch * 32 + 255) / 256), 256, 0, st>>>(view, ctx->n_ch, n_frames, ctx->d_peak);
		ctx->launches++;
	}
	if (timed) CK(cudaEventRecord(ctx->ev_tile[4 * tile_idx + 1], st));
	if (overlap) {
		CK(cudaEventRecord(ctx->ev_fir_done[tile_idx & 1], st));
		CK(cudaStreamWaitEvent(st_trk, ctx->ev_fir_done[tile_idx & 1], 0));
	}

	if (timed) CK(cudaEventRecord(ctx->ev_tile[4 * tile_idx + 2], st_trk));
	if (fp.ch) {
		if (fp.frames < n_frames) {
			nl = track_launch(signs + (fp.frames / 32) * ctx->n_ch, ctx->d_state, 0, fp.ch, ctx->n_ch, n_frames - fp.frames, out, st_trk);
			if (nl < 0)
				return fail(GAIS_ECUDA, "tracking launch failed: %s", cudaGetErrorString(cudaGetLastError()));
			ctx->launches += nl;
		}
		if (fp.ch < ctx->n_ch) {
			nl = track_launch(signs, ctx->d_state, fp.ch, ctx->n_ch, ctx->n_ch, n_frames, out, st_trk);
			if (nl < 0)
				return fail(GAIS_ECUDA, "tracking launch failed: %s", cudaGetErrorString(cudaGetLastError()));
			ctx->launches += nl;
		}
	} else {
		nl = track_launch(signs, ctx->d_state, 0, ctx->n_ch, ctx->n_ch, n_frames, out, st_trk);
		if (nl < 0)
			return fail(GAIS_ECUDA, "tracking launch failed: %s", cudaGetErrorString(cudaGetLastError()));
		ctx->launches += nl;
	}
	if (timed) CK(cudaEventRecord(ctx->ev_tile[4 * tile_idx + 3], st_trk));
	if (overlap)
		CK(cudaEventRecord(ctx->ev_trk_done[tile_idx & 1], st_trk));
	CK(cudaGetLastError());
	return 0;
}

static int finish(gais_ctx *ctx);

/* every argument has been validated by the caller: from here on the previous run's results are gone */
static int begin_run(gais_ctx *ctx, int64_t n_frames, cudaStream_t st)
{
	if (n_frames < 1 || n_frames > ctx->cfg.max_frames_per_run)
		return fail(GAIS_EINVAL, "n_frames %lld outside 1..max_frames_per_run (%lld)", (long long) n_frames,
			    (long long) ctx->cfg.max_frames_per_run);
	CK(cudaSetDevice(ctx->cfg.device));
	if (ctx->pending) {
		/* a run that was never synced: complete it first (its kernels may still be using the per-run
		 * buffers on another stream); its messages are superseded by this run's, an overflow it hit is not lost */
		int rc = finish(ctx);
		if (rc)
			return rc;
	}
	ctx->launches = 0;
	ctx->finished = 0;
	ctx->text_valid = 0;
	ctx->last_frames = n_frames;
	CK(cudaMemsetAsync(ctx->d_run_count, 0, sizeof(uint32_t) * ctx->n_ch, st));
	CK(cudaMemsetAsync(ctx->d_run_bits, 0, sizeof(uint32_t) * ctx->n_ch, st));
	if (ctx->d_bits)
		CK(cudaMemsetAsync(ctx->d_bits, 0, (size_t) ctx->n_ch * ctx->bits_row_words * 4, st));
	if (ctx->d_peak)
		CK(cudaMemsetAsync(ctx->d_peak, 0, (size_t) ctx->n_ch * 2, st));
	return 0;
}

extern "C" int gais_run_device(gais_ctx *ctx, const int16_t *d_samples, int64_t n_frames, int64_t stride, void *stream)
{
	if (!ctx || !d_samples)
		return fail(GAIS_EINVAL, "null argument");
	cudaStream_t st = (cudaStream_t) stream;
	const bool planar = ctx->cfg.layout == GAIS_LAYOUT_PLANAR;
	if (planar ? stride < n_frames : stride < ctx->n_ch)
		return fail(GAIS_EINVAL, "stride %lld too small for the layout", (long long) stride);
	int rc = begin_run(ctx, n_frames, st);
	if (rc)
		return rc;

	/* tile plan.  Where the fused kernel takes every channel there is nothing to stage between kernels, so a tile is
	 * as long as one launch may be; otherwise the sign words of a tile go through global memory (ctx->tile_frames) */
	SampleView v0;
	v0.ch_stride = planar ? stride : 1;
	v0.t_stride = planar ? 1 : stride;
	v0.base = d_samples;
	int64_t tile = ctx->tile_frames;
	const bool keep = (ctx->cfg.flags & GAIS_KEEP_SIGNS) != 0;
	const FusedPart fp0 = fused_part(ctx, v0, n_frames < tile ? n_frames : tile);
	const bool relay = relay_planar(ctx, v0);
	const bool fused_all = fp0.ch == ctx->n_ch && !keep && !relay;
	if (fused_all)
		tile = X_MAX_FRAMES;
	if (relay && (rc = ensure_planar(ctx, n_frames < tile ? n_frames : tile)) != 0)
		return rc;
	int n_tiles = (int) ((n_frames + tile - 1) / tile);
	if ((rc = ensure_tile_events(ctx, n_tiles)) != 0)
		return rc;
	{
		const int64_t nf0 = n_frames < tile ? n_frames : tile;
		int64_t words = tile_sign_words(ctx, fused_part(ctx, v0, nf0), nf0);
		if (n_tiles > 1) {
			const int64_t nfl = n_frames - (int64_t) (n_tiles - 1) * tile, wl = tile_sign_words(ctx, fused_part(ctx, v0, nfl), nfl);
			if (wl > words)
				words = wl;
		}
		if ((rc = ensure_signs(ctx, words)) != 0)
			return rc;
	}
	CK(cudaEventRecord(ctx->ev[EV_START], st));
	/* overlap mode (two-kernel path only): FIR on s_fir, tracking on the high-priority s_trk, both forked from / joined to st */
	cudaStream_t sF = st, sT = st;
	if (ctx->overlap && n_tiles > 1 && fp0.ch == 0) {
		sF = ctx->s_fir;
		sT = ctx->s_trk;
		CK(cudaStreamWaitEvent(sF, ctx->ev[EV_START], 0));
		CK(cudaStreamWaitEvent(sT, ctx->ev[EV_START], 0));
	}
	for (int t = 0; t < n_tiles; t++) {
		int64_t f0 = (int64_t) t * tile;
		int64_t nf = (n_frames - f0 < tile) ? n_frames - f0 : tile;
		SampleView v = v0;
		v.base = d_samples + f0 * v.t_stride;
		if ((rc = enqueue_tile(ctx, v, nf, t, f0 / 32, sF, sT, true)) != 0)
			return rc;
	}
	if ((rc = enqueue_post(ctx, sT)) != 0)
		return rc;
	if (sT != st) {
		CK(cudaEventRecord(ctx->ev_join, sT));
		CK(cudaStreamWaitEvent(st, ctx->ev_join, 0));
	}
	CK(cudaEventRecord(ctx->ev[EV_END], st));
	CK(cudaGetLastError());
	ctx->last_stream = st;
	ctx->pending = 1;
	ctx->n_tiles_last = n_tiles;
	return 0;
}

extern "C" int gais_run_host(gais_ctx *ctx, const int16_t *h_samples, int64_t n_frames, int64_t stride)
{
	if (!ctx || !h_samples)
		return fail(GAIS_EINVAL, "null argument");
	cudaStream_t st = ctx->s_own;
	const bool planar = ctx->cfg.layout == GAIS_LAYOUT_PLANAR;
	if (planar ? stride < n_frames : stride < ctx->n_ch)
		return fail(GAIS_EINVAL, "stride %lld too small for the layout", (long long) stride);
	int rc = begin_run(ctx, n_frames, st);
	if (rc)
		return rc;

	/* staging tiles: planar [n_ch][tile] (device stride = tile), interleaved [tile][stride] */
	int64_t tile = ctx->tile_frames;
	int64_t need = planar ? (int64_t) ctx->n_ch * tile : tile * stride;
	if (need > ctx->stage_elems) {
		CK(cudaStreamSynchronize(ctx->s_copy));
		CK(cudaStreamSynchronize(st));
		cudaFree(ctx->d_stage[0]); cudaFree(ctx->d_stage[1]);
		ctx->d_stage[0] = ctx->d_stage[1] = NULL;
		ctx->stage_elems = 0;
		CK(cudaMalloc(&ctx->d_stage[0], (size_t) need * 2));
		CK(cudaMalloc(&ctx->d_stage[1], (size_t) need * 2));
		ctx->stage_elems = need;
		CK(cudaEventRecord(ctx->ev_free[0], st));
		CK(cudaEventRecord(ctx->ev_free[1], st));
	}
	int n_tiles = (int) ((n_frames + tile - 1) / tile);
	if ((rc = ensure_tile_events(ctx, n_tiles)) != 0)
		return rc;
	{
		SampleView sv;
		sv.base = ctx->d_stage[0];
		sv.ch_stride = planar ? tile : 1;
		sv.t_stride = planar ? 1 : stride;
		const int64_t nf0 = n_frames < tile ? n_frames : tile, nfl = n_frames - (int64_t) (n_tiles - 1) * tile;
		int64_t words = tile_sign_words(ctx, fused_part(ctx, sv, nf0), nf0);
		const int64_t wl = tile_sign_words(ctx, fused_part(ctx, sv, nfl), nfl);
		if ((rc = ensure_signs(ctx, wl > words ? wl : words)) != 0)
			return rc;
		if (relay_planar(ctx, sv) && (rc = ensure_planar(ctx, nf0)) != 0)
			return rc;
	}
	CK(cudaEventRecord(ctx->ev[EV_START], st));
	for (int t = 0; t < n_tiles; t++) {
		int64_t f0 = (int64_t) t * tile;
		int64_t nf = (n_frames - f0 < tile) ? n_frames - f0 : tile;
		int b = t & 1;
		/* the copy into buffer b must wait until the kernels of tile t-2 have read it */
		CK(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_free[b], 0));
		SampleView v;
		if (planar) {
			CK(cudaMemcpy2DAsync(ctx->d_stage[b], (size_t) tile * 2, h_samples + f0, (size_t) stride * 2, (size_t) nf * 2,
					     (size_t) ctx->n_ch, cudaMemcpyHostToDevice, ctx->s_copy));
			v.base = ctx->d_stage[b]; v.ch_stride = tile; v.t_stride = 1;
		} else {
			CK(cudaMemcpyAsync(ctx->d_stage[b], h_samples + f0 * stride, (size_t) nf * stride * 2, cudaMemcpyHostToDevice,
					   ctx->s_copy));
			v.base = ctx->d_stage[b]; v.ch_stride = 1; v.t_stride = stride;
		}
		CK(cudaEventRecord(ctx->ev_copy[b], ctx->s_copy));
		CK(cudaStreamWaitEvent(st, ctx->ev_copy[b], 0));
		if ((rc = enqueue_tile(ctx, v, nf, t, f0 / 32, st, st, true)) != 0)
			return rc;
		CK(cudaEventRecord(ctx->ev_free[b], st));
	}
	if ((rc = enqueue_post(ctx, st)) != 0)
		return rc;
	CK(cudaEventRecord(ctx->ev[EV_END], st));
	CK(cudaGetLastError());
	ctx->last_stream = st;
	ctx->pending = 1;
	ctx->n_tiles_last = n_tiles;
	return 0;
}

extern "C" int gais_run_bits_device(gais_ctx *ctx, const uint8_t *d_bits, int64_t n_bits, int64_t stride, void *stream)
{
	if (!ctx || !d_bits)
		return fail(GAIS_EINVAL, "null argument");
	if (stride < n_bits)
		return fail(GAIS_EINVAL, "stride %lld smaller than n_bits", (long long) stride);
	cudaStream_t st = (cudaStream_t) stream;
	int rc = begin_run(ctx, n_bits, st);
	if (rc)
		return rc;
	if ((rc = ensure_tile_events(ctx, 1)) != 0)
		return rc;
	CK(cudaEventRecord(ctx->ev[EV_START], st));
	CK(cudaEventRecord(ctx->ev_tile[0], st));
	CK(cudaEventRecord(ctx->ev_tile[1], st));
	CK(cudaEventRecord(ctx->ev_tile[2], st));
	hdlc_bits_kernel<<<(ctx->n_ch + TRK_THREADS - 1) / TRK_THREADS, TRK_THREADS, 0, st>>>(d_bits, stride, n_bits, ctx->d_state, ctx->n_ch,
												 make_out(ctx));
	ctx->launches++;
	CK(cudaEventRecord(ctx->ev_tile[3], st));
	if ((rc = enqueue_post(ctx, st)) != 0)
		return rc;
	CK(cudaEventRecord(ctx->ev[EV_END], st));
	CK(cudaGetLastError());
	ctx->last_stream = st;
	ctx->pending = 1;
	ctx->n_tiles_last = 1;
	return 0;
}

extern "C" int gais_run_bits_host(gais_ctx *ctx, const uint8_t *h_bits, int64_t n_bits, int64_t stride)
{
	if (!ctx || !h_bits)
		return fail(GAIS_EINVAL, "null argument");
	if (n_bits < 1 || n_bits > ctx->cfg.max_frames_per_run)
		return fail(GAIS_EINVAL, "n_bits %lld outside 1..max_frames_per_run (%lld)", (long long) n_bits,
			    (long long) ctx->cfg.max_frames_per_run);
	if (stride < n_bits)
		return fail(GAIS_EINVAL, "stride %lld smaller than n_bits", (long long) stride);
	CK(cudaSetDevice(ctx->cfg.device));
	/* the sample staging buffers double as the bit staging buffer (they are at least one tile of int16 per channel) */
	const int64_t need = ((int64_t) ctx->n_ch * n_bits + 1) / 2;
	if (need > ctx->stage_elems) {
		CK(cudaStreamSynchronize(ctx->s_copy));
		CK(cudaStreamSynchronize(ctx->s_own));
		cudaFree(ctx->d_stage[0]); cudaFree(ctx->d_stage[1]);
		ctx->d_stage[0] = ctx->d_stage[1] = NULL;
		ctx->stage_elems = 0;
		CK(cudaMalloc(&ctx->d_stage[0], (size_t) need * 2));
		CK(cudaMalloc(&ctx->d_stage[1], (size_t) need * 2));
		ctx->stage_elems = need;
		CK(cudaEventRecord(ctx->ev_free[0], ctx->s_own));
		CK(cudaEventRecord(ctx->ev_free[1], ctx->s_own));
	}
	CK(cudaStreamSynchronize(ctx->s_own));        /* the previous run has read the staging buffer */
	CK(cudaMemcpy2DAsync(ctx->d_stage[0], (size_t) n_bits, h_bits, (size_t) stride, (size_t) n_bits, (size_t) ctx->n_ch,
			     cudaMemcpyHostToDevice, ctx->s_own));
	return gais_run_bits_device(ctx, (const uint8_t *) ctx->d_stage[0], n_bits, n_bits, ctx->s_own);
}

/* complete the last run: wait, size and fill the dense message array */
static int finish(gais_ctx *ctx)
{
	CK(cudaSetDevice(ctx->cfg.device));
	if (!ctx->pending)
		return 0;
	cudaStream_t st = ctx->last_stream;
	uint64_t total = 0;
	int32_t ovf = 0;
	CK(cudaMemcpyAsync(&total, ctx->d_offsets + ctx->n_ch, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(&ovf, ctx->d_overflow, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	if ((int64_t) total > ctx->dense_cap) {
		cudaFree(ctx->d_dense);
		ctx->d_dense = NULL;
		ctx->dense_cap = 0;
		int64_t cap = (int64_t) total + (int64_t) total / 8 + 1024;
		CK(cudaMalloc(&ctx->d_dense, (size_t) cap * sizeof(gais_msg)));
		ctx->dense_cap = cap;
	}
	CK(cudaEventRecord(ctx->ev[EV_POST0], st));
	if (total > 0) {
		int64_t threads = (int64_t) ctx->n_ch * 32;
		gather_msgs_kernel<<<(unsigned) ((threads + 255) / 256), 256, 0, st>>>(ctx->d_slots, ctx->slot_cap, ctx->d_run_count,
										      ctx->d_offsets, ctx->n_ch, (uint32_t) ctx->cfg.reserved[3], ctx->d_dense);
		ctx->launches++;
	}
	CK(cudaEventRecord(ctx->ev[EV_POST1], st));
	CK(cudaStreamSynchronize(st));
	CK(cudaGetLastError());
	ctx->n_msgs = (int64_t) total;
	ctx->pending = 0;
	ctx->finished = 1;

	gais_timing &tm = ctx->timing;
	memset(&tm, 0, sizeof(tm));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, ctx->ev[EV_START], ctx->ev[EV_END]));
	tm.total_ms = ms;
	for (int t = 0; t < ctx->n_tiles_last; t++) {
		CK(cudaEventElapsedTime(&ms, ctx->ev_tile[4 * t], ctx->ev_tile[4 * t + 1]));
		tm.fir_ms += ms;
		CK(cudaEventElapsedTime(&ms, ctx->ev_tile[4 * t + 2], ctx->ev_tile[4 * t + 3]));
		tm.track_ms += ms;
	}
	CK(cudaEventElapsedTime(&ms, ctx->ev[EV_POST0], ctx->ev[EV_POST1]));
	tm.post_ms = ms;
	tm.total_ms += ms;
	tm.launches = ctx->launches;
	if (ovf) {
		CK(cudaMemset(ctx->d_overflow, 0, 4));
		return fail(GAIS_EOVERFLOW, "a channel produced more than %d messages in one run (raise reserved[0])", ctx->slot_cap);
	}
	return 0;
}

extern "C" int gais_sync(gais_ctx *ctx)
{
	if (!ctx)
		return fail(GAIS_EINVAL, "null ctx");
	return finish(ctx);
}

extern "C" int gais_message_count(gais_ctx *ctx, int64_t *n_msgs)
{
	if (!ctx || !n_msgs)
		return fail(GAIS_EINVAL, "null argument");
	int rc = finish(ctx);
	*n_msgs = ctx->n_msgs;
	return rc;
}

extern "C" int gais_device_messages(gais_ctx *ctx, const gais_msg **d_msgs, int64_t *n_msgs)
{
	if (!ctx || !d_msgs || !n_msgs)
		return fail(GAIS_EINVAL, "null argument");
	int rc = finish(ctx);
	*d_msgs = ctx->d_dense;
	*n_msgs = ctx->n_msgs;
	return rc;
}

extern "C" int gais_get_messages(gais_ctx *ctx, gais_msg *h_out, int64_t cap, int64_t *n_msgs)
{
	if (!ctx || !n_msgs)
		return fail(GAIS_EINVAL, "null argument");
	int rc = finish(ctx);
	*n_msgs = ctx->n_msgs;
	if (rc)
		return rc;
	int64_t n = ctx->n_msgs < cap ? ctx->n_msgs : cap;
	if (n > 0 && h_out)
		CK(cudaMemcpy(h_out, ctx->d_dense, (size_t) n * sizeof(gais_msg), cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int gais_host_alloc(void **h_ptr, size_t bytes)
{
	if (!h_ptr)
		return fail(GAIS_EINVAL, "null argument");
	*h_ptr = NULL;
	cudaError_t e = cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(e == cudaErrorMemoryAllocation ? GAIS_ENOMEM : GAIS_ECUDA, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
	}
	return 0;
}

extern "C" void gais_host_free(void *h_ptr)
{
	if (h_ptr)
		cudaFreeHost(h_ptr);
}

extern "C" int gais_get_nmea(gais_ctx *ctx, gais_nmea_rec *h_out, int64_t cap, int64_t *n_msgs)
{
	if (!ctx || !n_msgs)
		return fail(GAIS_EINVAL, "null argument");
	int rc = finish(ctx);
	*n_msgs = ctx->n_msgs;
	if (rc)
		return rc;
	int64_t n = ctx->n_msgs < cap ? ctx->n_msgs : cap;
	if (n <= 0 || !h_out)
		return 0;
	if (ctx->n_msgs > ctx->nmea_cap) {
		cudaFree(ctx->d_nmea);
	cudaFree(ctx->d_text); cudaFree(ctx->d_text_off); cudaFree(ctx->d_blk_off); cudaFree(ctx->d_blk_tot);
	for (int i = 0; i < 2; i++)
		if (ctx->ev_nmea[i]) cudaEventDestroy(ctx->ev_nmea[i]);
		ctx->d_nmea = NULL;
		ctx->nmea_cap = 0;
		CK(cudaMalloc(&ctx->d_nmea, (size_t) ctx->n_msgs * sizeof(gais_nmea_rec)));
		ctx->nmea_cap = ctx->n_msgs;
	}
	nmea_kernel<<<(unsigned) ((ctx->n_msgs + 127) / 128), 128, 0, ctx->last_stream>>>(ctx->d_dense, ctx->n_msgs, ctx->d_nmea);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(h_out, ctx->d_nmea, (size_t) n * sizeof(gais_nmea_rec), cudaMemcpyDeviceToHost, ctx->last_stream));
	CK(cudaStreamSynchronize(ctx->last_stream));
	return 0;
}

/* packed NMEA text of the last run: lengths per block, scan of the block totals, one thread per message via shared memory */
static int build_text(gais_ctx *ctx)
{
	int rc = finish(ctx);
	if (rc)
		return rc;
	if (ctx->text_valid)
		return 0;
	cudaStream_t st = ctx->last_stream;
	const int64_t n = ctx->n_msgs;
	ctx->text_bytes = 0;
	ctx->timing.nmea_ms = 0;
	if (n > 0) {
		const int64_t nblk = (n + NM_BLOCK_ITEMS - 1) / NM_BLOCK_ITEMS;
		if (n > ctx->text_msgs_cap) {
			cudaFree(ctx->d_text_off); cudaFree(ctx->d_blk_off); cudaFree(ctx->d_blk_tot);
			ctx->d_text_off = ctx->d_blk_off = NULL; ctx->d_blk_tot = NULL;
			ctx->text_msgs_cap = 0;
			const int64_t cap = n + n / 8 + 1024, cblk = (cap + NM_BLOCK_ITEMS - 1) / NM_BLOCK_ITEMS;
			CK(cudaMalloc(&ctx->d_text_off, (size_t) (cap + 1) * 8));
			CK(cudaMalloc(&ctx->d_blk_off, (size_t) (cblk + 1) * 8));
			CK(cudaMalloc(&ctx->d_blk_tot, (size_t) cblk * 4));
			ctx->text_msgs_cap = cap;
		}
		if (!ctx->ev_nmea[0]) {
			CK(cudaEventCreate(&ctx->ev_nmea[0]));
			CK(cudaEventCreate(&ctx->ev_nmea[1]));
		}
		if (n * NM_MAX_TEXT > ctx->text_cap) {
			/* sized by the message count (113 bytes each at most, 49 typically) so that no byte count has to come back
			 * to the host between the kernels; kept between runs */
			cudaFree(ctx->d_text);
			ctx->d_text = NULL;
			ctx->text_cap = 0;
			const int64_t cap = (n + n / 8 + 1024) * NM_MAX_TEXT;
			CK(cudaMalloc(&ctx->d_text, (size_t) cap));
			ctx->text_cap = cap;
		}
		CK(cudaEventRecord(ctx->ev_nmea[0], st));
		nmea_len_kernel<<<(unsigned) nblk, NM_BLOCK_ITEMS, 0, st>>>(ctx->d_dense, n, ctx->d_blk_tot);
		scan_counts_kernel<<<1, 1024, 0, st>>>(ctx->d_blk_tot, (int) nblk, ctx->d_blk_off);
		nmea_write_kernel<<<(unsigned) nblk, NM_BLOCK_ITEMS, 0, st>>>(ctx->d_dense, n, ctx->d_blk_off, ctx->d_text, ctx->d_text_off);
		CK(cudaEventRecord(ctx->ev_nmea[1], st));
		uint64_t total = 0;
		CK(cudaMemcpyAsync(&total, ctx->d_blk_off + nblk, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		CK(cudaGetLastError());
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, ctx->ev_nmea[0], ctx->ev_nmea[1]));
		ctx->timing.nmea_ms = ms;
		ctx->text_bytes = (int64_t) total;
		ctx->launches += 3;
	}
	ctx->text_valid = 1;
	return 0;
}

extern "C" int gais_device_nmea(gais_ctx *ctx, const char **d_text, const uint64_t **d_offsets, int64_t *n_msgs, int64_t *n_bytes)
{
	if (!ctx || !d_text || !d_offsets || !n_msgs || !n_bytes)
		return fail(GAIS_EINVAL, "null argument");
	int rc = build_text(ctx);
	*d_text = ctx->d_text;
	*d_offsets = ctx->d_text_off;
	*n_msgs = ctx->n_msgs;
	*n_bytes = ctx->text_bytes;
	return rc;
}

extern "C" int gais_get_nmea_text(gais_ctx *ctx, char *h_text, int64_t cap, int64_t *n_bytes)
{
	if (!ctx || !n_bytes)
		return fail(GAIS_EINVAL, "null argument");
	int rc = build_text(ctx);
	*n_bytes = ctx->text_bytes;
	if (rc)
		return rc;
	const int64_t n = ctx->text_bytes < cap ? ctx->text_bytes : cap;
	if (n > 0 && h_text)
		CK(cudaMemcpy(h_text, ctx->d_text, (size_t) n, cudaMemcpyDeviceToHost));
	return 0;
}

struct StateRow { gais_counters cnt; gais_chan_state st; };

__global__ void export_state_kernel(const ChanState *__restrict__ st, int n, gais_counters *cnt, gais_chan_state *cs)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n)
		return;
	if (cnt) {
		cnt[c].ok = st[c].ok; cnt[c].crcfail = st[c].crcfail; cnt[c].sizefail = st[c].sizefail;
	}
	if (cs) {
		cs[c].pll = st[c].pll; cs[c].prev = st[c].prev; cs[c].lastbit = st[c].lastbit;
		cs[c].fsm_state = hdlc_public_state(st[c].fsm); cs[c].seqnr = st[c].seqnr; cs[c].n_bits = st[c].n_bits;
	}
}

static int export_state(gais_ctx *ctx, gais_counters *h_cnt, gais_chan_state *h_cs)
{
	int rc = finish(ctx);
	if (rc && rc != GAIS_EOVERFLOW)
		return rc;
	void *d = NULL;
	size_t bytes = (size_t) ctx->n_ch * (h_cnt ? sizeof(gais_counters) : sizeof(gais_chan_state));
	CK(cudaMalloc(&d, bytes));
	export_state_kernel<<<(ctx->n_ch + 127) / 128, 128>>>(ctx->d_state, ctx->n_ch, h_cnt ? (gais_counters *) d : NULL,
							       h_cs ? (gais_chan_state *) d : NULL);
	cudaError_t e = cudaMemcpy(h_cnt ? (void *) h_cnt : (void *) h_cs, d, bytes, cudaMemcpyDeviceToHost);
	cudaFree(d);
	CK(e);
	return 0;
}

extern "C" int gais_get_counters(gais_ctx *ctx, gais_counters *h_out)
{
	if (!ctx || !h_out)
		return fail(GAIS_EINVAL, "null argument");
	return export_state(ctx, h_out, NULL);
}

extern "C" int gais_get_state(gais_ctx *ctx, gais_chan_state *h_out)
{
	if (!ctx || !h_out)
		return fail(GAIS_EINVAL, "null argument");
	return export_state(ctx, NULL, h_out);
}

extern "C" int gais_get_totals(gais_ctx *ctx, int64_t totals[3])
{
	if (!ctx || !totals)
		return fail(GAIS_EINVAL, "null argument");
	int rc = finish(ctx);
	if (rc && rc != GAIS_EOVERFLOW)
		return rc;
	CK(cudaMemset(ctx->d_totals, 0, 24));
	totals_kernel<<<(ctx->n_ch + 255) / 256, 256>>>(ctx->d_state, ctx->n_ch, ctx->d_totals);
	CK(cudaMemcpy(totals, ctx->d_totals, 24, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int gais_bits_row_words(gais_ctx *ctx, int64_t *words_per_row)
{
	if (!ctx || !words_per_row)
		return fail(GAIS_EINVAL, "null argument");
	*words_per_row = ctx->bits_row_words;
	return 0;
}

extern "C" int gais_get_bits(gais_ctx *ctx, uint32_t *h_words, uint32_t *h_nbits)
{
	if (!ctx || !h_words || !h_nbits)
		return fail(GAIS_EINVAL, "null argument");
	if (!ctx->d_bits)
		return fail(GAIS_EINVAL, "context was created without GAIS_KEEP_BITS");
	int rc = finish(ctx);
	if (rc && rc != GAIS_EOVERFLOW)
		return rc;
	CK(cudaMemcpy(h_words, ctx->d_bits, (size_t) ctx->n_ch * ctx->bits_row_words * 4, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(h_nbits, ctx->d_run_bits, (size_t) ctx->n_ch * 4, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int gais_get_signs(gais_ctx *ctx, uint32_t *h_words, int64_t cap_words)
{
	if (!ctx || !h_words)
		return fail(GAIS_EINVAL, "null argument");
	if (!(ctx->cfg.flags & GAIS_KEEP_SIGNS))
		return fail(GAIS_EINVAL, "context was created without GAIS_KEEP_SIGNS");
	int rc = finish(ctx);
	if (rc && rc != GAIS_EOVERFLOW)
		return rc;
	int64_t words = ((ctx->last_frames + 31) / 32) * ctx->n_ch;
	if (words > cap_words)
		words = cap_words;
	CK(cudaMemcpy(h_words, ctx->d_signs[0], (size_t) words * 4, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int gais_get_peaks(gais_ctx *ctx, int16_t *h_out)
{
	if (!ctx || !h_out)
		return fail(GAIS_EINVAL, "null argument");
	if (!ctx->d_peak)
		return fail(GAIS_EINVAL, "context was created without GAIS_KEEP_PEAK");
	int rc = finish(ctx);
	if (rc && rc != GAIS_EOVERFLOW)
		return rc;
	CK(cudaMemcpy(h_out, ctx->d_peak, (size_t) ctx->n_ch * 2, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int gais_get_timing(gais_ctx *ctx, gais_timing *out)
{
	if (!ctx || !out)
		return fail(GAIS_EINVAL, "null argument");
	int rc = finish(ctx);
	*out = ctx->timing;
	return rc;
}

extern "C" int gais_nmea_format(const gais_msg *msg, char *out)
{
	if (!msg || !out)
		return fail(GAIS_EINVAL, "null argument");
	return gn_format(msg->payload, msg->nbits, msg->flags & 15, out);
}

/*
 * gais_kernels.cuh -- sm_100a kernels of the batched AIS receive path.
 *
 *   fir_sign_kernel   K1  int16 -> 36-tap FIR -> sign bit per sample (32 samples / word)
 *   save_hist_kernel      carries the last 36 samples of the run into the next one
 *   track_kernel      K2+K3  zero-crossing DPLL, slicer, NRZI, HDLC FSM (gais_track.cuh)
 *   crc/finalize          CRC-16, counters, seqnr, records (gais_track.cuh)
 *   scan/gather       K4  dense (channel, end_bit)-ordered message array
 *   nmea_kernel       K5  !AIVDM armouring on the GPU
 *
 * Reference behaviour each kernel reproduces is cited at the kernel.
 */
#ifndef GAIS_KERNELS_CUH
#define GAIS_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "gais_b200.h"
#include "gais_state.h"
#include "gais_nmea.h"

namespace gais {

__constant__ float c_taps[GAIS_NTAPS];

/* sample (c, n) of the run; n < 0 reads the carried history x[n_prev_end + n] */
struct SampleView {
	const int16_t *base;
	int64_t ch_stride, t_stride;
};

/* ------------------------------------------------------------------------------------------
 * K1 (exact flavour).  out[n] = sum_{i<36} tap[i] * x[n-36+i]  -- the window is the 36 samples
 * BEFORE the one just stored (src/filter.c:115-125) -- as a float32 sum, multiply then add,
 * strictly in tap order (src/filter.h:40-49), no FMA contraction, denormals honoured.  Only
 * (out[n] > 0) is ever consumed (src/receiver.c:110-111,126).  Taps 0,1,34,35 are exactly
 * +0.0f: their products are +-0 and s + (+-0) == s for every s this sum can reach (s is never
 * -0), so they are skipped.
 *
 * Block = 8 warps = 8 adjacent channels x one 1024-sample tile; lane l owns outputs
 * [32l, 32l+32) of its warp's channel -> one sign word.  signs layout: [word][channel].
 * ------------------------------------------------------------------------------------------ */
constexpr int K1_CH = 8;
constexpr int K1_TILE = 1024;
constexpr int K1_ROW = K1_TILE + GAIS_NTAPS + 4;   /* int16 elements per smem row (padded) */

__device__ __forceinline__ float exact_fir(const float *xs /* 36 window values */)
{
	float s = 0.0f;
#pragma unroll
	for (int i = 2; i < GAIS_NTAPS - 2; i++)
		s = __fadd_rn(s, __fmul_rn(xs[i], c_taps[i]));
	return s;
}

__global__ void __launch_bounds__(K1_CH * 32)
fir_sign_exact_kernel(SampleView in, const ChanState *__restrict__ st, int hist_sel, int c_begin, int c_end,
		      int64_t n_begin, int64_t n_end, int n_channels, uint32_t *__restrict__ signs)
{
	/* channels [c_begin, c_end), samples [n_begin, n_end) of the tile; n_begin % 32 == 0.
	 * Sign-word format (device): LSB first, bit j of word w = (out[32w + j] > 0). */
	__shared__ int16_t tile[K1_CH][K1_ROW];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int c = c_begin + blockIdx.x * K1_CH + warp;
	const int64_t n0 = n_begin + (int64_t) blockIdx.y * K1_TILE;

	if (c < c_end) {
		const int16_t *row = in.base + (int64_t) c * in.ch_stride;
		for (int i = lane; i < K1_TILE + GAIS_NTAPS; i += 32) {
			int64_t n = n0 - GAIS_NTAPS + i;
			int16_t v = 0;
			if (n < 0)
				v = st[c].hist[hist_sel][GAIS_NTAPS + n];
			else if (n < n_end)
				v = row[n * in.t_stride];
			tile[warp][i] = v;
		}
	}
	__syncwarp();
	if (c >= c_end)
		return;
	if (n0 + 32 * lane >= n_end)
		return;

	float xs[GAIS_NTAPS + 31];
#pragma unroll
	for (int i = 0; i < GAIS_NTAPS + 31; i++)
		xs[i] = (float) tile[warp][32 * lane + i];

	uint32_t word = 0;
#pragma unroll
	for (int j = 0; j < 32; j++) {
		float s = exact_fir(&xs[j]);
		word |= (s > 0.0f ? 1u : 0u) << j;
	}
	signs[(n0 / 32 + lane) * n_channels + c] = word;
}

/* next tile's history = last 36 samples seen (src/filter.c:129-134 keeps exactly these); channels
 * [c_begin, n_channels) -- the fast FIR kernel saves its own */
__global__ void save_hist_kernel(SampleView in, ChanState *st, int hist_sel, int c_begin, int n_channels, int64_t n_frames)
{
	int c = c_begin + blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_channels)
		return;
	const int16_t *row = in.base + (int64_t) c * in.ch_stride;
	int16_t v[GAIS_NTAPS];
#pragma unroll
	for (int i = 0; i < GAIS_NTAPS; i++) {
		int64_t n = n_frames - GAIS_NTAPS + i;
		v[i] = (n < 0) ? st[c].hist[hist_sel][GAIS_NTAPS + n] : row[n * in.t_stride];
	}
#pragma unroll
	for (int i = 0; i < GAIS_NTAPS; i++)
		st[c].hist[hist_sel ^ 1][i] = v[i];
}

/* Frame-interleaved samples [time][t_stride] (the reference's own buffer shape, src/receiver.c:102: channel c of frame n at
 * buf[n * num_ch + c]) -> planar rows [channel][out_stride], so that batches of interleaved channels take the same fast kernels
 * as planar input.  One block moves a 32-channel x 64-sample tile through shared memory (both sides in 64- / 128-byte runs). */
__global__ void __launch_bounds__(256)
deinterleave_kernel(const int16_t *__restrict__ in, int64_t t_stride, int n_channels, int64_t n_frames, int16_t *__restrict__ out,
		    int64_t out_stride)
{
	__shared__ int16_t tile[64][33];
	const int c0 = blockIdx.x * 32;
	const int64_t n0 = (int64_t) blockIdx.y * 64;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	for (int i = ty; i < 64; i += 8) {
		const int64_t n = n0 + i;
		tile[i][tx] = (n < n_frames && c0 + tx < n_channels) ? in[n * t_stride + c0 + tx] : (int16_t) 0;
	}
	__syncthreads();
	for (int i = ty; i < 32; i += 8) {
		const int c = c0 + i;
		if (c >= n_channels)
			continue;
		int16_t *row = out + (int64_t) c * out_stride + n0;
		if (n0 + tx < n_frames)
			row[tx] = tile[tx][i];
		if (n0 + 32 + tx < n_frames)
			row[32 + tx] = tile[32 + tx][i];
	}
}

/* The level filter_run_buf() returns (src/filter.c:112-119): the maximum of the samples of the call, starting from
 * 0 -- i.e. positive samples only, SURVEY.md H6 -- which receiver_run() turns into the "Level on ch" log line
 * (src/receiver.c:137-147).  Optional (GAIS_KEEP_PEAK): one warp per channel, max-combined over the tiles of a run. */
__global__ void peak_kernel(SampleView in, int n_channels, int64_t n_frames, int16_t *__restrict__ peak)
{
	const int c = (int) (((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (c >= n_channels)
		return;
	const int16_t *row = in.base + (int64_t) c * in.ch_stride;
	int m = 0;
	if (in.t_stride == 1 && (in.ch_stride % 8) == 0 && ((uintptr_t) in.base % 16) == 0) {
		const uint4 *v = reinterpret_cast<const uint4 *>(row);
		for (int64_t i = lane; i < n_frames / 8; i += 32) {
			const uint4 q = v[i];
			const uint32_t w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
			for (int k = 0; k < 4; k++) {
				m = max(m, (int) (int16_t) (w[k] & 0xffffu));
				m = max(m, (int) (int16_t) (w[k] >> 16));
			}
		}
		for (int64_t n = n_frames / 8 * 8 + lane; n < n_frames; n += 32)
			m = max(m, (int) row[n]);
	} else {
		for (int64_t n = lane; n < n_frames; n += 32)
			m = max(m, (int) row[n * in.t_stride]);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		m = max(m, __shfl_down_sync(0xffffffffu, m, o));
	if (lane == 0)
		peak[c] = (int16_t) max(m, (int) peak[c]);
}

/* ------------------------------------------------------------------------------------------
 * K2+K3.  One lane per channel, sequential in time.
 *   DPLL/slicer/NRZI: src/receiver.c:109-135.  HDLC FSM: src/protodec.c:988-1122.
 *   CRC: src/protodec.c:106-167.  type gate + seqnr: src/protodec.c:896-929.
 * ------------------------------------------------------------------------------------------ */
struct TrackOut {
	gais_msg *slots;          /* [n_channels][slot_cap] */
	uint32_t *run_count;      /* [n_channels] messages written this run */
	uint32_t *bits;           /* [n_channels][bits_row_words] or NULL */
	uint32_t *run_bits;       /* [n_channels] bits produced this run */
	int32_t slot_cap;
	int32_t bits_row_words;
	int32_t *overflow;        /* set to 1 if a channel ran out of slots */
};

/* ------------------------------------------------------------------------------------------
 * K4: exclusive scan of per-channel message counts (single block, sequential over chunks of
 * 1024 -- n_channels <= a few hundred thousand) and gather into the dense array.
 * ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const uint32_t *__restrict__ counts, int n, uint64_t *__restrict__ offsets /* n+1 */)
{
	__shared__ uint32_t warp_sums[32];
	__shared__ uint64_t carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < n; base += 1024) {
		int i = base + threadIdx.x;
		uint32_t v = (i < n) ? counts[i] : 0u, x = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= d) x += y;
		}
		if (lane == 31) warp_sums[warp] = x;
		__syncthreads();
		if (warp == 0) {
			uint32_t ws = warp_sums[lane], xs = ws;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				uint32_t y = __shfl_up_sync(0xffffffffu, xs, d);
				if (lane >= d) xs += y;
			}
			warp_sums[lane] = xs - ws;   /* exclusive */
		}
		__syncthreads();
		uint64_t excl = carry + warp_sums[warp] + (x - v);
		if (i < n) offsets[i] = excl;
		__syncthreads();
		if (threadIdx.x == 1023) carry = excl + v;
		__syncthreads();
	}
	if (threadIdx.x == 0) offsets[n] = carry;
}

/* one warp per channel: copy count[c] 64-byte records, 16 B per lane; the channel field (word 14 = third word of
 * every fourth 16-byte piece) becomes first_channel + c, the channel's number in the whole multi-GPU batch */
__global__ void gather_msgs_kernel(const gais_msg *__restrict__ slots, int slot_cap, const uint32_t *__restrict__ counts,
				   const uint64_t *__restrict__ offsets, int n_channels, uint32_t first_channel, gais_msg *__restrict__ dense)
{
	int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (c >= n_channels)
		return;
	uint32_t cnt = counts[c];
	const uint4 *src = reinterpret_cast<const uint4 *>(slots + (int64_t) c * slot_cap);
	uint4 *dst = reinterpret_cast<uint4 *>(dense + offsets[c]);
	for (uint32_t i = lane; i < cnt * 4u; i += 32) {
		uint4 v = src[i];
		if ((i & 3u) == 3u)
			v.z += first_channel;
		dst[i] = v;
	}
}

/* K5 (fixed-stride form, kept for gais_get_nmea): one thread per message */
__global__ void nmea_kernel(const gais_msg *__restrict__ msgs, int64_t n, gais_nmea_rec *__restrict__ out)
{
	int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	gais_msg m = msgs[i];
	char text[GAIS_NMEA_STRIDE];
	int len = gn_format(m.payload, m.nbits, m.flags & 15, text);
	gais_nmea_rec *r = &out[i];
	r->len = (uint8_t) len;
	for (int k = 0; k < len; k++)
		r->text[k] = text[k];
}

/* ------------------------------------------------------------------------------------------
 * K5, packed form: the "!AIVDM...\r\n" text of all messages of a run back to back, message i at
 * text[offsets[i] .. offsets[i+1]).  Same bytes as gn_format() (src/protodec.c:780-894, :896-929).
 *   nmea_len_kernel        text length of every message, summed per block of 256 messages
 *   (scan_counts_kernel)   exclusive scan of the block totals
 *   nmea_write_kernel      one THREAD per message writes its text into shared memory at its place inside the
 *                          block's text (lengths scanned again inside the block); the block then copies its
 *                          piece of the text to global memory, consecutive bytes by consecutive threads
 * A sentence is  "!AIVDM,n,s,"(11) + "q,," or ",A,"(3) + <= 61 six-bit characters + ",f*hh\r\n"(7); a message is
 * at most two sentences (426 payload bits).
 * ------------------------------------------------------------------------------------------ */
constexpr int NM_BLOCK_ITEMS = 256;
constexpr int NM_MAX_TEXT = 2 * 21 + 71;            /* 113 bytes per message at most */

/* words 0 and 13 of a record: type gate and text length (0 when the gate drops the message) */
__device__ __forceinline__ uint32_t gn_text_len(uint32_t w0, uint32_t w13, int &nch, int &nsent, int &fill)
{
	const int nbits = (int) (w13 >> 16);
	const unsigned type = (w0 & 0xffu) >> 2;
	fill = (nbits % 6) ? 6 - nbits % 6 : 0;
	const int total = nbits + fill;
	nch = total / 6;
	nsent = (total <= 366) ? 1 : 2;
	if (!((w13 >> 12) & 1u) || type < 1 || type > 24 || nch > 71)      /* flags bit 4: type gate passed */
		return 0u;
	return (uint32_t) (21 * nsent + nch);
}

/* inclusive scan of one value per thread over a block of 256; returns the exclusive prefix, *total = block sum */
__device__ __forceinline__ uint32_t block_scan_256(uint32_t v, uint32_t *warp_sums /* [8] shared */, uint32_t *total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t x = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
		if (lane >= d) x += y;
	}
	if (lane == 31) warp_sums[warp] = x;
	__syncthreads();
	uint32_t base = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < NM_BLOCK_ITEMS / 32; w++) {
		const uint32_t ws = warp_sums[w];
		if (w < warp) base += ws;
		tot += ws;
	}
	*total = tot;
	return base + x - v;
}

__global__ void __launch_bounds__(NM_BLOCK_ITEMS)
nmea_len_kernel(const gais_msg *__restrict__ msgs, int64_t n, uint32_t *__restrict__ block_tot)
{
	__shared__ uint32_t warp_sums[NM_BLOCK_ITEMS / 32];
	const int64_t i = (int64_t) blockIdx.x * NM_BLOCK_ITEMS + threadIdx.x;
	uint32_t len = 0;
	if (i < n) {
		const uint32_t *w = reinterpret_cast<const uint32_t *>(&msgs[i]);
		int nch, nsent, fill;
		len = gn_text_len(w[0], w[13], nch, nsent, fill);
	}
	uint32_t tot;
	block_scan_256(len, warp_sums, &tot);
	if (threadIdx.x == 0)
		block_tot[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(NM_BLOCK_ITEMS)
nmea_write_kernel(const gais_msg *__restrict__ msgs, int64_t n, const uint64_t *__restrict__ block_off, char *__restrict__ text,
		  uint64_t *__restrict__ offsets)
{
	__shared__ uint32_t warp_sums[NM_BLOCK_ITEMS / 32];
	__shared__ __align__(16) char buf[NM_BLOCK_ITEMS * NM_MAX_TEXT];
	const int64_t i = (int64_t) blockIdx.x * NM_BLOCK_ITEMS + threadIdx.x;
	uint32_t w[14], len = 0;
	int nch = 0, nsent = 0, fill = 0;
	if (i < n) {
		/* the record: four 16-byte loads, a warp reads 2 KB of consecutive records */
		const uint4 *src = reinterpret_cast<const uint4 *>(&msgs[i]);
		const uint4 a = src[0], b = src[1], c = src[2], d = src[3];
		w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
		w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w; w[12] = d.x; w[13] = d.y;
		len = gn_text_len(w[0], w[13], nch, nsent, fill);
	}
	uint32_t tot;
	const uint32_t at = block_scan_256(len, warp_sums, &tot);
	const uint64_t base = block_off[blockIdx.x];
	if (i < n) {
		offsets[i] = base + at;
		if (i == n - 1)
			offsets[n] = base + at + len;
	}
	if (len) {
		const int nbytes = (int) (w[13] >> 16) >> 3, seqnr = (int) ((w[13] >> 8) & 15u);
		w[13] &= 0xffu;                      /* payload byte 52; bytes at or beyond nbytes read as 0 below */
		char *o = buf + at;
		int ci = 0;                          /* payload character index */
		for (int s = 1; s <= nsent; s++) {
			unsigned cs = 0;
			const int ncs = (s == 1) ? (nch < 61 ? nch : 61) : nch - 61;
#define NM_PUT(ch_) do { const char c_ = (char) (ch_); *o++ = c_; cs ^= (unsigned char) c_; } while (0)
			*o++ = '!';
			NM_PUT('A'); NM_PUT('I'); NM_PUT('V'); NM_PUT('D'); NM_PUT('M'); NM_PUT(',');
			NM_PUT('0' + nsent); NM_PUT(','); NM_PUT('0' + s); NM_PUT(',');
			if (nsent > 1) { NM_PUT('0' + seqnr); NM_PUT(','); NM_PUT(','); }
			else { NM_PUT(','); NM_PUT('A'); NM_PUT(','); }
			/* 24 payload bits = 3 bytes = 4 characters at a time; byte j of the payload is byte j & 3 of word j >> 2 */
#pragma unroll
			for (int g = 0; g < 18; g++) {
				if (4 * g + 3 < ci || 4 * g >= ci + ncs)
					continue;                       /* group outside this sentence */
				unsigned v24 = 0;
#pragma unroll
				for (int k = 0; k < 3; k++) {
					const int j = 3 * g + k;
					const unsigned byte = (j < nbytes) ? (w[j >> 2] >> (8 * (j & 3))) & 255u : 0u;
					v24 = (v24 << 8) | byte;
				}
#pragma unroll
				for (int k = 0; k < 4; k++) {
					const int cidx = 4 * g + k;
					if (cidx >= ci && cidx < ci + ncs) {
						const unsigned v = (v24 >> (18 - 6 * k)) & 63u;
						NM_PUT(v < 40 ? v + 48 : v + 56);
					}
				}
			}
			ci += ncs;
			NM_PUT(',');
			NM_PUT('0' + ((nsent > 1 && s == nsent) ? fill : 0));
#undef NM_PUT
			*o++ = '*';
			*o++ = gn_hex((cs >> 4) & 15u);
			*o++ = gn_hex(cs & 15u);
			*o++ = '\r';
			*o++ = '\n';
		}
	}
	__syncthreads();
	/* the block's text leaves as one contiguous piece: consecutive bytes by consecutive threads */
	char *dst = text + base;
	for (uint32_t k = threadIdx.x; k < tot; k += NM_BLOCK_ITEMS)
		dst[k] = buf[k];
}

/* sum of counters over channels (3 x int64) */
__global__ void totals_kernel(const ChanState *__restrict__ st, int n_channels, unsigned long long *totals)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	long long a = 0, b = 0, d = 0;
	if (c < n_channels) { a = st[c].ok; b = st[c].crcfail; d = st[c].sizefail; }
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		a += __shfl_down_sync(0xffffffffu, a, o);
		b += __shfl_down_sync(0xffffffffu, b, o);
		d += __shfl_down_sync(0xffffffffu, d, o);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&totals[0], (unsigned long long) a);
		atomicAdd(&totals[1], (unsigned long long) b);
		atomicAdd(&totals[2], (unsigned long long) d);
	}
}

} /* namespace gais */
#endif

/*
 * gais_fir.cuh -- K1, the FIR-sign stage, and its launch logic.
 *
 * What must come out is sign-exact: cur = (filter_run_buf() output > 0) for every sample, with
 * the reference's float32 rounding (src/filter.h:40-49 sequential mul+add, src/filter.c:115-125
 * window = the 36 samples before the current one, src/receiver.c:110-111 only the sign is used).
 *
 * Fast kernel (sm_100a): "guard-banded" evaluation.
 *   tier 1   a10 = sum over the 10 centre taps 13..22 (packed fp32x2 FMAs, any rounding).
 *            |a10 - R| <= 0.3568 for ANY int16 input, R being the reference's rounded sum (u = 2^-24,
 *            |x| <= 32768; a product t_i x_i that enters the reference's chain at position i is rounded
 *            once by the multiply and once by each of the 34 - i additions that follow, so its relative
 *            error is at most (35 - i) u; the first addition, 0 + p, is exact):
 *              0.0855  32768 u * sum_i t_i (35 - i) = 32768 u * 43.7      (the reference's own rounding)
 *              0.2444  2 * (t_12 + t_11 + ...) * 32768                    (dropped taps)
 *              0.0269  32768 u * sum_k t_k (10 - k) = 32768 u * 13.8      (this kernel's 10-term FMA chain)
 *            so |a10| > E1 = 0.36  =>  sign(R) = sign(a10) and R != 0.
 *   tier 2   (only samples with |a10| <= E1, ~4e-4 of noisy audio) 12 taps 12..23, with a
 *            DATA-DEPENDENT bound: S12 = sum_{12 taps} t_i |x_i| is computed alongside and
 *              |a12 - R| <= gamma_32 * (S12 + 0.001775) + 0.001775 + 12 u S12
 *                        <= 0.00355 + 2.64e-6 * S12          (0.001775 = outer taps at full scale)
 *            -> E2 = 0.004 + 3.0e-6 * S12 (0.006 on idle noise, 0.09 inside a burst).
 *   tier 3   (|a12| <= E2, ~1e-5 of samples) the exact 32-term chain, __fmul_rn/__fadd_rn in tap order.
 *
 * Data movement: one CTA = 64 channels x a run of 256-sample stages, 8 warps, no producer warp.
 * Each stage is ONE 2-D tensor-map request (TMA, box 64 rows x (40 history + 256) int16) into a
 * two-buffer shared-memory ring, completion on an mbarrier.  Warp 0 issues the first two requests;
 * after that the LAST warp to finish with a buffer (shared-memory arrival counter, acq_rel) issues
 * the request that refills it, so nobody polls.  Every int16 is read from HBM once (the 40-sample
 * overlap of consecutive stages is an L2 hit).  Rows are 592 B (= 80 mod 128) so the 8 lanes of
 * a quarter-warp, which read 8 DIFFERENT rows at the same column, hit 8 distinct 16-byte bank
 * groups with their LDS.128.
 *
 * Work mapping: lane l of warp w computes the 32 outputs of word-column w for channels l and
 * l+32 of the group; the two channels ride in the two halves of fma.rn.f32x2 (SASS FFMA2).
 * Each lane ends with two sign words that go to the [word][channel] buffer as two coalesced
 * 128-byte warp stores.
 *
 * What bounds it (tools/ubench_fir.cu, tools/ubench_ffma2.cu, profiles/r1_ubench_fir.txt): on B200
 * an FFMA2 with three distinct operands occupies the dispatch port for ~2.3 cycles and does not
 * share them with ALU-pipe work (PRMT/LOP3/SHF cost ~1.5 more cycles each next to it), so a word
 * column costs ~1000 cycles per SM sub-partition whatever the order of the instructions: 320 FFMA2
 * + 41 FADD2 (730), the int16 -> f32 conversion (180), the sign / guard bookkeeping (125).  The
 * scalar-FFMA version of the same column is 11 % slower, the symmetric pre-add version 3 % slower.
 *
 * Device sign-word format: LSB first -- bit j of word w = (filtered[32w + j] > 0).
 */
#ifndef GAIS_FIR_CUH
#define GAIS_FIR_CUH

#include <cuda.h>
#include <cudaTypedefs.h>

#include "gais_kernels.cuh"

namespace gais {

constexpr int F_CH = 64;
constexpr int F_T = 256;
constexpr int F_HALO = 40;
constexpr int F_ROW_BYTES = (F_T + F_HALO) * 2;     /* 592 = 80 (mod 128): 8 consecutive rows start in 8 different 16-byte bank groups */
constexpr int F_STAGE_BYTES = F_CH * F_ROW_BYTES;   /* 37888 */
constexpr int F_NSTAGE = 2;
#ifndef F_OUT_WORDS
#define F_OUT_WORDS 1                               /* sign words (32 outputs) per lane per stage */
#endif
#ifndef F_GUARD_GROUP
#define F_GUARD_GROUP 4                            /* outputs per guard-band test (8, 4 or 2) */
#endif
#ifndef F_MIN_BLOCKS
#define F_MIN_BLOCKS 3
#endif
constexpr int F_W = F_OUT_WORDS;
constexpr int F_CWARPS = F_T / (32 * F_W);          /* warps of a CTA: each owns F_W word columns of a stage */
constexpr int F_THREADS = F_CWARPS * 32;
constexpr int F_STAGES_PER_BLOCK = 12;              /* 3072 samples of 64 channels per CTA: shorter-lived CTAs interleave better with the tracker (12: 26.0 ms per step, 16: 26.85, 8: 26.1, 24: 27.5) */
#ifndef F_E1
#define F_E1 0.36f
#endif
#define F_E2_BASE 0.004f       /* 2 * 0.001775 (taps <= 11 / >= 24 at full scale) rounded up */
#define F_E2_SLOPE 3.0e-6f     /* (gamma_32 + 12 u) = 2.64e-6 rounded up */

/* ---- small PTX helpers (all shared-memory operands are 32-bit shared-window addresses) ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"W_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
		     "r"(bytes), "r"(bar)
		     : "memory");
}
/* one TMA request per stage: box = 64 rows x 148 uint32 (= 296 int16) of the planar sample matrix */
__device__ __forceinline__ void tma_g2s_2d(uint32_t dst, const CUtensorMap *tmap, int x, int y, uint32_t bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
		     "l"(tmap), "r"(x), "r"(y), "r"(bar)
		     : "memory");
}
/* "this warp is done reading the buffer": acq_rel, so that the warp which counts the last arrival has
 * every other warp's reads of the buffer ordered before the refill it issues */
__device__ __forceinline__ uint32_t smem_add_acq_rel(uint32_t addr, uint32_t v)
{
	uint32_t old;
	asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
	return old;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}
/* one int16 from shared memory as float, again without the conversion unit (an I2F holds the dispatch
 * port for 8 cycles): the zero-extended halfword, biased and dropped into the mantissa of 2^23 + 2^22 by
 * one XOR, minus 2^23 + 2^22 + 2^15 -- exact */
__device__ __forceinline__ float lds_s16_f32(uint32_t addr)
{
	unsigned short v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
	return __fadd_rn(__uint_as_float((uint32_t) v ^ 0x4B408000u), -12615680.0f);
}

__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
	uint64_t r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b)
{
	uint64_t d;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c)
{
	uint64_t d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b)
{
	uint64_t d;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}

/* ---- tiers 2 and 3 for one doubtful sample; w36 = shared address of x[n-36] inside a row ------ */
__device__ __noinline__ uint32_t fir_sign_resolve(uint32_t w36)
{
	/* tier 2: the 12 taps 12..23 and S12 = sum t_i |x_i|, which makes the bound data dependent:
	 * |a12 - R| <= 0.00355 + 2.64e-6 * S12 (header comment) */
	float xs[GAIS_NTAPS];
#pragma unroll
	for (int i = 12; i <= 23; i++)
		xs[i] = lds_s16_f32(w36 + 2 * i);
	float a = 0.0f, sabs = 0.0f;
#pragma unroll
	for (int i = 12; i <= 23; i++) {
		a = fmaf(xs[i], c_taps[i], a);
		sabs = fmaf(fabsf(xs[i]), c_taps[i], sabs);
	}
	if (fabsf(a) > fmaf(F_E2_SLOPE, sabs, F_E2_BASE))
		return a > 0.0f ? 1u : 0u;
	/* tier 3: the reference's own arithmetic -- float32, multiply then add, tap order
	 * (src/filter.h:40-49).  An all-zero window gives exactly +0: not > 0. */
#pragma unroll
	for (int i = 2; i < 12; i++)
		xs[i] = lds_s16_f32(w36 + 2 * i);
#pragma unroll
	for (int i = 24; i < GAIS_NTAPS - 2; i++)
		xs[i] = lds_s16_f32(w36 + 2 * i);
	float s = 0.0f;
#pragma unroll
	for (int i = 2; i < GAIS_NTAPS - 2; i++)
		s = __fadd_rn(s, __fmul_rn(xs[i], c_taps[i]));
	return s > 0.0f ? 1u : 0u;
}

/*
 * int16 -> float32 without the conversion unit (I2F runs at 16 lanes/clk/SM on the XU pipe,
 * profiles/r1_ubench_b200.txt, and would cap this kernel below the FMA pipe).  The samples
 * are biased to unsigned (x ^ 0x8000), dropped into the low mantissa bytes of 2^23 + 2^22
 * (PRMT), and the constant 2^23 + 2^22 + 2^15 is removed again with one packed add -- every
 * step is exact for 16-bit integers.
 */
#define F_MAGIC_BITS 0x4B400000u
#define F_MAGIC_SUB (-12615680.0f)     /* -(2^23 + 2^22 + 2^15) */

__device__ __forceinline__ uint64_t cvt_pair(uint32_t ua, uint32_t ub, uint32_t sel, uint64_t sub)
{
	/* ua/ub: biased packed int16 pairs of rows A/B; sel picks the low (0x7610) or high (0x7632) half */
	const uint32_t fa = __byte_perm(ua, F_MAGIC_BITS, sel);
	const uint32_t fb = __byte_perm(ub, F_MAGIC_BITS, sel);
	return fadd2(pack2(__uint_as_float(fa), __uint_as_float(fb)), sub);
}

/* 8 samples of rows A and B -> 8 packed (A, B) float pairs */
__device__ __forceinline__ void fir_load_chunk(uint64_t *xs, uint32_t rowA, uint32_t rowB, int byte_ofs, uint64_t sub)
{
	uint4 a = lds128(rowA + byte_ofs);
	uint4 b = lds128(rowB + byte_ofs);
	a.x ^= 0x80008000u; a.y ^= 0x80008000u; a.z ^= 0x80008000u; a.w ^= 0x80008000u;
	b.x ^= 0x80008000u; b.y ^= 0x80008000u; b.z ^= 0x80008000u; b.w ^= 0x80008000u;
	xs[0] = cvt_pair(a.x, b.x, 0x7610, sub);
	xs[1] = cvt_pair(a.x, b.x, 0x7632, sub);
	xs[2] = cvt_pair(a.y, b.y, 0x7610, sub);
	xs[3] = cvt_pair(a.y, b.y, 0x7632, sub);
	xs[4] = cvt_pair(a.z, b.z, 0x7610, sub);
	xs[5] = cvt_pair(a.z, b.z, 0x7632, sub);
	xs[6] = cvt_pair(a.w, b.w, 0x7610, sub);
	xs[7] = cvt_pair(a.w, b.w, 0x7632, sub);
}

/* first stage of a tile (s == 0), whole warp: the row heads come from the carried history
 * (4 zeros + 36 samples, generic stores made visible by the release of the arrive), the 256
 * samples by 64 row copies that leave the heads alone */
__device__ __noinline__ void fir_issue_first(uint8_t *dst, const int16_t *gbase, int64_t ch_stride, const ChanState *st_cg, int hist_sel,
					      uint32_t bar, int lane)
{
	for (int i = lane; i < F_CH * F_HALO; i += 32) {
		const int r = i / F_HALO, k = i % F_HALO;
		const int16_t v = (k < F_HALO - GAIS_NTAPS) ? (int16_t) 0 : st_cg[r].hist[hist_sel][k - (F_HALO - GAIS_NTAPS)];
		*reinterpret_cast<int16_t *>(dst + r * F_ROW_BYTES + k * 2) = v;
	}
	__syncwarp();
	if (lane == 0)
		mbar_expect_tx(bar, (uint32_t) (F_CH * F_T * 2));
	__syncwarp();
#pragma unroll
	for (int h = 0; h < F_CH / 32; h++) {
		const int r = lane + 32 * h;
		bulk_g2s(smem_u32(dst) + r * F_ROW_BYTES + F_HALO * 2, gbase + (int64_t) r * ch_stride, F_T * 2, bar);
	}
}

/* the tile ends in this stage: its last 36 samples (row index 260..295) are the next tile's history
 * (src/filter.c:129-134); written to the other half of the double buffer */
__device__ __noinline__ void fir_save_hist(const uint8_t *stage, ChanState *st_cg, int hist_next, int lane)
{
	for (int i = lane; i < F_CH * GAIS_NTAPS; i += 32) {
		const int r = i / GAIS_NTAPS, k = i % GAIS_NTAPS;
		st_cg[r].hist[hist_next][k] = *reinterpret_cast<const int16_t *>(stage + r * F_ROW_BYTES + (F_HALO + F_T - GAIS_NTAPS + k) * 2);
	}
}

/*
 * No producer warp: the loads of the first F_NSTAGE stages are issued by warp 0 before it starts
 * computing; after that, the LAST warp to finish with a buffer (shared-memory arrival counter)
 * issues the tensor-map request that refills it F_NSTAGE stages ahead.  Nobody polls.
 */
/* refill of one ring buffer, by the one lane that counted the last arrival (out of line: 1 stage in 8
 * per warp takes it, and its ~20 instructions would otherwise be issued, predicated off, by every warp) */
__device__ __noinline__ void fir_issue_refill(uint32_t bar, uint32_t dst, const CUtensorMap *tmap, int x, int y)
{
	mbar_expect_tx(bar, (uint32_t) F_STAGE_BYTES);
	tma_g2s_2d(dst, tmap, x, y, bar);
}

/* a value the compiler must keep in a register instead of recomputing it from special registers */
__device__ __forceinline__ uint32_t pinned_u32(uint32_t v)
{
	uint32_t r;
	asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
	return r;
}

/* DBG: diagnostics only (GAIS_FIR_DBG, separate instantiations -- the product is DBG = 0):
 * 2 loads without compute, 4 compute without refills, 8 guard marks but no tier 2/3, 16 no guard marks */
template <int DBG>
#ifdef F_MAXNREG
__global__ void __maxnreg__(F_MAXNREG)
#else
__global__ void __launch_bounds__(F_THREADS, F_MIN_BLOCKS)
#endif
fir_sign_fast_kernel(const __grid_constant__ CUtensorMap tmap, const int16_t *__restrict__ base, int64_t ch_stride,
		     ChanState *__restrict__ st, int hist_sel, int n_channels, int n_stages, int stages_per_block,
		     uint32_t *__restrict__ signs, int save_hist)
{
	constexpr int dbg = DBG;
	extern __shared__ __align__(128) uint8_t tile[];
	__shared__ __align__(8) uint64_t full_bar_[F_NSTAGE];
	__shared__ uint32_t done_cnt_[F_NSTAGE];

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int cg = blockIdx.x * F_CH;
	const int s_begin = blockIdx.y * stages_per_block;
	const int n_it = min(stages_per_block, n_stages - s_begin);
	const uint32_t tile_a = pinned_u32(smem_u32(tile)), bar_a = pinned_u32(smem_u32(full_bar_)), cnt_a = pinned_u32(smem_u32(done_cnt_));

	if (tid == 0) {
#pragma unroll
		for (int i = 0; i < F_NSTAGE; i++) {
			mbar_init(bar_a + 8 * i, 1);
			done_cnt_[i] = 0;
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (warp == 0) {
		for (int it = 0; it < F_NSTAGE && it < n_it; it++) {
			const int s = s_begin + it;
			if (s == 0)
				fir_issue_first(tile + it * F_STAGE_BYTES, base + (int64_t) cg * ch_stride, ch_stride, st + cg, hist_sel,
						bar_a + 8 * it, lane);
			else if (lane == 0) {
				mbar_expect_tx(bar_a + 8 * it, (uint32_t) F_STAGE_BYTES);
				tma_g2s_2d(tile_a + it * F_STAGE_BYTES, &tmap, (s * F_T - F_HALO) >> 1, cg, bar_a + 8 * it);
			}
		}
	}

	/* ===== every warp: word columns [F_W*warp, F_W*warp + F_W) of each stage; lane l channels l and l+32 ===== */
	uint64_t T[6];      /* taps 12..17 (= 23..18), each in both halves */
#pragma unroll
	for (int k = 0; k < 6; k++)
		T[k] = pack2(c_taps[12 + k], c_taps[12 + k]);
	const uint64_t sub = pack2(F_MAGIC_SUB, F_MAGIC_SUB);
	/* shared address of xs[0] of this lane's row A in buffer 0; output j uses xs[j+1 .. j+10] */
	const uint32_t row0 = pinned_u32(tile_a + lane * F_ROW_BYTES + (32 * F_W * warp + 16) * 2);
	uint32_t *sp = signs + ((int64_t) s_begin * (F_T / 32) + F_W * warp) * n_channels + cg + lane;
	const int64_t sp_step = (int64_t) (F_T / 32) * n_channels;
	int tma_x = ((s_begin + F_NSTAGE) * F_T - F_HALO) >> 1;      /* of the next refill */
	const bool saver = save_hist && s_begin + n_it == n_stages && warp == F_CWARPS - 1;

	for (int it = 0; it < n_it; it++) {
		const uint32_t buf = (uint32_t) it % F_NSTAGE;
		if (!(dbg & 4) || it < F_NSTAGE)      /* GAIS_FIR_DBG=4 (diagnostics only): compute without refills */
			mbar_wait(bar_a + 8 * buf, (uint32_t) ((it / F_NSTAGE) & 1));
		const uint32_t rowA = row0 + buf * F_STAGE_BYTES;
		const uint32_t rowB = rowA + 32 * F_ROW_BYTES;
		uint32_t outA[F_W], outB[F_W];

		if (!(dbg & 2)) {       /* GAIS_FIR_DBG=2 (diagnostics only): loads without compute */
			uint64_t xs[16 + 32 * F_W];
			fir_load_chunk(xs + 0, rowA, rowB, 0, sub);
			fir_load_chunk(xs + 8, rowA, rowB, 16, sub);
#pragma unroll
			for (int wi = 0; wi < F_W; wi++) {
				uint32_t wordA = 0, wordB = 0;        /* sign bits (1 = negative), first sample ends at the MSB */
				uint32_t pendA = 0, pendB = 0;        /* outputs whose sign still needs tiers 2/3 (bit j) */
#pragma unroll
				for (int g4 = 0; g4 < 4; g4++) {
					const int g = 4 * wi + g4;
					fir_load_chunk(xs + 8 * g + 16, rowA, rowB, (8 * g + 16) * 2, sub);
#pragma unroll
					for (int h = 0; h < 8 / F_GUARD_GROUP; h++) {
						/* the guard test is made every F_GUARD_GROUP outputs: the accumulators have to stay in
						 * registers until it is, and the kernel is short of registers */
						uint64_t acc[F_GUARD_GROUP];
						float m = 3.0e38f;
#pragma unroll
						for (int jj = 0; jj < F_GUARD_GROUP; jj++) {
							const int j = 8 * g + F_GUARD_GROUP * h + jj;
							uint64_t a = fmul2(T[1], xs[j + 1]);
							a = ffma2(T[2], xs[j + 2], a);
							a = ffma2(T[3], xs[j + 3], a);
							a = ffma2(T[4], xs[j + 4], a);
							a = ffma2(T[5], xs[j + 5], a);
							a = ffma2(T[5], xs[j + 6], a);
							a = ffma2(T[4], xs[j + 7], a);
							a = ffma2(T[3], xs[j + 8], a);
							a = ffma2(T[2], xs[j + 9], a);
							a = ffma2(T[1], xs[j + 10], a);
							acc[jj] = a;
							float ya, yb;
							unpack2(a, ya, yb);
							m = fminf(m, fminf(fabsf(ya), fabsf(yb)));
							wordA = __funnelshift_l(__float_as_uint(ya), wordA, 1);
							wordB = __funnelshift_l(__float_as_uint(yb), wordB, 1);
						}
						if (m <= F_E1 && !(dbg & 16)) {      /* 16: diagnostics only, no guard marks */
							/* some of these signs are in doubt: only MARK them here (the bits just shifted in for
							 * them are placeholders).  They are settled after the word is done, by compact
							 * out-of-line code, so the unrolled hot path stays small enough for the instruction
							 * cache */
#pragma unroll
							for (int jj = 0; jj < F_GUARD_GROUP; jj++) {
								float ya, yb;
								unpack2(acc[jj], ya, yb);
								if (fabsf(ya) <= F_E1)
									pendA |= 1u << (8 * g4 + F_GUARD_GROUP * h + jj);
								if (fabsf(yb) <= F_E1)
									pendB |= 1u << (8 * g4 + F_GUARD_GROUP * h + jj);
							}
						}
					}
				}
				/* device sign-word format: LSB first, bit j of word w = (filtered[32w + j] > 0) */
				uint32_t oA = __brev(~wordA), oB = __brev(~wordB);
				/* tiers 2 and 3: x[n-36] of output j of this word sits 12 samples before xs[0] + 32*wi + j */
				const uint32_t wA = rowA + (32 * wi - 12) * 2;
				if (dbg & 8)                  /* 8: diagnostics only, marks but no tier 2/3 */
					pendA = pendB = 0;
				while (pendA) {
					const uint32_t j = (uint32_t) __ffs((int) pendA) - 1u;
					pendA &= pendA - 1u;
					oA = (oA & ~(1u << j)) | (fir_sign_resolve(wA + 2 * j) << j);
				}
				while (pendB) {
					const uint32_t j = (uint32_t) __ffs((int) pendB) - 1u;
					pendB &= pendB - 1u;
					oB = (oB & ~(1u << j)) | (fir_sign_resolve(wA + 32 * F_ROW_BYTES + 2 * j) << j);
				}
				outA[wi] = oA;
				outB[wi] = oB;
			}
		} else {
#pragma unroll
			for (int wi = 0; wi < F_W; wi++)
				outA[wi] = outB[wi] = 0u;
		}
		__syncwarp();
		if (lane == 0) {
			/* this warp is done with the buffer; the last one to say so refills it */
			const uint32_t old = smem_add_acq_rel(cnt_a + 4 * buf, 1u);
			if (old % F_CWARPS == F_CWARPS - 1 && it + F_NSTAGE < n_it && !(dbg & 4))
				fir_issue_refill(bar_a + 8 * buf, tile_a + buf * F_STAGE_BYTES, &tmap, tma_x, cg);
		}
		tma_x += F_T / 2;
#pragma unroll
		for (int wi = 0; wi < F_W; wi++) {
			sp[(int64_t) wi * n_channels] = outA[wi];
			sp[(int64_t) wi * n_channels + 32] = outB[wi];
		}
		sp += sp_step;
	}
	/* the tile ends in this CTA's last stage: its last 36 samples are the next tile's history.  Neither of
	 * the last two buffers is refilled, so the stage is still there after the loop */
	if (saver)
		fir_save_hist(tile + ((n_it - 1) % F_NSTAGE) * F_STAGE_BYTES, st + cg, hist_sel ^ 1, lane);
}

/* cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency */
static PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;

static inline int fir_setup(void)
{
	const int smem = F_NSTAGE * F_STAGE_BYTES;
	if (cudaFuncSetAttribute(fir_sign_fast_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
	    cudaFuncSetAttribute(fir_sign_fast_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
	    cudaFuncSetAttribute(fir_sign_fast_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
	    cudaFuncSetAttribute(fir_sign_fast_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
	    cudaFuncSetAttribute(fir_sign_fast_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
	    cudaFuncSetAttribute(fir_sign_fast_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
	    cudaFuncSetAttribute(fir_sign_fast_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
		return -1;
	if (!g_encode_tiled) {
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
			return -1;
		g_encode_tiled = (PFN_cuTensorMapEncodeTiled_v12000) fn;
	}
	return 0;
}

/* tensor map over the planar tile: rows = channels, uint32 elements = int16 pairs */
static inline bool fir_make_tmap(CUtensorMap *tm, const int16_t *base, int64_t ch_stride, int n_rows, int64_t n_frames)
{
	const cuuint64_t gdim[2] = { (cuuint64_t) (n_frames / 2), (cuuint64_t) n_rows };
	const cuuint64_t gstride[1] = { (cuuint64_t) ch_stride * 2 };
	const cuuint32_t box[2] = { (F_T + F_HALO) / 2, F_CH };
	const cuuint32_t estr[2] = { 1, 1 };
	return g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void *) base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} /* namespace gais */
#endif

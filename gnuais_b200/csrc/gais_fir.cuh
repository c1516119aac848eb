/*
 * gais_fir.cuh -- K1, the FIR-sign stage, and its launch logic.
 *
 * What must come out is sign-exact: cur = (filter_run_buf() output > 0) for every sample, with
 * the reference's float32 rounding (src/filter.h:40-49 sequential mul+add, src/filter.c:115-125
 * window = the 36 samples before the current one, src/receiver.c:110-111 only the sign is used).
 *
 * Fast kernel (sm_100a): "guard-banded" evaluation.
 *   tier 1   a10 = sum over the 10 centre taps 13..22 (packed fp32x2 FMAs, any rounding).
 *            |a10 - R| <= 0.4493 for ANY int16 input, R being the reference's rounded sum:
 *              0.1563  gamma_32 * sum|t_i x_i|   (reference's own rounding, sum t_i = 2.50001, |x| <= 32768)
 *              0.2444  2 * (t_12 + t_11 + ...) * 32768        (dropped taps)
 *              0.0488  10 roundings of the FMA chain
 *            so |a10| > E1 = 0.5  =>  sign(R) = sign(a10) and R != 0.
 *   tier 2   (only samples with |a10| <= E1, ~5e-4 of noisy audio) 12 taps 12..23, in registers,
 *            with a DATA-DEPENDENT bound: S12 = sum_{12 taps} t_i |x_i| is computed alongside and
 *              |a12 - R| <= gamma_32 * (S12 + 0.001775) + 0.001775 + 12 u S12
 *                        <= 0.00355 + 2.64e-6 * S12          (u = 2^-24; 0.001775 = outer taps at full scale)
 *            -> E2 = 0.004 + 3.0e-6 * S12 (0.006 on idle noise, 0.09 inside a burst).
 *   tier 3   (|a12| <= E2, ~1e-5 of samples) the exact 32-term chain, __fmul_rn/__fadd_rn in tap order.
 *
 * Data movement: one CTA = 64 channels x a run of 256-sample stages.  Each stage is 64 row
 * segments of (40 history + 256) int16 brought in by cp.async.bulk (TMA 1-D) into a
 * double-buffered shared-memory tile, completion on an mbarrier; every int16 is read from HBM
 * once (the 40-sample overlap of consecutive stages is an L2 hit).  Rows are padded to
 * 656 B (= 16 mod 128) so the 8 lanes of a quarter-warp, which read 8 DIFFERENT rows at the
 * same column, cover all 32 banks with their LDS.128.
 *
 * Work mapping: lane l of warp w computes the 32 outputs of word-column w for channels l and
 * l+32 of the group; the two channels ride in the two halves of fma.rn.f32x2 (SASS FFMA2: 2
 * MACs per issue slot on the heavy FMA pipe, profiles/r1_ubench_b200.txt).  Each lane ends with
 * two sign words that go to the [word][channel] buffer as two coalesced 128-byte warp stores.
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
constexpr int F_CWARPS = 8;                         /* consumer warps = word columns of a stage */
constexpr int F_THREADS = (F_CWARPS + 1) * 32;      /* + one producer warp */
constexpr int F_STAGES_PER_BLOCK = 16;              /* 4096 samples of 64 channels per CTA */
#define F_E1 0.5f
#define F_E2_BASE 0.004f       /* 2 * 0.001775 (taps <= 11 / >= 24 at full scale) rounded up */
#define F_E2_SLOPE 3.0e-6f     /* (gamma_32 + 12 u) = 2.64e-6 rounded up */

/* ---- small PTX helpers --------------------------------------------------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"W_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@!p bra W_%=;\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* producer side: poll politely, the consumers need the issue slots */
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity)
{
	uint32_t done = 0;
	for (;;) {
		asm volatile("{\n\t.reg .pred p;\n\t"
			     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		if (done)
			break;
		__nanosleep(256);
	}
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
		     "l"(src), "r"(bytes), "r"(smem_u32(bar))
		     : "memory");
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

/* ---- tiers 2 and 3 for one doubtful sample; w36 -> x[n-36] inside a shared-memory row ------ */
__device__ __noinline__ bool fir_sign_resolve(const int16_t *w36)
{
	/* tier 2: the 12 taps 12..23 and S12 = sum t_i |x_i|, which makes the bound data dependent:
	 * |a12 - R| <= 0.00355 + 2.64e-6 * S12 (header comment) */
	float xs[GAIS_NTAPS];
#pragma unroll
	for (int i = 12; i <= 23; i++)
		xs[i] = (float) w36[i];
	float a = 0.0f, sabs = 0.0f;
#pragma unroll
	for (int i = 12; i <= 23; i++) {
		a = fmaf(xs[i], c_taps[i], a);
		sabs = fmaf(fabsf(xs[i]), c_taps[i], sabs);
	}
	if (fabsf(a) > fmaf(F_E2_SLOPE, sabs, F_E2_BASE))
		return a > 0.0f;
	/* tier 3: the reference's own arithmetic -- float32, multiply then add, tap order
	 * (src/filter.h:40-49).  An all-zero window gives exactly +0: not > 0. */
#pragma unroll
	for (int i = 2; i < 12; i++)
		xs[i] = (float) w36[i];
#pragma unroll
	for (int i = 24; i < GAIS_NTAPS - 2; i++)
		xs[i] = (float) w36[i];
	float s = 0.0f;
#pragma unroll
	for (int i = 2; i < GAIS_NTAPS - 2; i++)
		s = __fadd_rn(s, __fmul_rn(xs[i], c_taps[i]));
	return s > 0.0f;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint64_t abs2(uint64_t a) { return a & 0x7fffffff7fffffffull; }

__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b)
{
	uint64_t d;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
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
#ifdef GAIS_FIR_SCALAR_CVT
	(void) sub;
	return pack2(__fadd_rn(__uint_as_float(fa), F_MAGIC_SUB), __fadd_rn(__uint_as_float(fb), F_MAGIC_SUB));
#else
	return fadd2(pack2(__uint_as_float(fa), __uint_as_float(fb)), sub);
#endif
}

/* 8 samples of rows A and B -> 8 packed (A, B) float pairs */
__device__ __forceinline__ void fir_load_chunk(uint64_t *xs, const uint8_t *rowA, const uint8_t *rowB, int byte_ofs, uint64_t sub)
{
	uint4 a = *reinterpret_cast<const uint4 *>(rowA + byte_ofs);
	uint4 b = *reinterpret_cast<const uint4 *>(rowB + byte_ofs);
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

/* one TMA request per stage: box = 64 rows x 148 uint32 (= 296 int16) of the planar sample matrix */
__device__ __forceinline__ void tma_g2s_2d(void *dst, const CUtensorMap *tmap, int x, int y, uint64_t *bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
			     smem_u32(dst)),
		     "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
		     : "memory");
}

__global__ void __launch_bounds__(F_THREADS, 3)
fir_sign_fast_kernel(const __grid_constant__ CUtensorMap tmap, const int16_t *__restrict__ base, int64_t ch_stride,
		     ChanState *__restrict__ st, int hist_sel, int n_channels, int n_stages, int stages_per_block,
		     uint32_t *__restrict__ signs, int dbg, int save_hist)
{
	extern __shared__ __align__(128) uint8_t tile[];
	__shared__ __align__(8) uint64_t full_bar[F_NSTAGE], empty_bar[F_NSTAGE];

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int cg = blockIdx.x * F_CH;
	const int s_begin = blockIdx.y * stages_per_block;
	const int s_end = min(s_begin + stages_per_block, n_stages);

	if (tid == 0) {
		for (int i = 0; i < F_NSTAGE; i++) {
			mbar_init(&full_bar[i], 1);
			mbar_init(&empty_bar[i], F_CWARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	if (warp == F_CWARPS) {
		/* ===== producer warp: lane l brings rows l and l+32 of every stage ===== */
		const int16_t *gbase = base + (int64_t) cg * ch_stride;
		for (int s = s_begin; s < s_end; s++) {
			const int it = s - s_begin, buf = it % F_NSTAGE;
			if (it >= F_NSTAGE)
				mbar_wait_sleep(&empty_bar[buf], (uint32_t) ((it / F_NSTAGE - 1) & 1));
			uint8_t *dst = tile + buf * F_STAGE_BYTES;
			const int64_t n0 = (int64_t) s * F_T;
			if (s == 0) {
				/* no samples before the run: rows start with 4 zeros + the 36 carried samples
				 * (generic stores; made visible to the consumers by the release of the arrive below) */
				for (int i = lane; i < F_CH * F_HALO; i += 32) {
					const int r = i / F_HALO, k = i % F_HALO;
					const int16_t v = (k < F_HALO - GAIS_NTAPS) ? (int16_t) 0 : st[cg + r].hist[hist_sel][k - (F_HALO - GAIS_NTAPS)];
					*reinterpret_cast<int16_t *>(dst + r * F_ROW_BYTES + k * 2) = v;
				}
			}
			__syncwarp();
			if (dbg & 1) {          /* diagnostics only (GAIS_FIR_DBG=1): no loads, compute on stale smem */
				if (lane == 0)
					mbar_expect_tx(&full_bar[buf], 0);
			} else if (s == 0) {
				/* first stage of a tile: 64 row copies that leave the history bytes alone */
				if (lane == 0)
					mbar_expect_tx(&full_bar[buf], (uint32_t) (F_CH * F_T * 2));
				__syncwarp();
#pragma unroll
				for (int h = 0; h < 2; h++) {
					const int r = lane + 32 * h;
					bulk_g2s(dst + r * F_ROW_BYTES + F_HALO * 2, gbase + (int64_t) r * ch_stride + n0, F_T * 2, &full_bar[buf]);
				}
			} else if (lane == 0) {
				/* every other stage: ONE tensor-map request for the whole 64 x 296 box */
				mbar_expect_tx(&full_bar[buf], (uint32_t) F_STAGE_BYTES);
				tma_g2s_2d(dst, &tmap, (int) ((n0 - F_HALO) >> 1), cg, &full_bar[buf]);
			}
		}
		return;
	}

	/* ===== consumer warps: warp w owns word column w; lane l channels l and l+32 ===== */
	uint64_t T[6];      /* taps 12..17 (= 23..18), each in both halves */
#pragma unroll
	for (int k = 0; k < 6; k++)
		T[k] = pack2(c_taps[12 + k], c_taps[12 + k]);
	const uint64_t sub = pack2(F_MAGIC_SUB, F_MAGIC_SUB);

	for (int s = s_begin; s < s_end; s++) {
		const int it = s - s_begin, buf = it % F_NSTAGE;
		mbar_wait(&full_bar[buf], (uint32_t) ((it / F_NSTAGE) & 1));
		const uint8_t *stage = tile + buf * F_STAGE_BYTES;
		const uint8_t *rowA = stage + lane * F_ROW_BYTES;
		const uint8_t *rowB = stage + (lane + 32) * F_ROW_BYTES;
		const int col0 = 32 * warp + 16;      /* row index of xs[0]; output j uses xs[j+1 .. j+10] */
		if (dbg & 2) {          /* diagnostics only (GAIS_FIR_DBG=2): loads without compute */
			__syncwarp();
			if (lane == 0)
				mbar_arrive(&empty_bar[buf]);
			continue;
		}

		uint64_t xs[48];
		fir_load_chunk(xs + 0, rowA, rowB, (col0 + 0) * 2, sub);
		fir_load_chunk(xs + 8, rowA, rowB, (col0 + 8) * 2, sub);

		uint32_t wordA = 0, wordB = 0;        /* sign bits (1 = negative), first sample ends at the MSB */
		uint32_t pendA = 0, pendB = 0;        /* outputs whose sign still needs the exact chain (bit j) */
#pragma unroll
		for (int g = 0; g < 4; g++) {
			fir_load_chunk(xs + 8 * g + 16, rowA, rowB, (col0 + 8 * g + 16) * 2, sub);
			uint64_t acc[8];
			float m = 3.0e38f;
#pragma unroll
			for (int jj = 0; jj < 8; jj++) {
				const int j = 8 * g + jj;
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
			if (m <= F_E1) {
				/* some of these 16 signs are in doubt: only MARK them here (the bits just shifted in for
				 * them are placeholders).  They are settled after the column is done, by compact
				 * out-of-line code, so the unrolled hot path stays small enough for the instruction cache */
#pragma unroll
				for (int jj = 0; jj < 8; jj++) {
					float ya, yb;
					unpack2(acc[jj], ya, yb);
					if (fabsf(ya) <= F_E1)
						pendA |= 1u << (8 * g + jj);
					if (fabsf(yb) <= F_E1)
						pendB |= 1u << (8 * g + jj);
				}
			}
		}
		/* device sign-word format: LSB first, bit j of word w = (filtered[32w + j] > 0) */
		uint32_t outA = __brev(~wordA), outB = __brev(~wordB);
		while (pendA | pendB) {      /* tiers 2 and 3: x[n-36] of output j sits at row index 32*warp + j + 4 */
			const bool isA = pendA != 0u;
			uint32_t &pend = isA ? pendA : pendB;
			const int j = __ffs((int) pend) - 1;
			pend &= pend - 1u;
			const bool pos = fir_sign_resolve(reinterpret_cast<const int16_t *>(isA ? rowA : rowB) + 32 * warp + j + 4);
			uint32_t &out = isA ? outA : outB;
			out = (out & ~(1u << j)) | ((pos ? 1u : 0u) << j);
		}
		if (save_hist && s == n_stages - 1 && warp == F_CWARPS - 1) {
			/* the tile ends here: its last 36 samples (row index 260..295) are the next tile's history
			 * (src/filter.c:129-134); written to the other half of the double buffer */
			for (int i = lane; i < F_CH * GAIS_NTAPS; i += 32) {
				const int r = i / GAIS_NTAPS, k = i % GAIS_NTAPS;
				st[cg + r].hist[hist_sel ^ 1][k] =
					*reinterpret_cast<const int16_t *>(stage + r * F_ROW_BYTES + (F_HALO + F_T - GAIS_NTAPS + k) * 2);
			}
		}
		__syncwarp();
		if (lane == 0)
			mbar_arrive(&empty_bar[buf]);     /* this warp is done with the buffer */
		const int64_t wrow = (int64_t) s * (F_T / 32) + warp;
		signs[wrow * n_channels + cg + lane] = outA;
		signs[wrow * n_channels + cg + lane + 32] = outB;
	}
}

/* cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda dependency */
static PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;

static inline int fir_setup(void)
{
	if (cudaFuncSetAttribute(fir_sign_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_NSTAGE * F_STAGE_BYTES) != cudaSuccess)
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

/*
 * Launch K1 for one time tile.  The fast kernel takes the part of the tile it is built for
 * (planar rows, 16-byte aligned, whole 64-channel groups, whole 256-sample stages); the exact
 * kernel sweeps up the ragged remainder (and everything in GAIS_FIR_EXACT mode).
 * Returns the number of kernels launched, < 0 on error.
 */
static inline int fir_launch(int fir_mode, int layout, SampleView view, ChanState *st, int hist_sel, int n_ch,
			     int64_t n_frames, uint32_t *signs, cudaStream_t stream, int *hist_saved_channels)
{
	int launches = 0;
	int fast_ch = 0;
	*hist_saved_channels = 0;
	int64_t fast_frames = 0;
	const bool aligned = layout == GAIS_LAYOUT_PLANAR && view.t_stride == 1 && (view.ch_stride % 8) == 0 &&
			     ((uintptr_t) view.base % 16) == 0;
	if (fir_mode == GAIS_FIR_GUARD && aligned) {
		fast_ch = n_ch / F_CH * F_CH;
		fast_frames = n_frames / F_T * F_T;
	}
	if (fast_ch > 0 && fast_frames > 0) {
		const int n_stages = (int) (fast_frames / F_T);
		static int spb = 0, dbg = 0;
		if (!spb) {
			const char *e = getenv("GAIS_FIR_SPB");
			spb = (e && atoi(e) > 0) ? atoi(e) : F_STAGES_PER_BLOCK;
			e = getenv("GAIS_FIR_DBG");
			dbg = e ? atoi(e) : 0;
		}
		dim3 grid((unsigned) (fast_ch / F_CH), (unsigned) ((n_stages + spb - 1) / spb));
		CUtensorMap tm;
		if (!fir_make_tmap(&tm, view.base, view.ch_stride, fast_ch, fast_frames))
			return -1;
		fir_sign_fast_kernel<<<grid, F_THREADS, F_NSTAGE * F_STAGE_BYTES, stream>>>(tm, view.base, view.ch_stride, st, hist_sel, n_ch,
											     n_stages, spb, signs, dbg, fast_frames == n_frames ? 1 : 0);
		*hist_saved_channels = (fast_frames == n_frames) ? fast_ch : 0;
		launches++;
	} else {
		fast_ch = 0;
		fast_frames = 0;
	}
	/* remainder in time for the fast channels: frames [fast_frames, n_frames) */
	if (fast_ch > 0 && fast_frames < n_frames) {
		dim3 grid((unsigned) ((fast_ch + K1_CH - 1) / K1_CH), (unsigned) ((n_frames - fast_frames + K1_TILE - 1) / K1_TILE));
		fir_sign_exact_kernel<<<grid, K1_CH * 32, 0, stream>>>(view, st, hist_sel, 0, fast_ch, fast_frames, n_frames, n_ch, signs);
		launches++;
	}
	/* remaining channels, all frames */
	if (fast_ch < n_ch) {
		dim3 grid((unsigned) ((n_ch - fast_ch + K1_CH - 1) / K1_CH), (unsigned) ((n_frames + K1_TILE - 1) / K1_TILE));
		fir_sign_exact_kernel<<<grid, K1_CH * 32, 0, stream>>>(view, st, hist_sel, fast_ch, n_ch, 0, n_frames, n_ch, signs);
		launches++;
	}
	return cudaGetLastError() == cudaSuccess ? launches : -1;
}

} /* namespace gais */
#endif

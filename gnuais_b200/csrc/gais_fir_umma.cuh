/*
 * gais_fir_umma.cuh -- K1 on the tensor pipe: the FIR-sign stage as an EXACT integer Toeplitz
 * contraction on tcgen05.mma.kind::i8 (SASS UTCIMMA), accumulators in TMEM.
 *
 * What must come out is the same as in gais_fir.cuh: cur = (filter_run_buf() output > 0) for every
 * sample, with the reference's float32 rounding (src/filter.h:40-49 sequential mul+add,
 * src/filter.c:115-125 window = the 36 samples before the current one, src/receiver.c:110-111 only
 * the sign is used).  The FFMA2 kernel of gais_fir.cuh spends 10 packed FMAs per output on the
 * FP32 pipe and is bound by instruction dispatch (profiles/r1_ncu_final_summary.txt); this kernel
 * moves the multiply-adds off the issue port altogether.
 *
 * Arithmetic.  Only the taps 12..23 are >= 2^-25 (the other 20 non-zero taps add up to 0.001775 at
 * full scale); each is quantised to T_i = round(t_i * 2^24) = 2^16 T1 + 2^8 T2 + T3 (bytes).  A sample
 * with bit 7 flipped, x ^ 0x0080, read as two SIGNED bytes (hi, lo') satisfies x = 256 hi + lo' + 128
 * exactly, so
 *     sum_i T_i x_i = 2^24 D24 + 2^16 D16 + 2^8 D8 + D0 + 128 sum T
 *     D24 = sum T1 hi        D16 = sum (T2 hi + T1 lo')        D8 = sum (T3 hi + T2 lo')        D0 = sum T3 lo'
 * D24, D16, D8 are three int32 accumulators of ONE MMA (N = 3 x 32 columns), exact by construction --
 * integer arithmetic has no rounding to model.  With V = 65536 D24 + 256 D16 + D8 + sum T / 2 (units of
 * 2^-16) and R the reference's rounded float32 sum,
 *     |V / 65536 - R| <= 0.0855 (the reference's own rounding, gais_fir.cuh header)
 *                      + 0.0018 (taps outside 12..23 at full scale) + 0.0045 (tap quantisation)
 *                      + 0.0112 (D0, |lo'| <= 128)                                  = 0.103 < 0.5,
 * so Q = round(V / 65536) decides: Q >= 1 -> R > 0, Q <= -1 -> R < 0, and Q == 0 (1e-3 of noisy audio)
 * goes to tiers 2/3 of gais_fir.cuh (12-tap FMA with a data-dependent bound, then the exact chain).
 * tools/ubench_umma_probe.cu pins the operand addressing this relies on; tests/test_fir_umma_model.py
 * replays the integer arithmetic in numpy against the reference's float32 chain.
 *
 * Operands.  A is NOT materialised: the 16 channel rows of a stage (40 history + 256 samples, int16,
 * bit 7 flipped) sit in shared memory in time order, stored through the 64-byte swizzle, and the A
 * descriptor describes OVERLAPPING rows -- MMA row m = 8 c + r is the 96-byte window that starts at
 * byte 32 + 64 r of channel c's row (row pitch 64 B = SWIZZLE_64B, 8-row groups 592 B apart = SBO), k-step
 * k starts 32 k bytes further.  Three MMAs (M = 128, N = 96, K = 32 bytes = 16 samples) give the 32
 * outputs of every 32-sample word of 16 channels x 256 samples: TMEM lane m, column 32 a + j.
 * B is the banded tap matrix, built once on the host: B[32 a + j][2 s + b] with i = s - j + 12.
 *
 * Data movement: global -> registers (coalesced 16-byte loads, prefetched one stage ahead) -> XOR
 * 0x0080 -> swizzled STS.  No TMA here on purpose: the sign flip needs a pass through registers anyway,
 * and the kernel's next bound after instruction issue is shared-memory bandwidth, which a
 * TMA write followed by an LDS/STS pass would double.
 */
#ifndef GAIS_FIR_UMMA_CUH
#define GAIS_FIR_UMMA_CUH

#include "gais_fir.cuh"

namespace gais {

constexpr int U_CH = 16;                               /* channels per CTA: 16 x 8 words = 128 MMA rows */
constexpr int U_T = 256;                               /* samples per stage */
constexpr int U_HALO = 40;
constexpr int U_ROW_BYTES = (U_HALO + U_T) * 2;        /* 592 */
constexpr int U_ROW_CHUNKS = U_ROW_BYTES / 16;         /* 37 */
constexpr int U_BUF_BYTES = U_CH * U_ROW_BYTES;        /* 9472 */
constexpr int U_BUF_STRIDE = (U_BUF_BYTES + 511) / 512 * 512;   /* 9728: a multiple of the swizzle period, so that
                                                          swz64(a + buf * stride) = swz64(a) + buf * stride */
constexpr int U_THREADS = 128;
constexpr int U_LD = 2 * ((U_ROW_CHUNKS + 7) / 8);      /* 16-byte loads per thread per stage it moves: 64 threads, 2 rows x 5 chunks each */
constexpr int U_BK_BYTES = 96 * 32;                    /* one k-step of B: 96 rows x 32 bytes */
constexpr int U_BMAT_BYTES = 3 * U_BK_BYTES;           /* 9216 */
constexpr int U_SMEM_BYTES = U_BMAT_BYTES + 2 * U_BUF_STRIDE + 1024;   /* + alignment slack */
constexpr int U_SMEM_REQUEST = 50 * 1024;              /* pins residency at 4 CTAs per SM = 4 x 128 TMEM columns */
constexpr int U_TMEM_COLS = 128;
#ifndef U_STAGES_PER_BLOCK
#define U_STAGES_PER_BLOCK 48
#endif
constexpr int U_TAP_LO = 12, U_TAP_HI = 23;            /* taps with T_i != 0 */
constexpr int U_VGUARD = 8192;                         /* 0.125 in units of 2^-16: above the 0.103 error bound of the header */

/* instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = s8, B = u8, both K-major, N = 96, M = 128 */
constexpr uint32_t U_IDESC = (2u << 4) | (1u << 7) | (0u << 10) | ((96u >> 3) << 17) | ((128u >> 4) << 24);

__device__ uint8_t g_umma_bmat[U_BMAT_BYTES];

/* tap bytes, B matrix image and the rounding constant, on the host */
struct UmmaTaps {
	uint8_t bmat[U_BMAT_BYTES];
	int32_t kc;            /* sum T / 2 + 32768: V + 32768 = 65536 D24 + 256 D16 + D8 + kc */
};

static inline void umma_build_taps(UmmaTaps *out)
{
	static const uint32_t half[18] = GAIS_TAP_BITS_HALF;
	int64_t T[GAIS_NTAPS], sum = 0;
	for (int i = 0; i < GAIS_NTAPS; i++) {
		const uint32_t b = half[i < 18 ? i : 35 - i];
		float t;
		memcpy(&t, &b, 4);
		T[i] = (i >= U_TAP_LO && i <= U_TAP_HI) ? (int64_t) ((double) t * 16777216.0 + 0.5) : 0;
		sum += T[i];
	}
	memset(out->bmat, 0, sizeof(out->bmat));
	for (int a = 0; a < 3; a++)
		for (int j = 0; j < 32; j++)
			for (int s = 0; s < 48; s++)
				for (int b = 0; b < 2; b++) {
					const int i = s - j + 12;
					if (i < U_TAP_LO || i > U_TAP_HI)
						continue;
					const int T1 = (int) ((T[i] >> 16) & 255), T2 = (int) ((T[i] >> 8) & 255), T3 = (int) (T[i] & 255);
					int v;
					if (a == 0) v = b ? T1 : 0;          /* D24 = sum T1 hi */
					else if (a == 1) v = b ? T2 : T1;    /* D16 = sum T2 hi + T1 lo' */
					else v = b ? T3 : T2;                /* D8  = sum T3 hi + T2 lo' */
					/* canonical K-major no-swizzle layout per k-step: 8-row groups 256 B apart, the two 16-byte
					 * K chunks of a group 128 B apart */
					const int n = 32 * a + j, k = 2 * s + b, ks = k / 32, kk = k % 32;
					out->bmat[ks * U_BK_BYTES + (n / 8) * 256 + (kk / 16) * 128 + (n % 8) * 16 + kk % 16] = (uint8_t) v;
				}
	out->kc = (int32_t) (sum / 2 + 32768);
}

/* shared-memory matrix descriptor (cute::UMMA::SmemDescriptor) without the start address */
__host__ __device__ constexpr uint64_t umma_desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type)
{
	return ((uint64_t) ((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t) ((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t) 1 << 46) |
	       ((uint64_t) (layout_type & 7u) << 61);
}
constexpr uint64_t U_ADESC = umma_desc_base(16, U_ROW_BYTES, 4);    /* SWIZZLE_64B, 8-row groups = channels, 592 B apart */
constexpr uint64_t U_BDESC = umma_desc_base(128, 256, 0);           /* no swizzle */

/* the 64-byte swizzle is a function of ABSOLUTE shared-memory address bits: 16-byte chunk index (bits 4-5)
 * ^= bits 7-8 (profiles/r2_umma_probe.txt) */
__device__ __forceinline__ uint32_t swz64(uint32_t a) { return a ^ (((a >> 7) & 3u) << 4); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
		  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
		: "r"(taddr));
}

/* one int16 of a swizzled, bit-7-flipped row as float (same exact trick as lds_s16_f32) */
__device__ __forceinline__ float umma_lds_f32(uint32_t lin)
{
	unsigned short v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(swz64(lin)));
	return __fadd_rn(__uint_as_float((uint32_t) v ^ 0x4B408080u), -12615680.0f);
}

/* tiers 2 and 3 for one doubtful output (gais_fir.cuh fir_sign_resolve, on the swizzled row);
 * w36 = linear shared address of x[n-36] */
__device__ __noinline__ uint32_t umma_sign_resolve(uint32_t w36)
{
	float xs[GAIS_NTAPS];
#pragma unroll
	for (int i = 12; i <= 23; i++)
		xs[i] = umma_lds_f32(w36 + 2 * i);
	float a = 0.0f, sabs = 0.0f;
#pragma unroll
	for (int i = 12; i <= 23; i++) {
		a = fmaf(xs[i], c_taps[i], a);
		sabs = fmaf(fabsf(xs[i]), c_taps[i], sabs);
	}
	if (fabsf(a) > fmaf(F_E2_SLOPE, sabs, F_E2_BASE))
		return a > 0.0f ? 1u : 0u;
	/* tier 3: the reference's own arithmetic (src/filter.h:40-49) */
#pragma unroll
	for (int i = 2; i < 12; i++)
		xs[i] = umma_lds_f32(w36 + 2 * i);
#pragma unroll
	for (int i = 24; i < GAIS_NTAPS - 2; i++)
		xs[i] = umma_lds_f32(w36 + 2 * i);
	float s = 0.0f;
#pragma unroll
	for (int i = 2; i < GAIS_NTAPS - 2; i++)
		s = __fadd_rn(s, __fmul_rn(xs[i], c_taps[i]));
	return s > 0.0f ? 1u : 0u;
}

/* 16 outputs from three accumulator slices.  rev collects the sign bits of Q (output 0 ends at bit 15);
 * the rare Q == 0 outputs are looked at again with the exact V = (p & 0xffff) - 32768: V <= -8192 is a
 * certain negative (clr), |V| < 8192 stays open (pend), V >= 8192 is a certain positive.  The unrolled
 * scan is only entered by warps in which some lane saw a zero, one 4-output chain at a time. */
__device__ __forceinline__ void umma_half_word(const uint32_t (&d24)[16], const uint32_t (&d16)[16], const uint32_t (&d8)[16], int kc,
					       uint32_t &neg, uint32_t &clr, uint32_t &pend)
{
	int q[16];
	uint32_t mn[4], rev = 0;
#pragma unroll
	for (int j = 0; j < 16; j++) {
		const int p = (int) d16[j] * 256 + (int) d8[j] + kc;
		q[j] = (int) d24[j] + (p >> 16);
		rev = __funnelshift_l((uint32_t) q[j], rev, 1);
	}
	neg = __brev(rev) >> 16;
#pragma unroll
	for (int c = 0; c < 4; c++)
		mn[c] = min(min((uint32_t) q[4 * c], (uint32_t) q[4 * c + 1]), min((uint32_t) q[4 * c + 2], (uint32_t) q[4 * c + 3]));
	pend = 0;
	clr = 0;
	if (min(min(mn[0], mn[1]), min(mn[2], mn[3])) == 0u) {
#pragma unroll
		for (int c = 0; c < 4; c++)
			if (mn[c] == 0u) {
#pragma unroll
				for (int j = 4 * c; j < 4 * c + 4; j++)
					if (q[j] == 0) {
						const int v = (((int) d16[j] * 256 + (int) d8[j] + kc) & 0xffff) - 32768;
						if (v <= -U_VGUARD)
							clr |= 1u << j;
						else if (v < U_VGUARD)
							pend |= 1u << j;
					}
			}
	}
}

/* ---- the open outputs (2.6e-4 of noisy audio) are not settled where they are found -- one lane of a warp
 * walking tiers 2/3 costs the whole warp ~200 issue slots -- but queued per CTA and settled 32+ at a time,
 * one per lane, from global memory; the sign word already written carries a provisional 1 that is cleared
 * with an atomic when the exact answer is "not > 0". ---- */
constexpr int U_QCAP = 128, U_QDRAIN = 64;

/* sample t of a channel row of the tile (t < 0: the carried history, src/filter.c:129-134) as float */
__device__ __forceinline__ float umma_sample(const int16_t *__restrict__ row, const int16_t *__restrict__ hist, int t)
{
	return (float) (t >= 0 ? row[t] : (t >= -GAIS_NTAPS ? hist[GAIS_NTAPS + t] : (int16_t) 0));
}

/* tiers 2 and 3 of gais_fir.cuh for output n of a channel, read from global memory */
__device__ __noinline__ uint32_t umma_resolve_global(const int16_t *__restrict__ row, const int16_t *__restrict__ hist, int n)
{
	float xs[GAIS_NTAPS];
#pragma unroll
	for (int i = 2; i < GAIS_NTAPS - 2; i++)
		xs[i] = umma_sample(row, hist, n - GAIS_NTAPS + i);
	float a = 0.0f, sabs = 0.0f;
#pragma unroll
	for (int i = 12; i <= 23; i++) {
		a = fmaf(xs[i], c_taps[i], a);
		sabs = fmaf(fabsf(xs[i]), c_taps[i], sabs);
	}
	if (fabsf(a) > fmaf(F_E2_SLOPE, sabs, F_E2_BASE))
		return a > 0.0f ? 1u : 0u;
	float s = 0.0f;            /* tier 3: the reference's own arithmetic (src/filter.h:40-49) */
#pragma unroll
	for (int i = 2; i < GAIS_NTAPS - 2; i++)
		s = __fadd_rn(s, __fmul_rn(xs[i], c_taps[i]));
	return s > 0.0f ? 1u : 0u;
}

/* queue the open outputs of one sign word; when the queue is full they are settled on the spot from the
 * shared-memory row.  Returns the bits to clear in that case (bits 0..15), bit 16 = something was queued,
 * bit 17 = the queue has reached the length at which the CTA settles it */
__device__ __noinline__ uint32_t umma_push(uint32_t pend, uint32_t item0, uint32_t qn_a, uint32_t q_a, uint32_t w36_0)
{
	uint32_t clr = 0;
	while (pend) {
		const uint32_t j = (uint32_t) __ffs((int) pend) - 1u;
		pend &= pend - 1u;
		uint32_t pos;
		asm volatile("atom.shared::cta.add.u32 %0, [%1], 1;" : "=r"(pos) : "r"(qn_a) : "memory");
		if (pos < (uint32_t) U_QCAP) {
			asm volatile("st.shared.u32 [%0], %1;" ::"r"(q_a + 4u * pos), "r"(item0 + j) : "memory");
			clr |= 1u << 16;
		} else if (umma_sign_resolve(w36_0 + 2u * j) == 0u)
			clr |= 1u << j;
		if (pos + 1u >= (uint32_t) U_QDRAIN)
			clr |= 1u << 17;
	}
	return clr;
}

struct UmmaRegs { uint4 v[U_LD]; };      /* the 16-byte chunks of a stage a thread moves (see the loader in the kernel) */

/* first stage of a tile: the 40 samples before the tile come from the carried history (4 zeros + 36 samples,
 * src/filter.c:57-71 / :129-134) -- chunks 0..4 of each row, i.e. k == 0 of the threads with tid % 8 < 5 */
__device__ __noinline__ uint4 umma_hist_chunk(const ChanState *__restrict__ st_row, int hist_sel, int ch)
{
	uint32_t w[4];
#pragma unroll
	for (int e = 0; e < 4; e++) {
		const int t0 = 8 * ch + 2 * e - (U_HALO - GAIS_NTAPS);      /* index into hist[] */
		const uint32_t lo = t0 >= 0 ? (uint16_t) st_row->hist[hist_sel][t0] : 0u;
		const uint32_t hi = t0 + 1 >= 0 ? (uint16_t) st_row->hist[hist_sel][t0 + 1] : 0u;
		w[e] = lo | (hi << 16);
	}
	return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(U_THREADS, 4)
fir_sign_umma_kernel(const int16_t *__restrict__ base, int64_t ch_stride, ChanState *__restrict__ st, int hist_sel, int n_channels,
		     int n_stages, int stages_per_block, uint32_t *__restrict__ signs, int save_hist, int kc, int dbg)
{
	extern __shared__ __align__(1024) uint8_t u_smem_raw[];
	__shared__ __align__(8) uint64_t mma_bar;
	__shared__ uint32_t tmem_base_s;
	__shared__ uint32_t q_n, q_item[U_QCAP];      /* open outputs: (channel of the group << 28) | sample index in the tile */

	const int tid = threadIdx.x, warp = tid >> 5;
	const int cg = blockIdx.x * U_CH;
	const int s_begin = blockIdx.y * stages_per_block;
	const int n_it = min(stages_per_block, n_stages - s_begin);
	const uint32_t smem0 = (smem_u32(u_smem_raw) + 1023u) & ~1023u;
	const uint32_t bmat_a = smem0, buf_a = smem0 + U_BMAT_BYTES;
	const uint32_t bar_a = smem_u32(&mma_bar);

	/* ---- one-time setup: B matrix, mbarrier, TMEM ---- */
	for (int i = tid; i < U_BMAT_BYTES / 16; i += U_THREADS) {
		const uint4 v = reinterpret_cast<const uint4 *>(g_umma_bmat)[i];
		asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(bmat_a + 16 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
	}
	if (tid == 0) {
		mbar_init(bar_a, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		q_n = 0;
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(U_TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}

	/* ---- loader.  A thread's global loads share one scoreboard, so loads issued for a later stage would
	 * hold up the use of an earlier one; instead the two warp pairs alternate: pair p moves the stages of
	 * parity p, issues the loads of stage k + 2 right after it has stored stage k, and uses them a full
	 * stage later.  64 threads x 10 chunks: rows (tid/8) and (tid/8)+8, chunks (tid%8) + 8 j, j = 0..4 ---- */
	const int lpar = warp >> 1, lt = tid & 63;
	const int lrow = lt >> 3, lsub = lt & 7;
	const bool last_valid = lsub < U_ROW_CHUNKS - 8 * (U_LD / 2 - 1);
	/* chunk lsub of row lrow of stage s_begin; stage it is 32 chunks (512 B) further, row lrow + 8 is gp8 */
	const uint4 *gp = reinterpret_cast<const uint4 *>(base + (int64_t) (cg + lrow) * ch_stride + (int64_t) s_begin * U_T - U_HALO) + lsub;
	const int64_t gp8 = ch_stride;                /* 8 rows further, in uint4 units (8 * ch_stride * 2 / 16) */
	uint32_t sdst[U_LD / 2];
#pragma unroll
	for (int j = 0; j < U_LD / 2; j++)
		sdst[j] = swz64(buf_a + lrow * U_ROW_BYTES + (lsub + 8 * j) * 16);
	/* rows lrow + 8 sit 8 * 592 = 4736 B further: 4736 = 9 * 512 + 128, so their swizzle phase differs */
	uint32_t sdst8[U_LD / 2];
#pragma unroll
	for (int j = 0; j < U_LD / 2; j++)
		sdst8[j] = swz64(buf_a + (lrow + 8) * U_ROW_BYTES + (lsub + 8 * j) * 16);

	UmmaRegs pre;
	auto load_stage = [&](int it) {
		const uint4 *g = gp + (int64_t) it * (U_T / 8);
		if (dbg & 1)          /* GAIS_FIR_DBG=1 (diagnostics only): no global loads after the first stages */
			return;
#pragma unroll
		for (int j = 0; j < U_LD / 2; j++)
			if (j < U_LD / 2 - 1 || last_valid) {
				pre.v[2 * j] = __ldg(g + 8 * j);
				pre.v[2 * j + 1] = __ldg(g + gp8 + 8 * j);
			}
	};
	auto store_stage = [&](int buf) {
#pragma unroll
		for (int j = 0; j < U_LD / 2; j++)
			if (j < U_LD / 2 - 1 || last_valid) {
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sdst[j] + buf * U_BUF_STRIDE), "r"(pre.v[2 * j].x ^ 0x00800080u),
					     "r"(pre.v[2 * j].y ^ 0x00800080u), "r"(pre.v[2 * j].z ^ 0x00800080u), "r"(pre.v[2 * j].w ^ 0x00800080u)
					     : "memory");
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sdst8[j] + buf * U_BUF_STRIDE), "r"(pre.v[2 * j + 1].x ^ 0x00800080u),
					     "r"(pre.v[2 * j + 1].y ^ 0x00800080u), "r"(pre.v[2 * j + 1].z ^ 0x00800080u), "r"(pre.v[2 * j + 1].w ^ 0x00800080u)
					     : "memory");
			}
		/* the tensor core reads shared memory through the async proxy */
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	};
	auto issue_mma = [&](int buf, uint32_t tmem) {
		/* one thread: three k-steps of M128 x N96 x K32, then commit -> mbarrier */
		const uint32_t a0 = buf_a + buf * U_BUF_STRIDE + 32u;
#pragma unroll
		for (int k = 0; k < 3; k++) {
			const uint64_t ad = U_ADESC | (uint64_t) (((a0 + 32u * k) & 0x3FFFFu) >> 4);
			const uint64_t bd = U_BDESC | (uint64_t) (((bmat_a + (uint32_t) (k * U_BK_BYTES)) & 0x3FFFFu) >> 4);
			asm volatile(
				"{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
				"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
				::"r"(tmem), "l"(ad), "l"(bd), "r"(U_IDESC), "r"(k ? 1u : 0u), "r"(0u) : "memory");
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
	};

	if (lpar == 0) {
		if (s_begin == 0) {
			/* the chunks before the tile start must not be read from global memory */
#pragma unroll
			for (int j = 0; j < U_LD / 2; j++)
				if (j < U_LD / 2 - 1 || last_valid) {
					if (j == 0 && lsub < U_HALO / 8) {
						pre.v[0] = umma_hist_chunk(st + cg + lrow, hist_sel, lsub);
						pre.v[1] = umma_hist_chunk(st + cg + lrow + 8, hist_sel, lsub);
					} else {
						pre.v[2 * j] = __ldg(gp + 8 * j);
						pre.v[2 * j + 1] = __ldg(gp + gp8 + 8 * j);
					}
				}
		} else
			load_stage(0);
	} else if (n_it > 1)
		load_stage(1);
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();                              /* B matrix, barrier init and the TMEM address are visible */
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmem_base_s;
	if (lpar == 0)
		store_stage(0);
	__syncthreads();
	if (tid == 0 && !(dbg & 4)) {
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		issue_mma(0, tmem);
	}

	/* this thread's MMA row: channel c = m >> 3 of the group, word r = m & 7 of the stage */
	const int m = tid, c = m >> 3, r = m & 7;
	const uint32_t taddr = tmem + ((uint32_t) (warp * 32) << 16);
	uint32_t *sp = signs + ((int64_t) s_begin * (U_T / 32) + r) * n_channels + cg + c;
	const int64_t sp_step = (int64_t) (U_T / 32) * n_channels;
	/* linear shared address of x[n-36] for output 0 of this row: the row starts at sample t0 - 40, output j of
	 * word r is sample t0 + 32 r + j */
	const uint32_t w36 = buf_a + c * U_ROW_BYTES + (32 * r + U_HALO - GAIS_NTAPS) * 2;
	uint32_t qflags = 0;          /* bit 0: this thread has queued something since the last settling; bit 1: queue long enough */

	for (int it = 0; it < n_it; it++) {
		const int buf = it & 1;
		if ((it & 1) == lpar && it + 2 < n_it)
			load_stage(it + 2);           /* this pair stored stage `it` a stage ago: its registers are free */
		if (!(dbg & 4))       /* GAIS_FIR_DBG=4 (diagnostics only): no tensor-core work */
			mbar_wait(bar_a, (uint32_t) (it & 1));
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

		uint32_t word = 0;
		if (!(dbg & 2))       /* GAIS_FIR_DBG=2 (diagnostics only): no epilogue */
#pragma unroll
		for (int h = 0; h < 2; h++) {
			uint32_t d24[16], d16[16], d8[16], neg, clr, pend;
			tmem_ld16(taddr + 16 * h, d24);
			tmem_ld16(taddr + 32 + 16 * h, d16);
			tmem_ld16(taddr + 64 + 16 * h, d8);
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			umma_half_word(d24, d16, d8, kc, neg, clr, pend);
			if (pend) {    /* queued with a provisional 1; settled below, 32+ at a time */
				clr |= umma_push(pend, ((uint32_t) c << 28) | (uint32_t) ((s_begin + it) * U_T + 32 * r + 16 * h), smem_u32(&q_n),
						 smem_u32(q_item), w36 + buf * U_BUF_STRIDE + 32 * h);
				qflags |= clr >> 16;
			}
			word |= (~(neg | clr) & 0xffffu) << (16 * h);      /* bit j = (filtered[32 w + j] > 0) */
		}
		*sp = word;
		sp += sp_step;

		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		if (((it + 1) & 1) == lpar && it + 1 < n_it)
			store_stage(buf ^ 1);
		/* the barrier: TMEM drained by every warp, next stage complete in shared memory; it also carries the
		 * decision to settle the queue (long enough, or the CTA's last stage with something in it), so that
		 * every thread takes the same branch */
		const int settle = __syncthreads_or((int) ((qflags & 2u) | (it + 1 == n_it ? (qflags & 1u) : 0u)));
		if (it + 1 < n_it && tid == 0 && !(dbg & 4)) {
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			issue_mma(buf ^ 1, tmem);
		}
		if (settle) {
			const uint32_t nq = min(q_n, (uint32_t) U_QCAP);
			qflags = 0;
			if ((uint32_t) tid < nq) {
				const uint32_t item = q_item[tid];
				const int qc = (int) (item >> 28), n = (int) (item & 0x0fffffffu);
				if (umma_resolve_global(base + (int64_t) (cg + qc) * ch_stride, st[cg + qc].hist[hist_sel], n) == 0u)
					atomicAnd(signs + (int64_t) (n >> 5) * n_channels + cg + qc, ~(1u << (n & 31)));
			}
			__syncthreads();
			if (tid == 0)
				q_n = 0;
			__syncthreads();
		}
	}

	/* the tile ends in this CTA's last stage: its last 36 samples are the next tile's history (src/filter.c:129-134) */
	if (save_hist && s_begin + n_it == n_stages && tid < U_CH) {
		const int16_t *row = base + (int64_t) (cg + tid) * ch_stride + (int64_t) n_stages * U_T - GAIS_NTAPS;
#pragma unroll 4
		for (int i = 0; i < GAIS_NTAPS; i++)
			st[cg + tid].hist[hist_sel ^ 1][i] = row[i];
	}
	__syncthreads();
	if (warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(U_TMEM_COLS) : "memory");
}

static int g_umma_kc = 0;

static inline int fir_umma_setup(void)
{
	static UmmaTaps taps;
	umma_build_taps(&taps);
	g_umma_kc = taps.kc;
	if (cudaMemcpyToSymbol(g_umma_bmat, taps.bmat, sizeof(taps.bmat)) != cudaSuccess)
		return -1;
	if (cudaFuncSetAttribute(fir_sign_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, U_SMEM_REQUEST) != cudaSuccess)
		return -1;
	return 0;
}


} /* namespace gais */
#endif

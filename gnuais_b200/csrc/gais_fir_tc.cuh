/*
 * gais_fir_tc.cuh -- K1, the FIR-sign stage, on the tensor pipe with a TMA-fed shared-memory ring.
 *
 * Same arithmetic as gais_fir_umma.cuh (exact integer Toeplitz contraction on tcgen05.mma.kind::i8: x ^ 0x0080
 * read as two signed bytes, three int32 accumulators D24/D16/D8 per output in TMEM, Q = round(V / 65536)
 * decides the sign, the rare Q == 0 outputs are refined and, if still open, queued and settled by tiers 2/3
 * from global memory); reference src/filter.h:40-49, src/filter.c:106-143, src/receiver.c:107-111.  What
 * differs is how the samples reach the tensor core.  The first kernel staged them through registers and could
 * keep 76 KB of loads in flight per SM: it stalled at 3.5 TB/s whatever the prefetch distance
 * (profiles/r2_experiments.txt).  Here:
 *
 *   * TMA writes the MMA layout directly.  The planar sample matrix is described as a 3-D tensor
 *     {32 samples (64 B), time blocks, channels}; a box {32, 9, 16} lands as 16 channel rows of 576 contiguous
 *     bytes (t0 - 32 .. t0 + 256) stored through the hardware's 64-byte swizzle -- the same function of the
 *     absolute shared-memory address the A descriptor (SWIZZLE_64B) reads through (profiles/r2_umma_probe.txt).
 *     MMA row m = 8 c + r is the 96-byte window at byte 16 + 64 r of channel c's row: overlapping rows,
 *     nothing is replicated.  Each CTA keeps a ring of P_NS stages in flight (4 CTAs x 4 x 9 KB = 147 KB per SM).
 *   * The one thing the threads still do to the samples is flip bit 7 (so that both bytes of a sample can be
 *     read as signed: an MMA has ONE A type), in place: every 16-byte chunk of the landed stage is loaded,
 *     XORed and stored back to the same address, so the swizzle never has to be computed.  The flip of stage
 *     k + 1 runs while the tensor core works on stage k.
 *
 * A persistent one-CTA-per-SM variant with dedicated producer / MMA-issuer / epilogue warps and the raw bytes
 * fed twice (A = s8 on the high bytes, A = u8 on the low bytes, no flip at all) was built and measured first:
 * bit-exact, but one issuing thread needs ~370 cycles to issue the 6 MMAs of a 16-channel stage plus ~300
 * cycles of mbarrier round trips, twice the time HBM needs to deliver the stage (profiles/r2_experiments.txt).
 * Short-lived CTAs, each issuing its own three MMAs per stage, also interleave better with the tracker CTAs of
 * the previous time tile, which run concurrently (gais_track.cuh).
 */
#ifndef GAIS_FIR_TC_CUH
#define GAIS_FIR_TC_CUH

#include "gais_fir_umma.cuh"

namespace gais {

constexpr int P_CH = 16;                                /* channels per stage */
constexpr int P_T = 256;                                /* samples per stage */
constexpr int P_BLKS = 9;                               /* 32-sample blocks per row: t0 - 32 .. t0 + 256 */
constexpr int P_ROW_BYTES = P_BLKS * 64;                /* 576 */
constexpr int P_STAGE_BYTES = P_CH * P_ROW_BYTES;       /* 9216 = 18 x 512: a whole number of swizzle periods */
constexpr int P_STAGE_CHUNKS = P_STAGE_BYTES / 16;      /* 576 */
#ifndef P_NS
#define P_NS 4
#endif
constexpr int P_THREADS = 128;
constexpr int P_SMEM_BYTES = 50 * 1024;                 /* B matrix 9 KB + ring 36 KB (+ alignment); pins residency at 4 CTAs per SM
                                                           = 4 x 128 TMEM columns */
static_assert(U_BMAT_BYTES + P_NS * P_STAGE_BYTES + 1024 <= P_SMEM_BYTES, "ring does not fit");
#ifndef P_STAGES_PER_BLOCK
#define P_STAGES_PER_BLOCK 48
#endif
constexpr uint64_t P_ADESC = umma_desc_base(16, P_ROW_BYTES, 4);     /* SWIZZLE_64B, 8-row groups = channels, 576 B apart */

__device__ __forceinline__ void tma_g2s_3d(uint32_t dst, const CUtensorMap *tmap, int x, int y, int z, uint32_t bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
		     "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar)
		     : "memory");
}

/* queue the open outputs of one half word (provisional bit 1 stays in the sign word); a full queue settles
 * them on the spot from global memory.  Returns bits to clear (0..15), bit 16 = queued, bit 17 = queue long */
__device__ __noinline__ uint32_t tc_push(uint32_t pend, uint32_t item0, uint32_t qn_a, uint32_t q_a, const int16_t *__restrict__ row,
					 const int16_t *__restrict__ hist)
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
		} else if (umma_resolve_global(row, hist, (int) ((item0 + j) & 0x0fffffffu)) == 0u)
			clr |= 1u << j;
		if (pos + 1u >= (uint32_t) U_QDRAIN)
			clr |= 1u << 17;
	}
	return clr;
}

struct TcArgs {
	const int16_t *base;       /* tile view: sample (c, n) at base[c * ch_stride + n] */
	int64_t ch_stride;
	ChanState *st;
	uint32_t *signs;           /* [word][channel] */
	int hist_sel, n_channels, n_stages, stages_per_block, save_hist, kc;
	int dbg;                   /* GAIS_FIR_DBG, diagnostics only: 2 no epilogue arithmetic, 4 no MMAs, 8 no bit flip */
};

/* named barrier of the four epilogue warps (128 threads) that also ORs a predicate over them */
__device__ __forceinline__ int epi_sync_or(int pred)
{
	int r;
	asm volatile(
		"{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %1, 0;\n\t"
		"barrier.cta.red.or.pred q, 1, 128, p;\n\tselp.b32 %0, 1, 0, q;\n\t}"
		: "=r"(r) : "r"(pred) : "memory");
	return r;
}
__device__ __forceinline__ void epi_sync(void)
{
	asm volatile("barrier.cta.sync 1, 128;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

/*
 * Roles.  Warps 0-3 (one TMEM lane quadrant each): flip stage k + 1, wait for the accumulators of stage k,
 * tcgen05.ld -> Q -> sign word -> global, then ARRIVE (not wait) on `ready`.  Warp 4, one lane: waits for the 128
 * arrivals (= accumulators drained and next stage flipped), issues the three MMAs of the next stage and the TMA
 * request that refills the buffer the tensor core has just finished with.  Issuing an MMA blocks until the tensor
 * pipe takes it (~60 cycles each); in a thread that also runs an epilogue that time was on every stage's critical
 * path, and the CTA-wide barrier it needed cost another fifth of the kernel (profiles/r2_experiments.txt).
 */
__global__ void __launch_bounds__(P_THREADS + 32, 4)
fir_sign_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcArgs a)
{
	extern __shared__ __align__(1024) uint8_t p_smem_raw[];
	__shared__ __align__(8) uint64_t bars[P_NS + 2];
	__shared__ uint32_t tmem_base_s;
	__shared__ uint32_t q_n, q_item[U_QCAP];      /* open outputs: (channel of the group << 28) | sample index in the tile */

	const int tid = threadIdx.x, warp = tid >> 5;
	const int cg = blockIdx.x * P_CH;
	const int s_begin = blockIdx.y * a.stages_per_block;
	const int n_it = min(a.stages_per_block, a.n_stages - s_begin);
	const uint32_t smem0 = (smem_u32(p_smem_raw) + 1023u) & ~1023u;
	const uint32_t bmat_a = smem0, ring_a = smem0 + U_BMAT_BYTES;
	const uint32_t full_a = smem_u32(bars), mma_a = full_a + 8 * P_NS, ready_a = mma_a + 8;

	/* ---- one-time setup: barriers and the first TMA requests, B matrix, TMEM ---- */
	if (tid == P_THREADS) {
		for (int i = 0; i < P_NS; i++)
			mbar_init(full_a + 8 * i, 1);
		mbar_init(mma_a, 1);
		mbar_init(ready_a, P_THREADS);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		q_n = 0;
		for (int i = 0; i < P_NS && i < n_it; i++) {
			mbar_expect_tx(full_a + 8 * i, (uint32_t) P_STAGE_BYTES);
			tma_g2s_3d(ring_a + i * P_STAGE_BYTES, &tmap, 0, 8 * (s_begin + i) - 1, cg, full_a + 8 * i);
		}
	}
	for (int i = tid; i < U_BMAT_BYTES / 16; i += P_THREADS + 32) {
		const uint4 v = reinterpret_cast<const uint4 *>(g_umma_bmat)[i];
		asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(bmat_a + 16 * i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(U_TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();                              /* barriers, B matrix and the TMEM address are visible */
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmem_base_s;

	if (warp == 4) {
		/* ===== MMA issuer + TMA refills ===== */
		if (tid == P_THREADS) {
			for (int it = 0; it < n_it; it++) {
				mbar_wait(ready_a, (uint32_t) (it & 1));      /* stage `it` flipped, accumulators of stage it - 1 drained */
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t a0 = ring_a + (it % P_NS) * P_STAGE_BYTES + 16u;
				if (!(a.dbg & 4))
#pragma unroll
				for (int k = 0; k < 3; k++) {
					const uint64_t ad = P_ADESC | (uint64_t) (((a0 + 32u * k) & 0x3FFFFu) >> 4);
					const uint64_t bd = U_BDESC | (uint64_t) (((bmat_a + (uint32_t) (k * U_BK_BYTES)) & 0x3FFFFu) >> 4);
					asm volatile(
						"{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
						"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
						::"r"(tmem), "l"(ad), "l"(bd), "r"(U_IDESC), "r"(k ? 1u : 0u), "r"(0u) : "memory");
				}
				asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mma_a) : "memory");
				if (it >= 1 && it - 1 + P_NS < n_it) {
					/* the buffer of stage it - 1 is free: the epilogue threads saw its MMAs complete before they arrived */
					const int b = (it - 1) % P_NS;
					mbar_expect_tx(full_a + 8 * b, (uint32_t) P_STAGE_BYTES);
					tma_g2s_3d(ring_a + b * P_STAGE_BYTES, &tmap, 0, 8 * (s_begin + it - 1 + P_NS) - 1, cg, full_a + 8 * b);
				}
			}
		}
	} else {
		/* ===== flip + epilogue ===== */
		/* flip bit 7 of every sample of a landed stage, in place (16-byte chunks tid, tid + 128, ...: each chunk is read
		 * and written by one thread at one address, so neither the swizzle nor any ordering between threads matters) */
		auto flip_stage = [&](int it) {
			const uint32_t p0 = ring_a + (it % P_NS) * P_STAGE_BYTES + 16 * tid;
			mbar_wait(full_a + 8 * (it % P_NS), (uint32_t) ((it / P_NS) & 1));
			if (!(a.dbg & 8)) {
#pragma unroll
				for (int k = 0; k < (P_STAGE_CHUNKS + P_THREADS - 1) / P_THREADS; k++)
					if ((k + 1) * P_THREADS <= P_STAGE_CHUNKS || tid + k * P_THREADS < P_STAGE_CHUNKS) {
						uint4 v = lds128(p0 + 16 * P_THREADS * k);
						asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(p0 + 16 * P_THREADS * k), "r"(v.x ^ 0x00800080u),
							     "r"(v.y ^ 0x00800080u), "r"(v.z ^ 0x00800080u), "r"(v.w ^ 0x00800080u)
							     : "memory");
					}
			}
			/* the tensor core (and the TMA request that will refill this buffer) go through the async proxy */
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		};

		if (s_begin == 0) {
			/* first stage of the tile: TMA zero-filled the block before the tile start; the carried history goes there
			 * (samples -32..-1 = hist[4..35], src/filter.c:129-134): 16 rows x 4 chunks, written through the swizzle */
			mbar_wait(full_a, 0u);
			if (tid < 4 * P_CH) {
				const int row = tid >> 2, ch = tid & 3;
				const int16_t *h = a.st[cg + row].hist[a.hist_sel] + 4 + 8 * ch;
				uint32_t wv[4];
#pragma unroll
				for (int e = 0; e < 4; e++)
					wv[e] = (uint32_t) (uint16_t) h[2 * e] | ((uint32_t) (uint16_t) h[2 * e + 1] << 16);
				asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(swz64(ring_a + row * P_ROW_BYTES + ch * 16)), "r"(wv[0]),
					     "r"(wv[1]), "r"(wv[2]), "r"(wv[3])
					     : "memory");
			}
			epi_sync();
		}
		flip_stage(0);
		mbar_arrive(ready_a);

		/* this thread's MMA row: channel c = m >> 3 of the group, word r = m & 7 of the stage */
		const int m = tid, c = m >> 3, r = m & 7;
		const int ch = cg + c;
		const uint32_t taddr = tmem + ((uint32_t) (warp * 32) << 16);
		uint32_t *sp = a.signs + ((int64_t) s_begin * (P_T / 32) + r) * a.n_channels + ch;
		const int64_t sp_step = (int64_t) (P_T / 32) * a.n_channels;
		const int16_t *grow = a.base + (int64_t) ch * a.ch_stride;
		uint32_t qflags = 0;          /* bit 0: this thread has queued something since the last settling */

		for (int it = 0; it < n_it; it++) {
			if (it + 1 < n_it)
				flip_stage(it + 1);           /* while the tensor core works on stage `it` */
			mbar_wait(mma_a, (uint32_t) (it & 1));
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

			uint32_t word = 0;
			if (!(a.dbg & 2))
#pragma unroll
			for (int h = 0; h < 2; h++) {
				uint32_t d24[16], d16[16], d8[16], neg, clr, pend;
				tmem_ld16(taddr + 16 * h, d24);
				tmem_ld16(taddr + 32 + 16 * h, d16);
				tmem_ld16(taddr + 64 + 16 * h, d8);
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
				if (h == 1) {
					/* the accumulators are in registers and stage it + 1 is flipped: the issuer may go on */
					asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
					mbar_arrive(ready_a);
				}
				umma_half_word(d24, d16, d8, a.kc, neg, clr, pend);
				if (pend) {    /* queued with a provisional 1; settled below */
					clr |= tc_push(pend, ((uint32_t) c << 28) | (uint32_t) ((s_begin + it) * P_T + 32 * r + 16 * h), smem_u32(&q_n),
						       smem_u32(q_item), grow, a.st[ch].hist[a.hist_sel]);
					qflags |= (clr >> 16) & 1u;
				}
				word |= (~(neg | clr) & 0xffffu) << (16 * h);      /* bit j = (filtered[32 w + j] > 0) */
			}
			if (a.dbg & 2)
				mbar_arrive(ready_a);
			*sp = word;
			sp += sp_step;

			/* every 16th stage, and after the last one, the four warps meet and settle the queue if anything is in it
			 * (expected: one open output per stage; 128 entries; an overflow is settled on the spot in tc_push) */
			if ((it & 15) == 15 || it + 1 == n_it) {
				if (epi_sync_or((int) qflags)) {
					const uint32_t nq = min(q_n, (uint32_t) U_QCAP);
					qflags = 0;
					if ((uint32_t) tid < nq) {
						const uint32_t item = q_item[tid];
						const int qc = cg + (int) (item >> 28), n = (int) (item & 0x0fffffffu);
						if (umma_resolve_global(a.base + (int64_t) qc * a.ch_stride, a.st[qc].hist[a.hist_sel], n) == 0u)
							atomicAnd(a.signs + (int64_t) (n >> 5) * a.n_channels + qc, ~(1u << (n & 31)));
					}
					epi_sync();
					if (tid == 0)
						q_n = 0;
					epi_sync();
				}
			}
		}

		/* the tile ends in this CTA's last stage: its last 36 samples are the next tile's history (src/filter.c:129-134) */
		if (a.save_hist && s_begin + n_it == a.n_stages && tid < P_CH) {
			const int16_t *row = a.base + (int64_t) (cg + tid) * a.ch_stride + (int64_t) a.n_stages * P_T - GAIS_NTAPS;
#pragma unroll 4
			for (int i = 0; i < GAIS_NTAPS; i++)
				a.st[cg + tid].hist[a.hist_sel ^ 1][i] = row[i];
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(U_TMEM_COLS) : "memory");
}

static inline int fir_tc_setup(int device)
{
	(void) device;
	if (cudaFuncSetAttribute(fir_sign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES) != cudaSuccess)
		return -1;
	return 0;
}

/* 3-D tensor map over a planar tile: {32 samples, n_frames / 32 blocks, rows}, box {32, 9, 16}, 64-byte swizzle;
 * blocks outside the tile (the one before its start) read as zeros */
static inline bool fir_tc_make_tmap(CUtensorMap *tm, const int16_t *base, int64_t ch_stride, int n_rows, int64_t n_frames)
{
	const cuuint64_t gdim[3] = { 32, (cuuint64_t) (n_frames / 32), (cuuint64_t) n_rows };
	const cuuint64_t gstride[2] = { 64, (cuuint64_t) ch_stride * 2 };
	const cuuint32_t box[3] = { 32, P_BLKS, P_CH };
	const cuuint32_t estr[3] = { 1, 1, 1 };
	return g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, (void *) base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline int fir_tc_launch(SampleView view, ChanState *st, int hist_sel, int n_ch, int fast_ch, int64_t fast_frames, int save_hist,
				uint32_t *signs, int spb, int dbg, cudaStream_t stream)
{
	CUtensorMap tm;
	if (!fir_tc_make_tmap(&tm, view.base, view.ch_stride, fast_ch, fast_frames))
		return -1;
	TcArgs a;
	a.base = view.base;
	a.ch_stride = view.ch_stride;
	a.st = st;
	a.signs = signs;
	a.hist_sel = hist_sel;
	a.n_channels = n_ch;
	a.n_stages = (int) (fast_frames / P_T);
	a.stages_per_block = spb;
	a.save_hist = save_hist;
	a.kc = g_umma_kc;
	a.dbg = dbg;
	dim3 grid((unsigned) (fast_ch / P_CH), (unsigned) ((a.n_stages + spb - 1) / spb));
	fir_sign_tc_kernel<<<grid, P_THREADS + 32, P_SMEM_BYTES, stream>>>(tm, a);
	return 0;
}

/* which fast kernel GAIS_FIR_GUARD uses: 2 = tensor-core kernel with the TMA ring (this file, default), 1 = the first
 * tensor-core kernel (gais_fir_umma.cuh: register-staged loads), 0 = FFMA2 guard band
 * (gais_fir.cuh).  GAIS_FIR_IMPL=umma|ffma2 keeps the older ones selectable for A/B runs and as second
 * witnesses in the tests. */
static inline int fir_impl(void)
{
	static int v = -1;
	if (v < 0) {
		const char *e = getenv("GAIS_FIR_IMPL");
		v = (e && strcmp(e, "ffma2") == 0) ? 0 : (e && strcmp(e, "umma") == 0) ? 1 : 2;
	}
	return v;
}

/*
 * Launch K1 for one time tile.  The fast kernel takes the part of the tile it is built for
 * (planar rows, 16-byte aligned, whole channel groups, whole 256-sample stages); the exact
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
	const int impl = fir_impl();
	const bool umma = impl != 0;
	if (fir_mode == GAIS_FIR_GUARD && aligned) {
		fast_ch = umma ? n_ch / U_CH * U_CH : n_ch / F_CH * F_CH;
		fast_frames = n_frames / F_T * F_T;
	}
	if (fast_ch > 0 && fast_frames > 0) {
		const int n_stages = (int) (fast_frames / F_T);
		static int spb = 0, dbg = 0;
		if (!spb) {
			const char *e = getenv("GAIS_FIR_SPB");
			spb = (e && atoi(e) > 0) ? atoi(e) : (umma ? U_STAGES_PER_BLOCK : F_STAGES_PER_BLOCK);
			e = getenv("GAIS_FIR_DBG");
			dbg = e ? atoi(e) : 0;
		}
		if (impl == 2) {
			if (fir_tc_launch(view, st, hist_sel, n_ch, fast_ch, fast_frames, fast_frames == n_frames ? 1 : 0, signs, spb, dbg, stream) != 0)
				return -1;
		} else if (umma) {
			dim3 grid((unsigned) (fast_ch / U_CH), (unsigned) ((n_stages + spb - 1) / spb));
			fir_sign_umma_kernel<<<grid, U_THREADS, U_SMEM_REQUEST, stream>>>(view.base, view.ch_stride, st, hist_sel, n_ch, n_stages, spb,
											    signs, fast_frames == n_frames ? 1 : 0, g_umma_kc, dbg);
		} else {
			dim3 grid((unsigned) (fast_ch / F_CH), (unsigned) ((n_stages + spb - 1) / spb));
			CUtensorMap tm;
			if (!fir_make_tmap(&tm, view.base, view.ch_stride, fast_ch, fast_frames))
				return -1;
#define F_LAUNCH(D) fir_sign_fast_kernel<D><<<grid, F_THREADS, F_NSTAGE * F_STAGE_BYTES, stream>>>(tm, view.base, view.ch_stride, st, \
		hist_sel, n_ch, n_stages, spb, signs, fast_frames == n_frames ? 1 : 0)
			switch (dbg) {          /* 0 is the product; the others are the diagnostics of profiles/r1_experiments.txt */
			case 2: F_LAUNCH(2); break;
			case 4: F_LAUNCH(4); break;
			case 6: F_LAUNCH(6); break;
			case 8: F_LAUNCH(8); break;
			case 16: F_LAUNCH(16); break;
			case 20: F_LAUNCH(20); break;
			default: F_LAUNCH(0); break;
			}
#undef F_LAUNCH
		}
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

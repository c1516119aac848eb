// tools/ubench_umma_rate.cu -- how long does one tcgen05.mma.kind::i8 (M128 x N x K32) take, issued back to back by
// one thread, for the operand layouts the tensor-pipe FIR could use?  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ubench_umma_rate tools/ubench_umma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
constexpr int IMG = 96 * 1024;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
struct Args { uint64_t a_desc, b_desc; uint32_t a_start, b_start, a_kstep, b_kstep, idesc; int n_mma, ksteps, cols; int alt; uint32_t idesc2; };

__global__ void __launch_bounds__(128, 1) rate_kernel(Args pa, long long *out)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	__shared__ __align__(8) uint64_t bar;
	__shared__ uint32_t tmem_base_s;
	const int tid = threadIdx.x, warp = tid >> 5;
	for (int i = tid; i < IMG / 16; i += 128)
		reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x01010101u * (i & 3), 0x02020202u, 0x01010101u, 0);
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmem_base_s;
	if (tid == 0) {
		const uint32_t base = smem_u32(smem);
		long long t0 = clock64();
		/* descriptors of the (at most 4) k-steps up front: nothing but the MMAs in the timed loop */
		uint64_t ad[4], bd[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int kk = k % pa.ksteps;
			ad[k] = pa.a_desc | (uint64_t) (((base + pa.a_start + kk * pa.a_kstep) & 0x3FFFFu) >> 4);
			bd[k] = pa.b_desc | (uint64_t) (((base + pa.b_start + kk * pa.b_kstep) & 0x3FFFFu) >> 4);
		}
		uint32_t slot = 0;
		for (int i = 0; i < pa.n_mma; i += 4) {
#pragma unroll
			for (int k = 0; k < 4; k++) {
				/* alt 0: same MMA every time.  alt 1: odd MMAs use the second instruction descriptor (A = u8, N = 64) on
				 * columns +32 of the same slot.  alt 2: second descriptor, same columns.  alt 3: same descriptor, columns +32. */
				const bool second = pa.alt && (k & 1);
				const uint32_t id = (second && pa.alt != 3) ? pa.idesc2 : pa.idesc;
				const uint32_t dd = tmem + slot + ((second && pa.alt != 2) ? 32u : 0u);
				asm volatile(
					"{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
					"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
					::"r"(dd), "l"(ad[k]), "l"(bd[k]), "r"(id), "r"(k ? 1u : 0u), "r"(0u) : "memory");
			}
			slot = (slot + pa.cols) & 511u;
		}
		long long t1 = clock64();
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
		asm volatile(
			"{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}"
			::"r"(smem_u32(&bar)), "r"(0u) : "memory");
		long long t2 = clock64();
		if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
static uint64_t make_desc(uint32_t lbo, uint32_t sbo, uint32_t layout)
{
	return ((uint64_t) ((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t) ((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t) 1 << 46) | ((uint64_t) (layout & 7) << 61);
}
static uint32_t idesc_i8(int M, int N) { return (2u << 4) | (1u << 7) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24); }
struct Cfg { const char *name; uint32_t a_lbo, a_sbo, a_layout, a_start, a_kstep; int N, ksteps; uint32_t b_lbo, b_sbo, b_layout, b_kstep; int alt; };
int main()
{
	CK(cudaSetDevice(0));
	CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG + 1024));
	long long *d_out, h[2];
	CK(cudaMalloc(&d_out, 16));
	const Cfg cfgs[] = {
		{ "FIR layout: A sw64 overlapped sbo592, N96, 3 ksteps, B none", 16, 592, 4, 32, 32, 96, 3, 128, 256, 0, 3072 },
		{ "A sw64 overlapped sbo592, N96, B sw32-like none lbo16?", 16, 592, 4, 32, 32, 96, 3, 16, 32, 0, 3072 },
		{ "A sw64 standard sbo512 (K 64B row), N96, 2 ksteps", 16, 512, 4, 0, 32, 96, 2, 128, 256, 0, 3072 },
		{ "A sw128 standard sbo1024, N96, 4 ksteps", 16, 1024, 2, 0, 32, 96, 4, 128, 256, 0, 3072 },
		{ "A none lbo128 sbo256, N96, 1 kstep", 128, 256, 0, 0, 0, 96, 1, 128, 256, 0, 0 },
		{ "A sw128 standard, N96, B sw128 standard sbo1024", 16, 1024, 2, 0, 32, 96, 4, 16, 1024, 2, 32 },
		{ "A sw128 standard, N256, B sw128 standard", 16, 1024, 2, 0, 32, 256, 4, 16, 1024, 2, 32 },
		{ "A sw128 standard, N32, B sw128 standard", 16, 1024, 2, 0, 32, 32, 4, 16, 1024, 2, 32 },
		{ "A sw64 overlapped sbo592, N32, 3 ksteps, B none", 16, 592, 4, 32, 32, 32, 3, 128, 256, 0, 1024 },
		{ "A sw64 overlapped sbo592, N96, B sw128 standard", 16, 592, 4, 32, 32, 96, 3, 16, 1024, 2, 32 },
		{ "FIR tc: alternate (s8,N96,d) / (u8,N64,d+32)", 16, 576, 4, 16, 32, 96, 4, 128, 256, 0, 3072, 1 },
		{ "alternate (s8,N96,d) / (u8,N64,d)", 16, 576, 4, 16, 32, 96, 4, 128, 256, 0, 3072, 2 },
		{ "alternate (s8,N96,d) / (s8,N96,d+32)", 16, 576, 4, 16, 32, 96, 4, 128, 256, 0, 3072, 3 },
	};
	for (int grid : { 1 })
		for (const Cfg &c : cfgs) {
			Args a;
			a.a_desc = make_desc(c.a_lbo, c.a_sbo, c.a_layout);
			a.b_desc = make_desc(c.b_lbo, c.b_sbo, c.b_layout);
			a.a_start = c.a_start; a.b_start = 64 * 1024; a.a_kstep = c.a_kstep; a.b_kstep = c.b_kstep;
			a.idesc = idesc_i8(128, c.N); a.alt = c.alt; a.idesc2 = (2u << 4) | ((uint32_t) (64 >> 3) << 17) | ((uint32_t) (128 >> 4) << 24); a.ksteps = c.ksteps; a.cols = c.N <= 128 ? 128 : 256;
			a.n_mma = 1200;
			rate_kernel<<<grid, 128, IMG + 1024>>>(a, d_out);
			CK(cudaDeviceSynchronize());
			CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
			printf("grid %3d  %-62s : issue %6.1f cyc/mma, complete %6.1f cyc/mma\n", grid, c.name, (double) h[0] / a.n_mma, (double) h[1] / a.n_mma);
		}
	return 0;
}

// tools/ubench_umma_probe.cu -- which shared-memory byte does tcgen05.mma fetch for A[row m][byte k]?
//
// The tensor-core FIR (gais_fir_umma.cuh) feeds the int16 sample rows to tcgen05.mma.kind::i8 as an
// OVERLAPPING A operand: row m of the MMA tile is the 96-byte window that starts 64 bytes after row
// m-1's.  That is only legal if the hardware computes operand addresses the way this probe assumes:
//     linear = start + (m % 8) * pitch + (m / 8) * SBO + k          (pitch = 16/32/64/128 by layout type)
//     fetched = linear ^ swizzle(linear)                              (XOR of ABSOLUTE address bits)
// The probe measures it instead of assuming it: B is a 32 x 32 identity selector, so D[m][n] is the
// byte fetched for (m, k = n); shared memory is filled with byte p of its own offset (two passes, p = 0
// and 1), which gives the 16-bit offset of every fetched byte.  Not part of the product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ubench_umma_probe tools/ubench_umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int IMG = 48 * 1024;     // bytes of the shared-memory image (A region first, B region at B_OFF)
constexpr int B_OFF = 40 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

struct ProbeArgs {
	uint64_t a_desc_hi_lo;   // A descriptor with start address 0 (start is added in the kernel)
	uint64_t b_desc_hi_lo;
	uint32_t a_start;        // byte offset of A's start inside the image
	uint32_t idesc;
	int n_cols;              // N
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t *__restrict__ img, ProbeArgs pa, int32_t *__restrict__ out)
{
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     /* image starts 1024-aligned */
	__shared__ __align__(8) uint64_t bar;
	__shared__ uint32_t tmem_base_s;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int i = tid; i < IMG / 16; i += 128)
		reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(img)[i];
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = tmem_base_s;
	if (tid == 0) {
		const uint32_t base = smem_u32(smem);
		const uint64_t a_desc = pa.a_desc_hi_lo | (uint64_t) (((base + pa.a_start) & 0x3FFFFu) >> 4);
		const uint64_t b_desc = pa.b_desc_hi_lo | (uint64_t) (((base + B_OFF) & 0x3FFFFu) >> 4);
		asm volatile(
			"{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
			"tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
			::"r"(tmem), "l"(a_desc), "l"(b_desc), "r"(pa.idesc), "r"(0u), "r"(0u) : "memory");
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
	}
	asm volatile(
		"{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}"
		::"r"(smem_u32(&bar)), "r"(0u) : "memory");
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	for (int c0 = 0; c0 < pa.n_cols; c0 += 32) {
		uint32_t v[32];
		const uint32_t taddr = tmem + ((uint32_t) (warp * 32) << 16) + (uint32_t) c0;
		asm volatile(
			"tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
			"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
			: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
			  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
			  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
			  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
			: "r"(taddr));
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
		for (int j = 0; j < 32; j++)
			out[(warp * 32 + lane) * pa.n_cols + c0 + j] = (int32_t) v[j];
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

static uint64_t make_desc(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type, uint32_t base_offset)
{
	uint64_t d = 0;
	d |= (uint64_t) ((lbo_bytes >> 4) & 0x3FFF) << 16;
	d |= (uint64_t) ((sbo_bytes >> 4) & 0x3FFF) << 32;
	d |= (uint64_t) 1 << 46;                      // descriptor version (Blackwell)
	d |= (uint64_t) (base_offset & 7) << 49;
	d |= (uint64_t) (layout_type & 7) << 61;
	return d;
}
static uint32_t make_idesc_i8(int M, int N, int a_signed, int b_signed)
{
	return (2u << 4) | ((uint32_t) a_signed << 7) | ((uint32_t) b_signed << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

struct Cfg { const char *name; uint32_t a_start, lbo, sbo, layout, base_off; int pitch, swz_bits; };

int main()
{
	int dev = 0;
	CK(cudaSetDevice(dev));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, dev));
	printf("device: %s sm_%d%d\n", prop.name, prop.major, prop.minor);
	CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG + 1024));
	uint8_t *d_img;
	int32_t *d_out;
	CK(cudaMalloc(&d_img, IMG));
	CK(cudaMalloc(&d_out, 128 * 32 * 4));
	std::vector<uint8_t> img(IMG);
	std::vector<int32_t> out(128 * 32), addr(128 * 32);

	// layout types: 0 none (pitch 16), 6 = 32B swizzle, 4 = 64B, 2 = 128B
	const Cfg cfgs[] = {
		{ "none  lbo128 sbo256 (canonical)", 0, 128, 256, 0, 0, 16, 0 },
		{ "none  lbo16 sbo128 (overlapping rows)", 0, 16, 128, 0, 0, 16, 0 },
		{ "sw64  start+0   sbo512", 0, 16, 512, 4, 0, 64, 2 },
		{ "sw64  start+32  sbo512", 32, 16, 512, 4, 0, 64, 2 },
		{ "sw64  start+64  sbo512", 64, 16, 512, 4, 0, 64, 2 },
		{ "sw64  start+96  sbo512", 96, 16, 512, 4, 0, 64, 2 },
		{ "sw64  start+128 sbo512", 128, 16, 512, 4, 0, 64, 2 },
		{ "sw64  start+32  sbo592", 32, 16, 592, 4, 0, 64, 2 },
		{ "sw64  start+592+32+64 sbo592", 592 + 96, 16, 592, 4, 0, 64, 2 },
		{ "sw64  start+1024+48 sbo1184", 1024 + 48, 16, 1184, 4, 0, 64, 2 },
		{ "sw128 start+0   sbo1024", 0, 16, 1024, 2, 0, 128, 3 },
		{ "sw128 start+64  sbo1024", 64, 16, 1024, 2, 0, 128, 3 },
		{ "sw128 start+128+32 sbo1184", 160, 16, 1184, 2, 0, 128, 3 },
		{ "sw32  start+0   sbo256", 0, 16, 256, 6, 0, 32, 1 },
		{ "sw32  start+32+16 sbo304", 48, 16, 304, 6, 0, 32, 1 },
	};
	int all_ok = 1;
	for (const Cfg &c : cfgs) {
		for (int pass = 0; pass < 2; pass++) {
			for (int o = 0; o < IMG; o++)
				img[o] = (uint8_t) ((o >> (8 * pass)) & 0xff);
			// B: 32 x 32 identity, canonical no-swizzle K-major: byte (n, k) at (n%8)*16 + (n/8)*256 + (k/16)*128 + k%16
			memset(&img[B_OFF], 0, IMG - B_OFF);
			for (int n = 0; n < 32; n++)
				img[B_OFF + (n % 8) * 16 + (n / 8) * 256 + (n / 16) * 128 + n % 16] = 1;
			CK(cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice));
			CK(cudaMemset(d_out, 0xff, 128 * 32 * 4));
			ProbeArgs pa;
			pa.a_desc_hi_lo = make_desc(c.lbo, c.sbo, c.layout, c.base_off);
			pa.b_desc_hi_lo = make_desc(128, 256, 0, 0);
			pa.a_start = c.a_start;
			pa.idesc = make_idesc_i8(128, 32, 0, 0);
			pa.n_cols = 32;
			probe_kernel<<<1, 128, IMG + 1024>>>(d_img, pa, d_out);
			CK(cudaDeviceSynchronize());
			CK(cudaMemcpy(out.data(), d_out, 128 * 32 * 4, cudaMemcpyDeviceToHost));
			for (int i = 0; i < 128 * 32; i++)
				addr[i] = pass == 0 ? (out[i] & 0xff) : (addr[i] | ((out[i] & 0xff) << 8));
		}
		// model: linear address, then XOR of absolute address bits.  The dynamic shared window of this kernel
		// starts 1024-aligned, so offsets and absolute addresses agree in the swizzled bits.
		int bad = 0, first_bad = -1;
		for (int m = 0; m < 128; m++)
			for (int k = 0; k < 32; k++) {
				uint32_t lin;
				if (c.layout == 0)
					lin = c.a_start + (m % 8) * 16 + (m / 8) * c.sbo + (k / 16) * c.lbo + k % 16;
				else
					lin = c.a_start + (m % 8) * c.pitch + (m / 8) * c.sbo + k;
				const uint32_t sw = lin ^ (((lin >> 7) & ((1u << c.swz_bits) - 1u)) << 4);
				if ((uint32_t) addr[m * 32 + k] != sw) {
					if (first_bad < 0) first_bad = m * 32 + k;
					bad++;
				}
			}
		printf("%-40s : %s", c.name, bad ? "MODEL MISMATCH" : "model ok");
		if (bad) {
			all_ok = 0;
			printf(" (%d of 4096, first at m=%d k=%d)\n", bad, first_bad / 32, first_bad % 32);
			for (int m = 0; m < 18; m++) {
				printf("   m=%3d:", m);
				for (int k = 0; k < 32; k += 16)
					printf("  k=%2d -> %5d", k, addr[m * 32 + k]);
				printf("   (k=1 -> %5d)\n", addr[m * 32 + 1]);
			}
		} else
			printf("\n");
	}
	printf(all_ok ? "PROBE: every configuration follows linear-address + absolute-bit XOR\n" : "PROBE: see mismatches above\n");
	return 0;
}

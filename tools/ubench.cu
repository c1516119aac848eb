// tools/ubench.cu -- per-SM issue throughput of the instructions the FIR / tracking kernels
// are built from, measured on the GPU box (B200, sm_100a).  Not part of the product; the
// numbers it prints (profiles/ubench_*.txt) justify the kernel design in DESIGN.md.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define NCHAIN 8
#define UNROLL 16

#define OPS(X) \
	X(0, "ffma_rrr      ", "fma.rn.f32 %0, %0, %1, %2;", "f") \
	X(1, "ffma_imm      ", "fma.rn.f32 %0, %0, 0f3F50B242, %2;", "f") \
	X(2, "fmul_rn       ", "mul.rn.f32 %0, %0, %1;", "f") \
	X(3, "fadd_rn       ", "add.rn.f32 %0, %0, %1;", "f") \
	X(4, "fmnmx         ", "min.f32 %0, %0, %1;", "f") \
	X(5, "dp2a_lo s32u32", "dp2a.lo.s32.u32 %0, %1, %2, %0;", "r") \
	X(6, "dp4a s32u32   ", "dp4a.s32.u32 %0, %1, %2, %0;", "r") \
	X(7, "imad          ", "mad.lo.s32 %0, %0, %1, %2;", "r") \
	X(8, "iadd3         ", "add.s32 %0, %0, %1;", "r") \
	X(9, "lop3          ", "lop3.b32 %0, %0, %1, %2, 0x96;", "r") \
	X(10, "shf.l.wrap    ", "shf.l.wrap.b32 %0, %1, %0, 1;", "r") \
	X(11, "prmt          ", "prmt.b32 %0, %0, %1, 0x5410;", "r") \
	X(12, "umin          ", "min.u32 %0, %0, %1;", "r") \
	X(13, "clz           ", "clz.b32 %0, %0;", "r") \
	X(14, "popc          ", "popc.b32 %0, %0;", "r") \
	X(15, "max.s16x2     ", "max.s16x2 %0, %0, %1;", "r") \
	X(16, "brev          ", "brev.b32 %0, %0;", "r") \
	X(17, "bfe.u32       ", "bfe.u32 %0, %0, 3, 9;", "r") \
	X(18, "cvt.f32.s32   ", "cvt.rn.f32.s32 %0, %0;", "r") \
	X(20, "isetp+selp    ", "{ .reg .pred q; setp.lt.s32 q, %0, %1; selp.b32 %0, %1, %2, q; }", "r") \
	X(21, "mov           ", "mov.b32 %0, %1;", "r")

template <int OP>
__global__ void __launch_bounds__(1024) k_int(uint32_t *out, int iters, unsigned long long *cyc)
{
	uint32_t r[NCHAIN];
	uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x * 40503u + 7u;
#pragma unroll
	for (int j = 0; j < NCHAIN; j++) r[j] = a + j * 977u;
	unsigned long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
#pragma unroll
			for (int j = 0; j < NCHAIN; j++) {
#define X(id, name, txt, cls) if (OP == id && cls[0] == 'r') asm volatile(txt : "+r"(r[j]) : "r"(a), "r"(b));
				OPS(X)
#undef X
			}
		}
	}
	unsigned long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int j = 0; j < NCHAIN; j++) s ^= r[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
__global__ void __launch_bounds__(1024) k_flt(uint32_t *out, int iters, unsigned long long *cyc)
{
	float r[NCHAIN];
	float a = 1.0f + (threadIdx.x & 7) * 1e-7f, b = 1e-9f * blockIdx.x;
#pragma unroll
	for (int j = 0; j < NCHAIN; j++) r[j] = 1.0f + j;
	unsigned long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
#pragma unroll
			for (int j = 0; j < NCHAIN; j++) {
#define X(id, name, txt, cls) if (OP == id && cls[0] == 'f') asm volatile(txt : "+f"(r[j]) : "f"(a), "f"(b));
				OPS(X)
#undef X
			}
		}
	}
	unsigned long long t1 = clock64();
	float s = 0;
#pragma unroll
	for (int j = 0; j < NCHAIN; j++) s += r[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(s);
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// packed fp32x2 FMA: two MACs per instruction
__global__ void __launch_bounds__(1024) k_ffma2(uint32_t *out, int iters, unsigned long long *cyc)
{
	unsigned long long r[NCHAIN];
	unsigned long long a = 0x3F8000013F800001ull, b = 0x3089705F3089705Full;
#pragma unroll
	for (int j = 0; j < NCHAIN; j++) r[j] = 0x3F8000003F800000ull + j + threadIdx.x;
	unsigned long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
#pragma unroll
			for (int j = 0; j < NCHAIN; j++)
				asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[j]) : "l"(a), "l"(b));
		}
	}
	unsigned long long t1 = clock64();
	unsigned long long s = 0;
#pragma unroll
	for (int j = 0; j < NCHAIN; j++) s ^= r[j];
	out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t) (s ^ (s >> 32));
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// s16 -> f32 conversion (I2F) and the magic-number alternative (PRMT + FADD)
__global__ void __launch_bounds__(1024) k_i2f(uint32_t *out, int iters, unsigned long long *cyc)
{
	uint32_t r[NCHAIN];
	float acc[NCHAIN];
#pragma unroll
	for (int j = 0; j < NCHAIN; j++) { r[j] = threadIdx.x * 31u + j; acc[j] = 0; }
	unsigned long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
#pragma unroll
			for (int j = 0; j < NCHAIN; j++) {
				float f;
				asm volatile("{ .reg .s16 lo, hi; mov.b32 {lo, hi}, %1; cvt.rn.f32.s16 %0, hi; }" : "=f"(f) : "r"(r[j]));
				r[j] = __float_as_uint(f) ^ (r[j] >> 3);
			}
		}
	}
	unsigned long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int j = 0; j < NCHAIN; j++) s ^= r[j] ^ __float_as_uint(acc[j]);
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mixes: n_fma FFMA(imm) + n_alu SHF per group, to see whether the two pipes co-issue
template <int NF, int NA, bool PACKED>
__global__ void __launch_bounds__(1024) k_mix(uint32_t *out, int iters, unsigned long long *cyc)
{
	float f[8];
	unsigned long long p[8];
	uint32_t r[8];
	uint32_t a = threadIdx.x * 2654435761u;
#pragma unroll
	for (int j = 0; j < 8; j++) { f[j] = 1.0f + j; r[j] = a + j; p[j] = 0x3F8000003F800000ull + j; }
	unsigned long long t0 = clock64();
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
#pragma unroll
			for (int j = 0; j < NF; j++) {
				if (PACKED)
					asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[j & 7]) : "l"(0x3F8000013F800001ull));
				else
					asm volatile("fma.rn.f32 %0, %0, 0f3F50B242, %1;" : "+f"(f[j & 7]) : "f"(1e-9f));
			}
#pragma unroll
			for (int j = 0; j < NA; j++)
				asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(r[j & 7]) : "r"(a));
		}
	}
	unsigned long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int j = 0; j < 8; j++) s ^= r[j] ^ __float_as_uint(f[j]) ^ (uint32_t) p[j] ^ (uint32_t) (p[j] >> 32);
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static uint32_t *d_out;
static unsigned long long *d_cyc, h_cyc[1024];
static int n_sm;

template <typename F>
static void run(const char *name, F launch, double ops_per_thread_iter, int iters)
{
	int blocks = n_sm * 2;
	launch(blocks, 8);            // warm
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	launch(blocks, iters);
	cudaEventRecord(e1);
	cudaDeviceSynchronize();
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	cudaMemcpy(h_cyc, d_cyc, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost);
	double cyc = 0;
	for (int i = 0; i < blocks; i++) cyc += (double) h_cyc[i];
	cyc /= blocks;
	double ops_sm = 2048.0 * ops_per_thread_iter * iters;   // 2 blocks x 1024 threads per SM
	printf("%-28s %8.1f lane-ops/clk/SM   (%.0f cycles, %.3f ms, %.2f Tops/s chip, eff clk %.0f MHz)\n", name, ops_sm / cyc, cyc, ms,
	       ops_sm * n_sm / (ms * 1e-3) / 1e12, cyc / (ms * 1e-3) / 1e6);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	n_sm = prop.multiProcessorCount;
	printf("device: %s, %d SMs, sm_%d%d, clock %d kHz\n", prop.name, n_sm, prop.major, prop.minor, prop.clockRate);
	cudaMalloc(&d_out, 4ull * 1024 * 2 * n_sm);
	cudaMalloc(&d_cyc, 8ull * 2 * n_sm);
	const int iters = 2000;
	const double per = NCHAIN * UNROLL;
#define X(id, name, txt, cls) \
	if (cls[0] == 'r') run(name, [&](int b, int it) { k_int<id><<<b, 1024>>>(d_out, it, d_cyc); }, per, iters); \
	else run(name, [&](int b, int it) { k_flt<id><<<b, 1024>>>(d_out, it, d_cyc); }, per, iters);
	OPS(X)
#undef X
	run("ffma2 (instr; x2 MACs)", [&](int b, int it) { k_ffma2<<<b, 1024>>>(d_out, it, d_cyc); }, per, iters);
	run("i2f.s16 + lop (pairs)", [&](int b, int it) { k_i2f<<<b, 1024>>>(d_out, it, d_cyc); }, per, iters);
	run("mix 8 ffma_imm + 0 shf", [&](int b, int it) { k_mix<8, 0, false><<<b, 1024>>>(d_out, it, d_cyc); }, 8 * 8, iters);
	run("mix 8 ffma_imm + 4 shf", [&](int b, int it) { k_mix<8, 4, false><<<b, 1024>>>(d_out, it, d_cyc); }, 8 * 12, iters);
	run("mix 8 ffma_imm + 8 shf", [&](int b, int it) { k_mix<8, 8, false><<<b, 1024>>>(d_out, it, d_cyc); }, 8 * 16, iters);
	run("mix 8 ffma2 + 0 shf", [&](int b, int it) { k_mix<8, 0, true><<<b, 1024>>>(d_out, it, d_cyc); }, 8 * 8, iters);
	run("mix 8 ffma2 + 4 shf", [&](int b, int it) { k_mix<8, 4, true><<<b, 1024>>>(d_out, it, d_cyc); }, 8 * 12, iters);
	run("mix 8 ffma2 + 8 shf", [&](int b, int it) { k_mix<8, 8, true><<<b, 1024>>>(d_out, it, d_cyc); }, 8 * 16, iters);
	run("mix 4 ffma2 + 8 shf", [&](int b, int it) { k_mix<4, 8, true><<<b, 1024>>>(d_out, it, d_cyc); }, 8 * 12, iters);
	return 0;
}

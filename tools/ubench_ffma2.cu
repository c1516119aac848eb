// tools/ubench_ffma2.cu -- does FFMA2 share its issue cycles with ALU-pipe instructions?  80 packed FMAs of the
// FIR inner block (10 taps x 8 outputs) in two orders, with 0/40/80 SHF interleaved.  Not part of the product.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ void ffma2(unsigned long long &acc, unsigned long long x, unsigned long long t)
{
	asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(x), "l"(t));
}
__device__ __forceinline__ void shf(uint32_t &r, uint32_t a)
{
	asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(r) : "r"(a));
}
__device__ __forceinline__ void iadd(uint32_t &r, uint32_t a)
{
	asm volatile("add.u32 %0, %1, %0;" : "+r"(r) : "r"(a));
}

// ORDER 0: input stationary (same x for 10 consecutive FMAs, 10 accumulators rotate)
// ORDER 1: output stationary, 8 chains interleaved (same tap for 8 consecutive FMAs)
// ORDER 2: output stationary, one chain after the other (10 dependent FMAs in a row)
// BCAST: taps as {t,t} pairs the assembler can turn into a scalar .F32 operand
template <int ORDER, int NALU, bool BCAST, int ALUOP>
__global__ void __launch_bounds__(256, 3) k(uint32_t *out, int iters, const float *tp)
{
	unsigned long long x[18], acc[10], t[5];
	uint32_t r[8];
	const uint32_t a = threadIdx.x * 2654435761u;
#pragma unroll
	for (int i = 0; i < 18; i++) x[i] = 0x3F8000003F800000ull + ((unsigned long long) (threadIdx.x + i) << 32) + i;
#pragma unroll
	for (int i = 0; i < 10; i++) acc[i] = 0;
#pragma unroll
	for (int i = 0; i < 5; i++) {
		const uint32_t lo = __float_as_uint(tp[i]), hi = BCAST ? lo : __float_as_uint(tp[i + 5]);
		asm("mov.b64 %0, {%1, %2};" : "=l"(t[i]) : "r"(lo), "r"(hi));
	}
#pragma unroll
	for (int i = 0; i < 8; i++) r[i] = a + i;
	for (int it = 0; it < iters; it++) {
		int n = 0;
#define ALU() do { if (NALU && (n * NALU / 80) != ((n + 1) * NALU / 80)) { if (ALUOP == 0) shf(r[n & 7], a); else iadd(r[n & 7], a); } n++; } while (0)
		if (ORDER == 0) {
#pragma unroll
			for (int i = 0; i < 8; i++)
#pragma unroll
				for (int kk = 0; kk < 10; kk++) { ffma2(acc[(i + kk) % 10], x[i + 10], t[kk < 5 ? kk : 9 - kk]); ALU(); }
		} else if (ORDER == 1) {
#pragma unroll
			for (int kk = 0; kk < 10; kk++)
#pragma unroll
				for (int j = 0; j < 8; j++) { ffma2(acc[j], x[j + kk], t[kk < 5 ? kk : 9 - kk]); ALU(); }
		} else {
#pragma unroll
			for (int j = 0; j < 8; j++)
#pragma unroll
				for (int kk = 0; kk < 10; kk++) { ffma2(acc[j], x[j + kk], t[kk < 5 ? kk : 9 - kk]); ALU(); }
		}
	}
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < 10; i++) s ^= (uint32_t) acc[i] ^ (uint32_t) (acc[i] >> 32);
#pragma unroll
	for (int i = 0; i < 8; i++) s ^= r[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static uint32_t *d_out;
static float *d_t;
static int n_sm, clk_khz;

template <int ORDER, int NALU, bool BCAST, int ALUOP>
static void run(const char *name)
{
	const int iters = 3000, blocks = n_sm * 3;
	k<ORDER, NALU, BCAST, ALUOP><<<blocks, 256>>>(d_out, 16, d_t);
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 3; rep++) {
		cudaEventRecord(e0);
		k<ORDER, NALU, BCAST, ALUOP><<<blocks, 256>>>(d_out, iters, d_t);
		cudaEventRecord(e1);
		cudaDeviceSynchronize();
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		best = ms < best ? ms : best;
	}
	const double cyc = best * 1e-3 * clk_khz * 1e3 / (6.0 * iters);   /* per block of 80 FFMA2 (+NALU), per SMSP with 6 warps */
	printf("%-64s %6.1f cycles per 80 FFMA2 + %2d ALU  (%.2f per FFMA2)  %s\n", name, cyc, NALU, cyc / 80.0, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	n_sm = prop.multiProcessorCount;
	cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	printf("device: %s, %d SMs, %d kHz; 3 CTAs x 8 warps per SM\n", prop.name, n_sm, clk_khz);
	cudaMalloc(&d_out, 4ull * 256 * 3 * n_sm);
	float ht[10] = { 2.2e-4f, 5.9e-3f, 6.9e-2f, 0.358f, 0.815f, 2.3e-4f, 5.8e-3f, 6.8e-2f, 0.359f, 0.814f };
	cudaMalloc(&d_t, sizeof(ht));
	cudaMemcpy(d_t, ht, sizeof(ht), cudaMemcpyHostToDevice);
	run<0, 0, true, 0>("input-stationary, scalar taps");
	run<0, 40, true, 0>("input-stationary, scalar taps, +40 SHF");
	run<0, 80, true, 0>("input-stationary, scalar taps, +80 SHF");
	run<0, 80, true, 1>("input-stationary, scalar taps, +80 IADD");
	run<0, 0, false, 0>("input-stationary, 64-bit taps");
	run<0, 80, false, 0>("input-stationary, 64-bit taps, +80 SHF");
	run<1, 0, true, 0>("output-stationary x8 interleaved, scalar taps");
	run<1, 40, true, 0>("output-stationary x8 interleaved, scalar taps, +40 SHF");
	run<1, 80, true, 0>("output-stationary x8 interleaved, scalar taps, +80 SHF");
	run<1, 80, false, 0>("output-stationary x8 interleaved, 64-bit taps, +80 SHF");
	run<2, 0, true, 0>("output-stationary chains, scalar taps");
	run<2, 80, true, 0>("output-stationary chains, scalar taps, +80 SHF");
	return 0;
}

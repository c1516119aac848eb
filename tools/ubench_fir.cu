// tools/ubench_fir.cu -- where do the cycles of one FIR word column go?  Runs the column code of
// gais_fir.cuh from shared memory (no HBM traffic, 3 CTAs x 8 warps per SM like the product) with parts
// switched off, and prints cycles per column per SM sub-partition.  Not part of the product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Ignuais_b200/csrc -o build/ubench_fir tools/ubench_fir.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include "gais_fir.cuh"

using namespace gais;

enum { NO_CVT = 1, NO_SIGN = 2, NO_MIN = 4, FMA_ONLY = 8, SYM = 16, ORDERED = 32 };

template <int V>
__global__ void __launch_bounds__(256, 3) col_kernel(const int16_t *__restrict__ src, uint32_t *__restrict__ out, int iters)
{
	extern __shared__ __align__(128) uint8_t tile[];
	for (int i = threadIdx.x; i < F_NSTAGE * F_STAGE_BYTES / 2; i += blockDim.x)
		reinterpret_cast<int16_t *>(tile)[i] = src[i];
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint64_t T[6];
#pragma unroll
	for (int k = 0; k < 6; k++)
		T[k] = pack2(c_taps[12 + k], c_taps[12 + k]);
	const uint64_t sub = pack2(F_MAGIC_SUB, F_MAGIC_SUB);
	const uint32_t row0 = smem_u32(tile) + lane * F_ROW_BYTES + (32 * warp + 16) * 2;
	uint32_t chk = 0;
	for (int it = 0; it < iters; it++) {
		const uint32_t rowA = row0 + (it & 1) * F_STAGE_BYTES, rowB = rowA + 32 * F_ROW_BYTES;
		uint64_t xs[48];
		auto load = [&](int g) {
			if (V & NO_CVT) {
				const uint4 a = lds128(rowA + g * 16), b = lds128(rowB + g * 16);
				xs[8 * g + 0] = ((uint64_t) b.x << 32) | a.x; xs[8 * g + 1] = ((uint64_t) b.y << 32) | a.y;
				xs[8 * g + 2] = ((uint64_t) b.z << 32) | a.z; xs[8 * g + 3] = ((uint64_t) b.w << 32) | a.w;
				xs[8 * g + 4] = ((uint64_t) a.x << 32) | b.x; xs[8 * g + 5] = ((uint64_t) a.y << 32) | b.y;
				xs[8 * g + 6] = ((uint64_t) a.z << 32) | b.z; xs[8 * g + 7] = ((uint64_t) a.w << 32) | b.w;
			} else
				fir_load_chunk(xs + 8 * g, rowA, rowB, g * 16, sub);
		};
		load(0);
		load(1);
		uint32_t wordA = 0, wordB = 0, pend = 0, keep = 0;
#pragma unroll
		for (int g = 0; g < 4; g++) {
			load(g + 2);
			uint64_t acc[8];
			float m = 3.0e38f;
#pragma unroll
			for (int jj = 0; jj < 8; jj++) {
				const int j = 8 * g + jj;
				uint64_t a;
				if (V & SYM) {
					/* symmetric taps: add the mirrored samples first (5 FADD2 + 5 FFMA2 instead of 10 FFMA2) */
					a = fmul2(T[1], fadd2(xs[j + 1], xs[j + 10]));
					a = ffma2(T[2], fadd2(xs[j + 2], xs[j + 9]), a);
					a = ffma2(T[3], fadd2(xs[j + 3], xs[j + 8]), a);
					a = ffma2(T[4], fadd2(xs[j + 4], xs[j + 7]), a);
					a = ffma2(T[5], fadd2(xs[j + 5], xs[j + 6]), a);
				} else {
					a = fmul2(T[1], xs[j + 1]);
					a = ffma2(T[2], xs[j + 2], a);
					a = ffma2(T[3], xs[j + 3], a);
					a = ffma2(T[4], xs[j + 4], a);
					a = ffma2(T[5], xs[j + 5], a);
					a = ffma2(T[5], xs[j + 6], a);
					a = ffma2(T[4], xs[j + 7], a);
					a = ffma2(T[3], xs[j + 8], a);
					a = ffma2(T[2], xs[j + 9], a);
					a = ffma2(T[1], xs[j + 10], a);
				}
				acc[jj] = a;
				float ya, yb;
				unpack2(a, ya, yb);
				if (V & FMA_ONLY) {
					keep ^= __float_as_uint(ya) & __float_as_uint(yb);     /* one LOP3 per output keeps it alive */
					continue;
				}
				if (!(V & NO_MIN))
					m = fminf(m, fminf(fabsf(ya), fabsf(yb)));
				if (!(V & NO_SIGN)) {
					wordA = __funnelshift_l(__float_as_uint(ya), wordA, 1);
					wordB = __funnelshift_l(__float_as_uint(yb), wordB, 1);
				} else if (V & NO_MIN)
					keep ^= __float_as_uint(ya) & __float_as_uint(yb);
			}
			if (!(V & (NO_MIN | FMA_ONLY)) && m <= F_E1) {
#pragma unroll
				for (int jj = 0; jj < 8; jj++) {
					float ya, yb;
					unpack2(acc[jj], ya, yb);
					if (fabsf(ya) <= F_E1) pend |= 1u << (8 * g + jj);
					if (fabsf(yb) <= F_E1) pend |= 1u << (8 * g + jj);
				}
			}
		}
		chk += __brev(~wordA) ^ __brev(~wordB) ^ pend ^ keep;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = chk;
}

/* the same column with scalar FFMA and immediate taps: lane l still owns channels l and l+32, as two scalar streams */
#define TAPF(k) __uint_as_float(k == 1 ? 0x396a68bfu : k == 2 ? 0x3bc2cc99u : k == 3 ? 0x3d8e92d5u : k == 4 ? 0x3eb7cd8au : 0x3f50b242u)
template <int V>
__global__ void __launch_bounds__(256, 3) col_kernel_scalar(const int16_t *__restrict__ src, uint32_t *__restrict__ out, int iters)
{
	extern __shared__ __align__(128) uint8_t tile[];
	for (int i = threadIdx.x; i < F_NSTAGE * F_STAGE_BYTES / 2; i += blockDim.x)
		reinterpret_cast<int16_t *>(tile)[i] = src[i];
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t row0 = smem_u32(tile) + lane * F_ROW_BYTES + (32 * warp + 16) * 2;
	uint32_t chk = 0;
	for (int it = 0; it < iters; it++) {
		const uint32_t rowA = row0 + (it & 1) * F_STAGE_BYTES, rowB = rowA + 32 * F_ROW_BYTES;
		float xa[48], xb[48];
		auto cv = [&](uint32_t w, uint32_t sel) { return __fadd_rn(__uint_as_float(__byte_perm(w, F_MAGIC_BITS, sel)), F_MAGIC_SUB); };
		auto load = [&](int g) {
			uint4 a = lds128(rowA + g * 16), b = lds128(rowB + g * 16);
			a.x ^= 0x80008000u; a.y ^= 0x80008000u; a.z ^= 0x80008000u; a.w ^= 0x80008000u;
			b.x ^= 0x80008000u; b.y ^= 0x80008000u; b.z ^= 0x80008000u; b.w ^= 0x80008000u;
			xa[8 * g + 0] = cv(a.x, 0x7610); xa[8 * g + 1] = cv(a.x, 0x7632); xa[8 * g + 2] = cv(a.y, 0x7610); xa[8 * g + 3] = cv(a.y, 0x7632);
			xa[8 * g + 4] = cv(a.z, 0x7610); xa[8 * g + 5] = cv(a.z, 0x7632); xa[8 * g + 6] = cv(a.w, 0x7610); xa[8 * g + 7] = cv(a.w, 0x7632);
			xb[8 * g + 0] = cv(b.x, 0x7610); xb[8 * g + 1] = cv(b.x, 0x7632); xb[8 * g + 2] = cv(b.y, 0x7610); xb[8 * g + 3] = cv(b.y, 0x7632);
			xb[8 * g + 4] = cv(b.z, 0x7610); xb[8 * g + 5] = cv(b.z, 0x7632); xb[8 * g + 6] = cv(b.w, 0x7610); xb[8 * g + 7] = cv(b.w, 0x7632);
		};
		load(0);
		load(1);
		uint32_t wordA = 0, wordB = 0, pend = 0, keep = 0;
#pragma unroll
		for (int g = 0; g < 4; g++) {
			load(g + 2);
			float ya[8], yb[8];
			float m = 3.0e38f;
#pragma unroll
			for (int jj = 0; jj < 8; jj++) {
				const int j = 8 * g + jj;
				auto fir = [&](const float *x) {
					float a = x[j + 1] * TAPF(1);
					a = fmaf(x[j + 2], TAPF(2), a); a = fmaf(x[j + 3], TAPF(3), a); a = fmaf(x[j + 4], TAPF(4), a);
					a = fmaf(x[j + 5], TAPF(5), a); a = fmaf(x[j + 6], TAPF(5), a); a = fmaf(x[j + 7], TAPF(4), a);
					a = fmaf(x[j + 8], TAPF(3), a); a = fmaf(x[j + 9], TAPF(2), a); a = fmaf(x[j + 10], TAPF(1), a);
					return a;
				};
				ya[jj] = fir(xa);
				yb[jj] = fir(xb);
				if (V & FMA_ONLY) {
					keep ^= __float_as_uint(ya[jj]) & __float_as_uint(yb[jj]);
					continue;
				}
				m = fminf(m, fminf(fabsf(ya[jj]), fabsf(yb[jj])));
				wordA = __funnelshift_l(__float_as_uint(ya[jj]), wordA, 1);
				wordB = __funnelshift_l(__float_as_uint(yb[jj]), wordB, 1);
			}
			if (!(V & FMA_ONLY) && m <= F_E1) {
#pragma unroll
				for (int jj = 0; jj < 8; jj++) {
					if (fabsf(ya[jj]) <= F_E1) pend |= 1u << (8 * g + jj);
					if (fabsf(yb[jj]) <= F_E1) pend |= 1u << (8 * g + jj);
				}
			}
		}
		chk += __brev(~wordA) ^ __brev(~wordB) ^ pend ^ keep;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = chk;
}

template <int V, bool SCALAR = false>
static void run(const char *name, const int16_t *d_src, uint32_t *d_out, int n_sm, int iters)
{
	const int smem = F_NSTAGE * F_STAGE_BYTES;
	auto kern = SCALAR ? col_kernel_scalar<V> : col_kernel<V>;
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	int occ = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
	cudaFuncAttributes fa;
	cudaFuncGetAttributes(&fa, kern);
	const int blocks = n_sm * 3;
	kern<<<blocks, 256, smem>>>(d_src, d_out, 64);
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e30f;
	for (int rep = 0; rep < 3; rep++) {
		cudaEventRecord(e0);
		kern<<<blocks, 256, smem>>>(d_src, d_out, iters);
		cudaEventRecord(e1);
		cudaDeviceSynchronize();
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		best = ms < best ? ms : best;
	}
	int clk_khz = 0;
	cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
	/* columns per SM sub-partition: 3 CTAs x 8 warps / 4 = 6 warps, each `iters` columns */
	const double cyc_per_col = best * 1e-3 * clk_khz * 1e3 / (6.0 * iters);
	const double tile_ms = 65536.0 * 32768.0 / 2048.0 / (n_sm * 4.0) * cyc_per_col / (clk_khz * 1e3) * 1e3;
	printf("%-44s %7.1f cycles/column/SMSP (at %d kHz)  regs %3d occ %d  -> %.3f ms per 65536x32768 tile  %s\n", name, cyc_per_col, clk_khz,
	       fa.numRegs, occ, tile_ms, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	const int n_sm = prop.multiProcessorCount;
	printf("device: %s, %d SMs\n", prop.name, n_sm);
	static const uint32_t half[18] = GAIS_TAP_BITS_HALF;
	uint32_t full[GAIS_NTAPS];
	for (int i = 0; i < GAIS_NTAPS; i++)
		full[i] = half[i < 18 ? i : 35 - i];
	cudaMemcpyToSymbol(c_taps, full, sizeof(full));
	const int n = F_NSTAGE * F_STAGE_BYTES / 2;
	int16_t *h = (int16_t *) malloc(n * 2), *d_src;
	uint32_t s = 12345u;
	for (int i = 0; i < n; i++) {
		s = s * 1664525u + 1013904223u;
		h[i] = (int16_t) ((int) (s >> 16) % 12000 - 6000 + ((i / 7) & 1 ? 9000 : -9000));
	}
	cudaMalloc(&d_src, n * 2);
	cudaMemcpy(d_src, h, n * 2, cudaMemcpyHostToDevice);
	uint32_t *d_out;
	cudaMalloc(&d_out, 4ull * 256 * 3 * n_sm);
	const int iters = 4000;
	run<0>("full column (cvt + 320 FFMA2 + sign + min)", d_src, d_out, n_sm, iters);
	run<NO_CVT>("  without int16->f32 conversion", d_src, d_out, n_sm, iters);
	run<NO_SIGN>("  without sign SHF", d_src, d_out, n_sm, iters);
	run<NO_MIN>("  without min/guard test", d_src, d_out, n_sm, iters);
	run<NO_SIGN | NO_MIN>("  without sign and min (1 LOP3/output)", d_src, d_out, n_sm, iters);
	run<FMA_ONLY>("  cvt + FFMA2 only (1 LOP3/output)", d_src, d_out, n_sm, iters);
	run<FMA_ONLY | NO_CVT>("  LDS + FFMA2 only (1 LOP3/output)", d_src, d_out, n_sm, iters);
	run<SYM>("symmetric pre-add (5 FADD2 + 5 FFMA2)", d_src, d_out, n_sm, iters);
	run<SYM | FMA_ONLY | NO_CVT>("  symmetric, LDS + FMA only", d_src, d_out, n_sm, iters);
	run<0, true>("scalar FFMA, immediate taps: full column", d_src, d_out, n_sm, iters);
	run<FMA_ONLY, true>("  scalar: cvt + FFMA only (1 LOP3/output)", d_src, d_out, n_sm, iters);
	return 0;
}

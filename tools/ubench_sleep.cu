// How long do nanosleep and mbarrier.try_wait (with and without a suspend-time hint) really hold a warp?  B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(long long *out, int other_warps_busy)
{
	__shared__ __align__(8) uint64_t bar;
	const uint32_t b = (uint32_t) __cvta_generic_to_shared(&bar);
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
	}
	__syncthreads();
	if (threadIdx.x >= 32) {
		if (other_warps_busy) {	/* keep the scheduler busy */
			float x = threadIdx.x;
			for (int i = 0; i < 200000; i++) x = x * 1.0001f + 0.5f;
			if (x == 12345.f) out[63] = 1;
		}
		return;
	}
	long long t0, t1;
	int idx = 0;
#define T(stmt) do { t0 = clock64(); for (int i = 0; i < 64; i++) { stmt; } t1 = clock64(); if (threadIdx.x == 0) out[idx] = (t1 - t0) / 64; idx++; } while (0)
	T(asm volatile("nanosleep.u32 100;" ::: "memory"));
	T(asm volatile("nanosleep.u32 400;" ::: "memory"));
	T(asm volatile("nanosleep.u32 2000;" ::: "memory"));
	T(asm volatile("nanosleep.u32 10000;" ::: "memory"));
	{ uint32_t ns = 400; T(asm volatile("nanosleep.u32 %0;" ::"r"(ns) : "memory")); }
	T(asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;}" ::"r"(b) : "memory"));
	T(asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, %1;}" ::"r"(b), "r"(4000u) : "memory"));
	T(asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, %1;}" ::"r"(b), "r"(1000000u) : "memory"));
	T(asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%0], 0;}" ::"r"(b) : "memory"));
}
int main()
{
	long long *d, h[64];
	cudaMalloc(&d, sizeof(h));
	const char *names[] = { "nanosleep 100", "nanosleep 400", "nanosleep 2000", "nanosleep 10000", "nanosleep reg 400", "try_wait (pending)",
				"try_wait hint 4000", "try_wait hint 1e6", "test_wait" };
	for (int busy = 0; busy < 2; busy++) {
		cudaMemset(d, 0, sizeof(h));
		k<<<1, busy ? 256 : 32>>>(d, busy);
		cudaError_t e = cudaDeviceSynchronize();
		cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
		printf("%s (%s)\n", busy ? "7 busy warps beside" : "alone", cudaGetErrorString(e));
		for (int i = 0; i < 9; i++)
			printf("  %-22s %8lld cycles per call\n", names[i], h[i]);
	}
	return 0;
}

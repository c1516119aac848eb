/*
 * gais_synth.cu -- host and device front-ends of the integer-only workload generator
 * (synth_core.h).  gais_synth_host() needs no GPU (it is what the CPU tests and the oracle
 * comparisons use); gais_synth_device() fills HBM directly for the large bench configs.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "gais_b200.h"
#include "synth_core.h"

#include "synth_host_impl.h"

extern "C" int gais_synth_host(const gais_synth *p, uint32_t first_channel, int32_t n_channels, int64_t n_frames,
			       int16_t *h_out, int32_t layout, int64_t stride)
{
	return gs_fill_host(p, first_channel, n_channels, n_frames, h_out, layout, stride);
}

/* one warp per (channel, slot pair): lane 0 builds the bursts in shared memory, all lanes
 * synthesise samples (coalesced 64-byte stores per warp in the planar layout) */
__global__ void __launch_bounds__(128)
synth_kernel(gais_synth_params q, uint32_t first_channel, int32_t n_channels, int64_t n_frames, int64_t n_pairs,
	     int16_t *out, int64_t ch_stride, int64_t t_stride)
{
	__shared__ gs_burst sb[4][2];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int64_t unit = (int64_t) blockIdx.x * 4 + warp;
	if (unit >= n_pairs * n_channels)
		return;
	const int32_t c = (int32_t) (unit / n_pairs);
	const uint32_t pair = (uint32_t) (unit % n_pairs);
	const uint32_t ck = gs_channel_key(q.seed, first_channel + (uint32_t) c);
	if (lane == 0)
		gs_build_pair(sb[warp], ck, pair, &q);
	__syncwarp();
	const int64_t n0 = (int64_t) pair * GS_PAIR_SAMPLES;
	int16_t *row = out + (int64_t) c * ch_stride;
	for (int m = lane; m < GS_PAIR_SAMPLES; m += 32) {
		int64_t n = n0 + m;
		if (n < n_frames)
			row[n * t_stride] = gs_sample(sb[warp], ck, (uint32_t) n, m, &q);
	}
}

extern "C" int gais_synth_device(const gais_synth *p, uint32_t first_channel, int32_t n_channels, int64_t n_frames,
				 int16_t *d_out, int32_t layout, int64_t stride, void *stream)
{
	if (!p || !d_out || n_channels < 1 || n_frames < 1)
		return GAIS_EINVAL;
	gais_synth_params q = gs_to_params(p);
	const int64_t ch_stride = (layout == GAIS_LAYOUT_PLANAR) ? stride : 1;
	const int64_t t_stride = (layout == GAIS_LAYOUT_PLANAR) ? 1 : stride;
	const int64_t n_pairs = (n_frames + GS_PAIR_SAMPLES - 1) / GS_PAIR_SAMPLES;
	const int64_t units = n_pairs * n_channels;
	const int64_t blocks = (units + 3) / 4;
	if (blocks > 0x7fffffffLL)
		return GAIS_EINVAL;
	synth_kernel<<<(unsigned) blocks, 128, 0, (cudaStream_t) stream>>>(q, first_channel, n_channels, n_frames, n_pairs, d_out,
									  ch_stride, t_stride);
	return cudaGetLastError() == cudaSuccess ? 0 : GAIS_ECUDA;
}

/*
 * gais_track.cuh -- launch wrappers of the tracking stage (K2+K3).
 */
#ifndef GAIS_TRACK_CUH
#define GAIS_TRACK_CUH

#include "gais_kernels.cuh"

namespace gais {

static inline int track_launch(const uint32_t *signs, ChanState *st, int n_ch, int64_t n_frames, const TrackOut &out,
			       cudaStream_t stream)
{
	track_simple_kernel<<<(n_ch + 127) / 128, 128, 0, stream>>>(signs, st, n_ch, n_frames, out);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

} /* namespace gais */
#endif

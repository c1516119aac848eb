/*
 * synth_host_impl.h -- the host loop of the workload generator (synth_core.h), shared by libgaisb200.so
 * (gais_synth_host of the C-ABI) and libgais_synth.so (the same entry point in a library of its own: plain C, no
 * CUDA, no product code -- what bench.py --impl reference and the CPU tests load, so that the reference arm maps
 * nothing of the product).
 */
#ifndef GAIS_SYNTH_HOST_IMPL_H
#define GAIS_SYNTH_HOST_IMPL_H

#include "gais_b200.h"
#include "synth_core.h"

static inline gais_synth_params gs_to_params(const gais_synth *p)
{
	gais_synth_params q;
	q.seed = p->seed;
	q.amplitude = p->amplitude;
	q.noise_q16 = p->noise_q16;
	q.rho_q16 = p->rho_q16;
	q.jitter = p->jitter;
	return q;
}

static inline int gs_fill_host(const gais_synth *p, uint32_t first_channel, int32_t n_channels, int64_t n_frames, int16_t *h_out,
			       int32_t layout, int64_t stride)
{
	if (!p || !h_out || n_channels < 1 || n_frames < 1)
		return GAIS_EINVAL;
	gais_synth_params q = gs_to_params(p);
	const int64_t ch_stride = (layout == GAIS_LAYOUT_PLANAR) ? stride : 1;
	const int64_t t_stride = (layout == GAIS_LAYOUT_PLANAR) ? 1 : stride;
	for (int32_t c = 0; c < n_channels; c++) {
		uint32_t ck = gs_channel_key(q.seed, first_channel + (uint32_t) c);
		int16_t *row = h_out + (int64_t) c * ch_stride;
		for (int64_t n0 = 0; n0 < n_frames; n0 += GS_PAIR_SAMPLES) {
			gs_burst b[2];
			gs_build_pair(b, ck, (uint32_t) (n0 / GS_PAIR_SAMPLES), &q);
			int64_t lim = (n_frames - n0 < GS_PAIR_SAMPLES) ? n_frames - n0 : GS_PAIR_SAMPLES;
			for (int64_t m = 0; m < lim; m++)
				row[(n0 + m) * t_stride] = gs_sample(b, ck, (uint32_t) (n0 + m), (int32_t) m, &q);
		}
	}
	return 0;
}

#endif

/* gais_synth_host.c -- libgais_synth.so: the workload generator's host entry point on its own (see synth_host_impl.h) */
#include <stdint.h>
#include "synth_host_impl.h"

int gais_synth_host(const gais_synth *p, uint32_t first_channel, int32_t n_channels, int64_t n_frames, int16_t *h_out, int32_t layout,
		    int64_t stride)
{
	return gs_fill_host(p, first_channel, n_channels, n_frames, h_out, layout, stride);
}

/*
 * gais_fir.cuh -- launch wrappers of the FIR-sign stage (K1).
 */
#ifndef GAIS_FIR_CUH
#define GAIS_FIR_CUH

#include "gais_kernels.cuh"

namespace gais {

static inline int fir_setup(void) { return 0; }

/* returns the number of kernels launched, < 0 on error */
static inline int fir_launch(int fir_mode, int layout, SampleView view, const ChanState *st, int hist_sel, int n_ch,
			     int64_t n_frames, uint32_t *signs, cudaStream_t stream)
{
	(void) fir_mode; (void) layout;
	dim3 grid((unsigned) ((n_ch + K1_CH - 1) / K1_CH), (unsigned) ((n_frames + K1_TILE - 1) / K1_TILE));
	fir_sign_exact_kernel<<<grid, K1_CH * 32, 0, stream>>>(view, st, hist_sel, n_ch, n_frames, signs);
	return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

} /* namespace gais */
#endif

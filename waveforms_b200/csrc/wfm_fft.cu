// wfm_fft.cu — K3: shared-memory Stockham FFT + frequency-domain filter.
#include <cuda_runtime.h>
#include "wfm_internal.h"

extern "C" int wfm_fft_filter(const double* x, double* y, int64_t n_sig, int64_t n, int64_t stride, const double* H,
                              void* stream) {
  (void)x; (void)y; (void)n_sig; (void)n; (void)stride; (void)H; (void)stream;
  return WFM_EUNSUPPORTED;
}
extern "C" int wfm_fft_c2c(double* data, int64_t n_sig, int64_t n, int64_t stride, int32_t sign, void* stream) {
  (void)data; (void)n_sig; (void)n; (void)stride; (void)sign; (void)stream;
  return WFM_EUNSUPPORTED;
}

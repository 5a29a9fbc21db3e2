// wfm_iir.cu — K2: cascaded-biquad IIR (scipy.signal.sosfilt semantics).
#include <cuda_runtime.h>
#include "wfm_internal.h"

extern "C" int wfm_sosfilt(const double* sos, int32_t n_sections, double initial, const double* x, double* y,
                           int64_t n_sig, int64_t n, int64_t stride, const double* zi, double* zf, void* stream) {
  (void)sos; (void)n_sections; (void)initial; (void)x; (void)y; (void)n_sig; (void)n; (void)stride; (void)zi; (void)zf; (void)stream;
  return WFM_EUNSUPPORTED;
}

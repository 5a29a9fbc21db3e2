/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * Extended-precision (x87 long double, 64-bit mantissa) restatement of the IIR
 * recurrences the reference applies through SciPy:
 *   scipy.signal.sosfilt  (waveforms/waveform.py:200-203, :249)  cascaded biquads, DF2T
 *   scipy.signal.lfilter  (waveforms/distortion.py:321)          one DF2T section of order M
 * Same operation order as SciPy's C loops, every product and sum kept in long double:
 * the "truth" against which bench.py and the tests measure how far the float64
 * results (SciPy's, the GPU's exact mode, the GPU's scan mode) are from the exact
 * filter output — the exp-decay filters have poles at 0.995..0.998, whose rounding-noise
 * gain makes SciPy's own float64 result uncertain at a few 1e-12.
 *
 * Build: gcc -O2 -fPIC -shared -o oracle/_c/libwfm_ld.so oracle/csrc/ld_filters.c
 */
#include <stddef.h>

/* y = sosfilt(sos, x): sos[n_sec][6] = b0 b1 b2 a0 a1 a2 (a0 == 1), zero initial state */
void wfm_ld_sosfilt(const double* sos, int n_sec, const double* x, double* y, long n) {
  long double z0[16] = {0}, z1[16] = {0};
  if (n_sec > 16) n_sec = 16;
  for (long i = 0; i < n; ++i) {
    long double v = x[i];
    for (int s = 0; s < n_sec; ++s) {
      const long double b0 = sos[6 * s], b1 = sos[6 * s + 1], b2 = sos[6 * s + 2];
      const long double a1 = sos[6 * s + 4], a2 = sos[6 * s + 5];
      const long double out = b0 * v + z0[s];
      z0[s] = b1 * v - a1 * out + z1[s];
      z1[s] = b2 * v - a2 * out;
      v = out;
    }
    y[i] = (double)v;
  }
}

/* y = lfilter(b, a, x, zi): order M = max(nb, na) - 1 <= 32, coefficients normalised by a[0] */
void wfm_ld_lfilter(const double* b, int nb, const double* a, int na, const double* x, double* y, long n,
                    const double* zi) {
  long double bb[33] = {0}, aa[33] = {0}, z[33] = {0};
  int m = (nb > na ? nb : na) - 1;
  if (m > 32) m = 32;
  for (int k = 0; k < nb && k <= 32; ++k) bb[k] = (long double)b[k] / a[0];
  for (int k = 0; k < na && k <= 32; ++k) aa[k] = (long double)a[k] / a[0];
  if (zi) for (int k = 0; k < m; ++k) z[k] = zi[k];
  for (long i = 0; i < n; ++i) {
    const long double v = x[i];
    const long double out = (m > 0 ? z[0] : 0.0L) + bb[0] * v;
    for (int k = 0; k + 1 < m; ++k) z[k] = z[k + 1] + bb[k + 1] * v - aa[k + 1] * out;
    if (m > 0) z[m - 1] = bb[m] * v - aa[m] * out;
    y[i] = (double)out;
  }
}

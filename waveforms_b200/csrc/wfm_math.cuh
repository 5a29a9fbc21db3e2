// wfm_math.cuh — correctly-rounded, never-contracted fp64 primitives.
// NumPy / SciPy evaluate a*b+c as two rounded operations; these wrappers keep
// the compiler from fusing them (see wfm_basis.cuh, wfm_iir.cu).
#pragma once

namespace wfm {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }

// sin and cos of one argument with a Cody-Waite reduction that stays on the fast
// path up to |a| = 2^30 (CUDA's sincos() leaves it at 105615 rad and then calls a
// Payne-Hanek routine with a local-memory scratch array: a 200 MHz carrier passes
// that threshold 84 us into a frame).  pi/2 = HI + MID + LO (round-to-nearest
// splits); q*HI is subtracted exactly by the first FMA (the difference is a
// multiple of 2^-52 below 2), the next two carry the tail, so the reduced
// argument has a relative error of ~2^-53 even next to a multiple of pi/2.
// Kernels: the fdlibm minimax polynomials on [-pi/4, pi/4].  Measured against
// mpmath over +-1e9 (incl. neighbours of k*pi/2): relative error < 2 * 2^-53.
static __device__ __noinline__ void sincos_huge(double a, double* sn, double* cs) { sincos(a, sn, cs); }

__device__ __forceinline__ void sincos_cw(double a, double* sn, double* cs) {
  if (!(fabs(a) <= 1073741824.0)) {  // also NaN / inf: CUDA's Payne-Hanek path, out of line
    sincos_huge(a, sn, cs);
    return;
  }
  const double q = rint(a * 0.6366197723675814);
  double r = fma(-q, 1.5707963267948966, a);
  r = fma(-q, 6.123233995736766e-17, r);
  r = fma(-q, -1.4973849048591698e-33, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double s = fma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double hz = 0.5 * z;
  const double w = 1.0 - hz;
  const double c = w + (((1.0 - w) - hz) + z * z * pc);
  const int n = (int)q;
  const double ss = (n & 1) ? c : s;
  const double cc = (n & 1) ? s : c;
  *sn = (n & 2) ? -ss : ss;
  *cs = ((n + 1) & 2) ? -cc : cc;
}

__device__ __forceinline__ double sin_cw(double a) {
  double s, c;
  sincos_cw(a, &s, &c);
  return s;
}

__device__ __forceinline__ double cos_cw(double a) {
  double s, c;
  sincos_cw(a, &s, &c);
  return c;
}

}  // namespace wfm

// wfm_basis.cuh — device implementations of the reference's basis functions.
//
// Each function receives t = x - shift (ONE fp64 subtraction, as
// /root/reference/waveforms/_waveform.pyx:130-131 hands `x - shift` to the
// basis function) and reproduces the reference's operation ORDER on the
// argument path.  All parity-critical arithmetic uses __dmul_rn/__dadd_rn/
// __dsub_rn/__ddiv_rn, which the compiler never contracts into FMAs: NumPy
// evaluates `w * t`, `t / s`, `a*t**2 + b*t` ... as separate correctly-rounded
// ufunc passes, and at carrier phases of 1e4..1e5 rad a fused or re-associated
// argument moves the result by ~1e-11 (SURVEY.md §7 "Argument path").
//
// Argument packing (a0, a1, pool) is produced by waveforms_b200/lowering.py
// (PACKERS) and multy_drag.py; keep the two in sync.
#pragma once
#include <math_constants.h>
#include "../../include/wfm_b200.h"
#include "wfm_math.cuh"
#include "wfm_erf_table.h"

namespace wfm {

constexpr double kPi = 3.141592653589793;       // numpy.pi
constexpr double kTwoPi = 6.283185307179586;    // 2 * numpy.pi

// value ** n for a small non-zero integer n (np.power fast paths: 2 -> square,
// -1 -> reciprocal; the rest agree with libm pow to a few ulp).
__device__ __forceinline__ double pow_small_int(double v, int n) {
  unsigned m = n < 0 ? (unsigned)(-n) : (unsigned)n;
  double r = 1.0, b = v;
  bool first = true;
  while (m) {
    if (m & 1u) { r = first ? b : mul(r, b); first = false; }
    m >>= 1;
    if (m) b = mul(b, b);
  }
  return n < 0 ? dvd(1.0, r) : r;
}

// _waveform.pyx:294-295  np.exp(-(t / std_sq2)**2)
__device__ __forceinline__ double f_gaussian(double t, double s) {
  double u = dvd(t, s);
  return exp(-mul(u, u));
}

// _waveform.pyx:303-304  scipy.special.erf(t / std_sq2): piecewise Taylor table
// (wfm_erf_table.h; 24 intervals on [0, 6), degree 12, within 1 ulp of 1 of erf).
// ~40 instructions against ~125 for CUDA's erf(), whose 64-bit immediates cost two
// moves per FMA; the coefficient loads are independent of the FMA chain.
// tab: kErfTab itself (global memory, L1-cached) or a shared-memory copy of it
__device__ __forceinline__ double erf_tab(double x, const double* __restrict__ tab = &kErfTab[0][0]) {
  const double ax = fabs(x);
  double r;
  if (ax < 6.0) {
    const int k = (int)(ax * 4.0);
    const double y = ax - ((double)k * 0.25 + 0.125);
    const double* __restrict__ c = tab + k * (kErfDegree + 1);
    r = c[kErfDegree];
#pragma unroll
    for (int j = kErfDegree - 1; j >= 0; --j) r = fma(r, y, c[j]);
  } else {
    r = (ax != ax) ? ax : 1.0;
  }
  return copysign(r, x);
}

// _waveform.pyx:311-312  np.sinc(bw*t): y = pi*where(x==0, 1e-20, x); sin(y)/y
__device__ __forceinline__ double f_sinc(double t, double bw) {
  double u = mul(bw, t);
  double y = mul(kPi, u == 0.0 ? 1.0e-20 : u);
  return dvd(sin_cw(y), y);
}

// _waveform.pyx:319-320  np.interp(t, np.linspace(start, stop, n), points)
// pool: [n, step, points...].  xp[j] = j*step + start (unfused), xp[n-1] = stop.
__device__ __forceinline__ double interp_xp(int j, int n, double step, double start, double stop) {
  return (j == n - 1 && n > 1) ? stop : add(mul((double)j, step), start);
}
__device__ inline double f_interp(double t, double start, double stop, const double* __restrict__ pool) {
  const int n = (int)pool[0];
  const double step = pool[1];
  const double* fp = pool + 2;
  if (t != t) return t;
  if (n == 1) return fp[0];
  if (t < start) return fp[0];          // left  = fp[0]
  if (t > stop) return fp[n - 1];       // right = fp[-1]
  // largest j with xp[j] <= t (np.interp's binary search result)
  int j = (step > 0.0) ? (int)floor(dvd(sub(t, start), step)) : 0;
  j = max(0, min(j, n - 1));
  while (j > 0 && interp_xp(j, n, step, start, stop) > t) --j;
  while (j < n - 1 && interp_xp(j + 1, n, step, start, stop) <= t) ++j;
  double xj = interp_xp(j, n, step, start, stop);
  if (j == n - 1 || xj == t) return fp[j];
  double xj1 = interp_xp(j + 1, n, step, start, stop);
  double slope = dvd(sub(fp[j + 1], fp[j]), sub(xj1, xj));
  double r = add(mul(slope, sub(t, xj)), fp[j]);
  if (r != r) {  // numpy retries from the right knot when the left gives NaN
    r = add(mul(slope, sub(t, xj1)), fp[j + 1]);
    if (r != r && fp[j] == fp[j + 1]) r = fp[j];
  }
  return r;
}

// _waveform.pyx:343-356.  a0 = t0, a1 = o = pi/width,
// pool: [k1 = 2pi(freq+delta), k2 = 2pi*delta*t0 + phase, has_y, c3 = -b*o, o2 = 2*o]
__device__ __forceinline__ double f_drag(double t, double t0, double o, const double* __restrict__ pool) {
  double dt = sub(t, t0);
  double s1 = sin_cw(mul(o, dt));
  double ox = mul(s1, s1);
  double wt = sub(mul(pool[0], t), pool[1]);
  const SinCos sc = sincos_cw(wt);
  const double sw = sc.s, cw = sc.c;
  if (pool[2] == 0.0) return mul(ox, cw);
  double oy = mul(pool[3], sin_cw(mul(pool[4], dt)));
  return add(mul(ox, cw), mul(oy, sw));
}

// _waveform.pyx:359-371.  a0 = r, a1 = d; pool (d > 0): [r**d, ncoef, coefs...]
__device__ inline double f_mollifier(double t, double r, double dd, const double* __restrict__ pool) {
  double u = dvd(t, r);
  double au = fabs(u);
  double q = sub(mul(au, au), 1.0);
  const int d = (int)dd;
  if (d == 0) return q >= 0.0 ? 0.0 : exp(add(dvd(1.0, q), 1.0));
  double base = 0.0;
  if (!(q >= 0.0)) {
    double den = (d == 1) ? mul(-q, -q) : pow(-q, (double)(2 * d));
    base = dvd(exp(add(dvd(1.0, q), 1.0)), den);
  }
  const int nc = (int)pool[1];
  double p = 0.0;  // np.polyval: y = y*x + c, two ufunc passes per step
  for (int k = 0; k < nc; ++k) p = add(mul(p, u), pool[2 + k]);
  return dvd(mul(base, p), pool[0]);
}

// _waveform.pyx:298-300: (-1)**n / s**n * hermite(n)(t/s) * exp(-(t/s)**2).
// scipy.special.hermite(n)(x) dispatches to eval_hermite: the three-term
// recurrence of He_n at sqrt(2)*x, scaled by 2**(n/2).  a0 = s, a1 = n,
// pool: [c = (-1)**n / s**n]
__device__ inline double f_dgaussian(double t, double s, double nn, const double* __restrict__ pool) {
  const int n = (int)nn;
  double u = dvd(t, s);
  double h;
  if (n == 0) {
    h = 1.0;
  } else {
    double xs = mul(1.4142135623730951, u);
    if (n == 1) {
      h = xs;
    } else {
      double y3 = 0.0, y2 = 1.0, y1;
      for (int k = n; k > 1; --k) {
        y1 = sub(mul(xs, y2), mul((double)k, y3));
        y3 = y2;
        y2 = y1;
      }
      h = sub(mul(xs, y2), y3);
    }
    h = mul(h, exp2(0.5 * (double)n));
  }
  return mul(mul(pool[0], h), exp(-mul(u, u)));
}

// multi-notch DRAG envelopes, ids 16/17 (multy_drag.py:30-174); see
// wfm_multidrag.cuh
__device__ inline double f_drag_sin(double t, double t0, double o, const double* __restrict__ pool, bool sinx);

// one basis-function row: func id, shift, the two inline scalars, offset of its
// block in the argument pool (WfmFactor / DFactor carry exactly these)
struct FacArgs {
  int func;
  int arg_off;
  double shift, a0, a1;
};

// every basis function; out of line (one copy per kernel): the sampling kernel
// inlines the hot ones itself (eval_factors in wfm_sample.cu)
static __device__ __noinline__ double eval_factor(const FacArgs& f, double x, const double* __restrict__ args) {
  const double t = sub(x, f.shift);
  switch (f.func) {
    case WFM_LINEAR: return t;
    case WFM_GAUSSIAN: return f_gaussian(t, f.a0);
    case WFM_ERF: return erf_tab(dvd(t, f.a0));
    case WFM_COS: return cos_cw(mul(f.a0, t));
    case WFM_SINC: return f_sinc(t, f.a0);
    case WFM_EXP: return exp(mul(f.a0, t));
    case WFM_INTERP: return f_interp(t, f.a0, f.a1, args + f.arg_off);
    case WFM_LINEARCHIRP: {
      // sin(phi0 + 2pi*((f1-f0)/(2T)*t**2 + f0*t)); a0 = f0, a1 = phi0, pool [c, 2pi]
      const double* p = args + f.arg_off;
      double inner = add(mul(p[0], mul(t, t)), mul(f.a0, t));
      return sin_cw(add(f.a1, mul(p[1], inner)));
    }
    case WFM_EXPONENTIALCHIRP: {
      // sin(phi0 + 2pi*f0*(exp(alpha*t)-1)/alpha); a0 = alpha, a1 = phi0, pool [2pi*f0]
      const double* p = args + f.arg_off;
      double g = sub(exp(mul(f.a0, t)), 1.0);
      return sin_cw(add(f.a1, dvd(mul(p[0], g), f.a0)));
    }
    case WFM_HYPERBOLICCHIRP: {
      // sin(phi0 + 2pi*f0/k*log(1+k*t)); a0 = k, a1 = phi0, pool [2pi*f0/k]
      const double* p = args + f.arg_off;
      return sin_cw(add(f.a1, mul(p[0], log(add(1.0, mul(f.a0, t))))));
    }
    case WFM_COSH: return cosh(mul(f.a0, t));
    case WFM_SINH: return sinh(mul(f.a0, t));
    case WFM_DRAG: return f_drag(t, f.a0, f.a1, args + f.arg_off);
    case WFM_MOLLIFIER: return f_mollifier(t, f.a0, f.a1, args + f.arg_off);
    case WFM_D_GAUSSIAN: return f_dgaussian(t, f.a0, f.a1, args + f.arg_off);
    case WFM_DRAG_SIN: return f_drag_sin(t, f.a0, f.a1, args + f.arg_off, false);
    case WFM_DRAG_SINX: return f_drag_sin(t, f.a0, f.a1, args + f.arg_off, true);
    default: return CUDART_NAN;
  }
}

}  // namespace wfm

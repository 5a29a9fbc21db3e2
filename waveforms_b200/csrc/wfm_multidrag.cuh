// wfm_multidrag.cuh — device side of the multi-notch DRAG envelopes
// (ids 16 DRAG_SIN / 17 DRAG_SINX; /root/reference/waveforms/multy_drag.py:30-174).
//
// The reference builds, per call, the notch matrices B (multy_drag.py:9-15),
// the derivative table A of sin^m (:18-27), evaluates
//     d_i(t) = sum_p A[i,p] * S^p * (C if p odd)        S = sin(o(t-t0)), C = cos(..)
//     Omega_j(t) = sum_i B[i,j,0] * d_i(t)  (/ coeff for DRAG_SIN)
// and, for DRAG_SINX, replaces d_i inside the two "tab" windows by the i-th
// derivative of a fitted polynomial.  Everything sample-independent is folded on
// the host (waveforms_b200/multy_drag.py: pack_drag_sin / pack_drag_sinx):
//     G[j,p] = sum_i B[i,j,0] A[i,p] (/coeff),  plateau constants P[j],
//     tab polynomial derivative rows.
// pool layout (a0 = t0, a1 = o = pi/width):
//   [0] k1 = 2pi(freq+delta)  [1] k2 = 2pi*delta*t0+phase  [2] tm1 = t0+width/2
//   [3] tm2 = t0+plateau+width/2  [4] plateau  [5] m  [6] Px  [7] Py
//   [8 .. 8+m] Gx   [9+m .. 9+2m] Gy
//   then: tl, tr, width/2, rows, and if rows > 0: L, B0[rows], B1[rows],
//         left rows[rows*L], right rows[rows*L]   (coefficients, highest first)
#pragma once
#include "wfm_basis.cuh"

namespace wfm {

__device__ __forceinline__ double polyval_rows(const double* __restrict__ c, int L, double x) {
  double y = 0.0;  // np.polyval: y = y*x + c, separate multiply and add
  for (int k = 0; k < L; ++k) y = add(mul(y, x), c[k]);
  return y;
}

__device__ inline double f_drag_sin(double t, double t0, double o, const double* __restrict__ pool, bool sinx) {
  const double tm1 = pool[2], tm2 = pool[3], plateau = pool[4];
  const int m = (int)pool[5];
  const double* __restrict__ gx = pool + 8;
  const double* __restrict__ gy = gx + (m + 1);
  const double* __restrict__ tb = gy + (m + 1);
  const double dt = sub(t, t0);
  double ox = 0.0, oy = 0.0;
  bool in_tab = false;
  if (sinx) {
    const int rows = (int)tb[3];
    if (rows > 0) {
      const double tl = tb[0], tr = tb[1], hw = tb[2];
      const int L = (int)tb[4];
      const double* __restrict__ b0 = tb + 5;
      const double* __restrict__ b1 = b0 + rows;
      const double* __restrict__ rows_p = nullptr;
      double tau = 0.0;
      if (t >= tm2 && t <= tr) {  // right window is assigned last in the reference loop
        rows_p = b1 + rows + (size_t)rows * L;
        tau = sub(sub(dt, plateau), hw);
      } else if (t >= tl && t <= tm1) {
        rows_p = b1 + rows;
        tau = sub(dt, hw);
      }
      if (rows_p) {
        in_tab = true;
        for (int n = 0; n < rows; ++n) {
          const double v = polyval_rows(rows_p + (size_t)n * L, L, tau);
          ox = add(ox, mul(b0[n], v));
          oy = add(oy, mul(b1[n], v));
        }
      }
    }
  }
  if (!in_tab) {
    if (t > tm1 && t < tm2) {
      ox = pool[6];
      oy = pool[7];
    } else {
      const double arg = (t >= tm2) ? mul(o, sub(dt, plateau)) : mul(o, dt);
      const SinCos sa = sincos_cw(arg);
      const double S = sa.s, Cc = sa.c;
      double sp = 1.0;
      for (int p = 0; p <= m; ++p) {
        const double basis = (p & 1) ? mul(sp, Cc) : sp;
        ox = add(ox, mul(gx[p], basis));
        oy = add(oy, mul(gy[p], basis));
        sp = mul(sp, S);
      }
    }
  }
  const double wt = sub(mul(pool[0], t), pool[1]);
  const SinCos sc = sincos_cw(wt);
  const double sw = sc.s, cw = sc.c;
  return add(mul(ox, cw), mul(oy, sw));
}

}  // namespace wfm

// wfm_iir.cu — K2: cascaded-biquad IIR, scipy.signal.sosfilt semantics.
//
// Replaces the sosfilt call sites /root/reference/waveforms/waveform.py:200-203,
// :249 and (through first/second-order sections) lfilter in
// /root/reference/waveforms/distortion.py:321.  Per section, direct form II
// transposed, exactly the recurrence scipy runs (scipy/signal/_sosfilt.pyx):
//     y  = b0*x + z0
//     z0 = b1*x - a1*y + z1
//     z1 = b2*x - a2*y
// with every product and sum rounded separately (no FMA; built with -fmad=false
// and written with __dmul_rn/__dadd_rn).
//
// Two kernels:
//
//  * WFM_IIR_EXACT — one thread per signal walks time sequentially with the
//    recurrence above.  Bit-identical to scipy.  Parallelism = n_sig.
//
//  * WFM_IIR_SCAN — block-parallel associative scan.  One CTA per signal walks
//    4096-sample tiles; inside a tile each of 256 threads owns 16 consecutive
//    samples.  The biquad is the affine map z' = A z + B x, A = [[-a1,1],[-a2,0]]
//    on its state z = (z0, z1); because A is the same for every sample, the scan
//    over thread chunks only has to carry 2-vectors and uses the precomputed
//    powers A^(16*2^d):
//       pass a: every thread runs its chunk from the zero state -> f_i
//               (thread 0 starts from the tile's carry-in state);
//       scan  : E_i = sum_{j<i} A^(16 (i-1-j)) f_j  (warp shuffles, then 8 warp
//               totals through shared memory);
//       pass c: every thread re-runs its chunk from E_i with the faithful
//               recurrence and writes y.
//    Cascaded sections are processed one after the other on the register-resident
//    tile.  The result differs from the sequential one only through the rounding
//    of the carried states: ~1e-16 * (filter noise gain).  For poles within 5e-3
//    of the unit circle (the exp-decay predistortion filters) that noise gain
//    makes scipy's OWN sequential result uncertain at the 3e-12 level, which no
//    re-association can track; for well-conditioned filters the
//    two modes agree to ~1e-15.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "wfm_internal.h"
#include "wfm_math.cuh"

namespace wfm {

#ifndef WFM_IIR_THREADS
#define WFM_IIR_THREADS 256
#endif
#ifndef WFM_IIR_T
#define WFM_IIR_T 16
#endif
#ifndef WFM_IIR_MINB
#define WFM_IIR_MINB 2
#endif
constexpr int kIirThreads = WFM_IIR_THREADS;
constexpr int kIirT = WFM_IIR_T;                   // samples per thread per tile
constexpr int kIirTile = kIirThreads * kIirT;      // 4096
constexpr int kIirRow = kIirT + 1;                 // padded smem row (bank spread)
constexpr int kMaxSections = 8;

struct Biquad {
  double b0, b1, b2, a1, a2;
};

struct IirParams {
  int n_sections;
  int shift_in, shift_out;  // joint scan only: subtract `initial` at the input / add it at the output of THIS launch
  double initial;
  Biquad sec[kMaxSections];
};

// powers of A used by the scan, per section; 2x2 row-major
struct IirScanTables {
  double lane[kMaxSections][32][4];  // A^(T*l), l = 0..31
  double lvl[kMaxSections][5][4];    // A^(T*2^d), d = 0..4
  double warp[kMaxSections][4];      // A^(T*32)
};

__device__ __forceinline__ void biquad_step(const Biquad& q, double x, double& z0, double& z1, double& y) {
  y = add(mul(q.b0, x), z0);
  z0 = add(sub(mul(q.b1, x), mul(q.a1, y)), z1);
  z1 = sub(mul(q.b2, x), mul(q.a2, y));
}

// ---------------------------------------------------------------------------
// exact: ONE WARP per signal, sequential in time.  The warp loads 32 consecutive
// samples with one coalesced access (the next block is already in flight), every lane
// runs the same recurrence on the broadcast samples (the state is replicated, so no
// lane waits for another) and lane l keeps output l: loads and stores are full lines
// and the only serial cost left is the recurrence's own dependency chain.
// ---------------------------------------------------------------------------
constexpr int kExactWarps = 4;  // warps (= signals) per CTA

__device__ __forceinline__ double shfl_f64(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__global__ void __launch_bounds__(32 * kExactWarps) sosfilt_exact_kernel(const __grid_constant__ IirParams P,
                                                                         const double* __restrict__ x, double* y,
                                                                         int64_t n_sig, int64_t n, int64_t stride,
                                                                         const double* __restrict__ zi,
                                                                         double* __restrict__ zf) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * kExactWarps + (threadIdx.x >> 5);
  if (s >= n_sig) return;
  const int S = P.n_sections;
  double z0[kMaxSections], z1[kMaxSections];
#pragma unroll
  for (int k = 0; k < kMaxSections; ++k) {
    z0[k] = (zi && k < S) ? zi[(s * S + k) * 2 + 0] : 0.0;
    z1[k] = (zi && k < S) ? zi[(s * S + k) * 2 + 1] : 0.0;
  }
  const double* __restrict__ xs = x + s * stride;
  double* ys = y + s * stride;
  const bool shift = P.initial != 0.0;
  double nxt = lane < n ? xs[lane] : 0.0;
  for (int64_t base = 0; base < n; base += 32) {
    const double mine = nxt;
    if (base + 32 + lane < n) nxt = xs[base + 32 + lane];  // in flight while this block is filtered
    const int cnt = (int)min((int64_t)32, n - base);
    double out = 0.0;
    if (S == 1 && cnt == 32) {  // full blocks of the common cases: no branch between the samples, so the
      const Biquad q = P.sec[0];  // 32 broadcasts are issued ahead of the recurrence that consumes them
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        double cur = shfl_f64(mine, l), o;
        if (shift) cur = sub(cur, P.initial);
        biquad_step(q, cur, z0[0], z1[0], o);
        if (lane == l) out = o;
      }
    } else if (S == 2 && cnt == 32) {
      // software-pipelined across the sections: section 1 of sample l next to section 2 of sample l - 1 — two
      // independent dependency chains in flight, the same operations in the same order per section (bit-identical)
      const Biquad q0 = P.sec[0], q1 = P.sec[1];
      double o1_prev;
      {
        double cur = shfl_f64(mine, 0);
        if (shift) cur = sub(cur, P.initial);
        biquad_step(q0, cur, z0[0], z1[0], o1_prev);
      }
#pragma unroll
      for (int l = 1; l <= 32; ++l) {
        double o1 = 0.0, o2;
        if (l < 32) {
          double cur = shfl_f64(mine, l);
          if (shift) cur = sub(cur, P.initial);
          biquad_step(q0, cur, z0[0], z1[0], o1);
        }
        biquad_step(q1, o1_prev, z0[1], z1[1], o2);
        if (lane == l - 1) out = o2;
        o1_prev = o1;
      }
    } else {
#pragma unroll 1
      for (int l = 0; l < cnt; ++l) {
        double cur = shfl_f64(mine, l);
        if (shift) cur = sub(cur, P.initial);
#pragma unroll
        for (int k = 0; k < kMaxSections; ++k)
          if (k < S) {
            double o;
            biquad_step(P.sec[k], cur, z0[k], z1[k], o);
            cur = o;
          }
        if (lane == l) out = cur;
      }
    }
    if (lane < cnt) ys[base + lane] = shift ? add(out, P.initial) : out;
  }
  if (zf && lane == 0) {
#pragma unroll
    for (int k = 0; k < kMaxSections; ++k)
      if (k < S) {
        zf[(s * S + k) * 2 + 0] = z0[k];
        zf[(s * S + k) * 2 + 1] = z1[k];
      }
  }
}

// ---------------------------------------------------------------------------
// scan: one CTA per signal, tiles of 4096 samples
// ---------------------------------------------------------------------------
__device__ __forceinline__ void matvec(const double* __restrict__ M, double a, double b, double& ra, double& rb) {
  ra = fma(M[0], a, M[1] * b);
  rb = fma(M[2], a, M[3] * b);
}

__global__ void __launch_bounds__(kIirThreads, WFM_IIR_MINB) sosfilt_scan_kernel(const __grid_constant__ IirParams P,
                                                                        const __grid_constant__ IirScanTables TT,
                                                                        const double* __restrict__ x, double* y,
                                                                     int64_t n, int64_t stride,
                                                                     const double* __restrict__ zi,
                                                                     double* __restrict__ zf) {
  const IirScanTables* __restrict__ T = &TT;  // ~10 KB of kernel parameters (constant bank): no table upload per call
  __shared__ double s_tile[kIirThreads * kIirRow];
  __shared__ double s_tot[kIirThreads / 32][2];
  __shared__ double s_carry[kMaxSections][2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t sig = blockIdx.x;
  const double* __restrict__ xs = x + sig * stride;
  double* ys = y + sig * stride;
  const int S = P.n_sections;
  const bool shift = P.initial != 0.0;
  if (tid < S) {
    s_carry[tid][0] = zi ? zi[(sig * S + tid) * 2 + 0] : 0.0;
    s_carry[tid][1] = zi ? zi[(sig * S + tid) * 2 + 1] : 0.0;
  }
  __syncthreads();

  for (int64_t base = 0; base < n; base += kIirTile) {
    const int cnt = (int)min((int64_t)kIirTile, n - base);
    // coalesced load -> padded rows (thread r owns row r)
#pragma unroll
    for (int k = 0; k < kIirT; ++k) {
      const int e = k * kIirThreads + tid;
      double v = 0.0;
      if (e < cnt) {
        v = xs[base + e];
        if (shift) v = sub(v, P.initial);
      }
      s_tile[(e / kIirT) * kIirRow + (e % kIirT)] = v;
    }
    __syncthreads();
    double v[kIirT];
#pragma unroll
    for (int k = 0; k < kIirT; ++k) v[k] = s_tile[tid * kIirRow + k];
    const int valid = max(0, min(kIirT, cnt - tid * kIirT));  // my valid samples

    for (int sct = 0; sct < S; ++sct) {
      const Biquad q = P.sec[sct];
      const double c0 = s_carry[sct][0], c1 = s_carry[sct][1];
      // pass a: zero-state response of my chunk (thread 0 carries the tile state)
      double f0 = tid == 0 ? c0 : 0.0, f1 = tid == 0 ? c1 : 0.0;
#pragma unroll
      for (int k = 0; k < kIirT; ++k) {
        double out;
        biquad_step(q, v[k], f0, f1, out);
      }
      // inclusive warp scan of the affine maps (constant matrix per level)
#pragma unroll
      for (int d = 0; d < 5; ++d) {
        const double p0 = __shfl_up_sync(0xffffffffu, f0, 1 << d);
        const double p1 = __shfl_up_sync(0xffffffffu, f1, 1 << d);
        if (lane >= (1 << d)) {
          double r0, r1;
          matvec(T->lvl[sct][d], p0, p1, r0, r1);
          f0 += r0;
          f1 += r1;
        }
      }
      if (lane == 31) {
        s_tot[warp][0] = f0;
        s_tot[warp][1] = f1;
      }
      // state entering my chunk from the lanes before me in this warp
      double e0 = __shfl_up_sync(0xffffffffu, f0, 1);
      double e1 = __shfl_up_sync(0xffffffffu, f1, 1);
      if (lane == 0) e0 = e1 = 0.0;
      __syncthreads();
      // carry entering my warp: C_w = Q C_{w-1} + tot_{w-1}
      double w0 = 0.0, w1 = 0.0;
      for (int k = 0; k < warp; ++k) {
        double r0, r1;
        matvec(T->warp[sct], w0, w1, r0, r1);
        w0 = r0 + s_tot[k][0];
        w1 = r1 + s_tot[k][1];
      }
      if (warp > 0) {
        double r0, r1;
        matvec(T->lane[sct][lane], w0, w1, r0, r1);
        e0 += r0;
        e1 += r1;
      }
      if (tid == 0) { e0 = c0; e1 = c1; }
      // pass c: faithful recurrence from the carried-in state
      double z0 = e0, z1 = e1;
#pragma unroll
      for (int k = 0; k < kIirT; ++k) {
        if (k == valid && valid < kIirT && (tid * kIirT + k == cnt)) {
          // first padded sample of the signal's tail: this is the final state
          s_carry[sct][0] = z0;
          s_carry[sct][1] = z1;
        }
        double out;
        biquad_step(q, v[k], z0, z1, out);
        v[k] = out;
      }
      __syncthreads();  // all reads of s_carry/s_tot for this section are done
      if (cnt == kIirTile && tid == kIirThreads - 1) {
        s_carry[sct][0] = z0;
        s_carry[sct][1] = z1;
      }
      __syncthreads();
    }
    // registers -> padded rows -> coalesced store
#pragma unroll
    for (int k = 0; k < kIirT; ++k) s_tile[tid * kIirRow + k] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kIirT; ++k) {
      const int e = k * kIirThreads + tid;
      if (e < cnt) {
        double o = s_tile[(e / kIirT) * kIirRow + (e % kIirT)];
        ys[base + e] = shift ? add(o, P.initial) : o;
      }
    }
    __syncthreads();
  }
  if (zf && tid < S) {
    zf[(sig * S + tid) * 2 + 0] = s_carry[tid][0];
    zf[(sig * S + tid) * 2 + 1] = s_carry[tid][1];
  }
}

// ---------------------------------------------------------------------------
// scan, all sections at once (S <= 2): the cascade is ONE linear system on the joint state
// s = (z0, z1) of every section (D = 2S components): s' = M s + g x.  Per tile
//   pass a: every thread runs its 16 samples from the zero state in the STATE-SPACE form
//           z0' = -a1 z0 + (z1 + c1 x), z1' = -a2 z0 + c2 x  (c1 = b1 - a1 b0, c2 = b2 - a2 b0):
//           one FMA per step on the critical path instead of the four dependent operations of
//           the faithful recurrence; only the chunk's END state f_i leaves this pass, and f_i
//           only feeds the carries, whose rounding is the same noise as before;
//   scan  : E_i = sum_{j<i} (M^16)^(i-1-j) f_j with D x D powers of M^16 (host, long double);
//           the per-lane powers sit in shared memory;
//   pass c: the FAITHFUL recurrence from E_i, the sections software-pipelined (section 1 of
//           sample k+1 next to section 2 of sample k: two independent chains).
// One scan and two passes per tile instead of two scans and four passes, and a shorter chain
// in pass a: 0.63 -> ~0.3 ms on cfg4 (256 x 400 000 samples; its two exponential decays make ONE biquad section; a
// two-section cascade of the same size takes 0.50 ms).
// ---------------------------------------------------------------------------
constexpr int kMaxJoint = 4;  // D = 2 S, S <= 2
struct IirJointTables {
  double lane[32][kMaxJoint * kMaxJoint];  // (M^16)^l, l = 0..31, row-major D x D
  double lvl[5][kMaxJoint * kMaxJoint];    // (M^16)^(2^d)
  double warp[kMaxJoint * kMaxJoint];      // (M^16)^32
  double c1[2], c2[2];                     // state-space input gains per section
};

template <int D>
__device__ __forceinline__ void matvec_d(const double* __restrict__ M, const double (&a)[D], double (&r)[D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double acc = M[i * D] * a[0];
#pragma unroll
    for (int j = 1; j < D; ++j) acc = fma(M[i * D + j], a[j], acc);
    r[i] = acc;
  }
}

// Tile staging of the joint kernel: rows of 16 samples at a pitch of 18 doubles (144 B): thread r reads / writes its
// row with 16-byte accesses without bank conflicts, and 16-byte chunks stay aligned for cp.async.  Two buffers: the
// NEXT tile is already on its way (cp.async, no registers held) while this one is filtered — the one-buffer kernel
// spent three quarters of its time waiting for its own loads (ncu: long-scoreboard stalls on the first use of every
// loaded sample).
constexpr int kJointPitch = kIirT + 2;                                       // doubles per staged row
constexpr int kJointTileBytes = kIirThreads * kJointPitch * (int)sizeof(double);  // 36 864
extern __shared__ __align__(16) unsigned char iir_smem[];

__device__ __forceinline__ void cp_async16_zfill(void* dst_smem, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// kAligned: x and y rows start on 16-byte boundaries (cp.async loads, 16-byte stores); else plain 8-byte accesses
template <int S, bool kAligned>
__global__ void __launch_bounds__(kIirThreads, WFM_IIR_MINB) sosfilt_scan_joint_kernel(
    const __grid_constant__ IirParams P, const __grid_constant__ IirJointTables TT, const double* __restrict__ x, double* y,
    int64_t n, int64_t stride, const double* __restrict__ zi, double* __restrict__ zf, int zs) {
  // zs: state words between two signals in zi / zf (2 x the sections of the WHOLE cascade: a long cascade runs as a chain
  // of launches over its sections two by two, each on its slice of the states)
  constexpr int D = 2 * S;
  double* s_buf0 = reinterpret_cast<double*>(iir_smem);
  double* s_buf1 = reinterpret_cast<double*>(iir_smem + kJointTileBytes);
  __shared__ double s_lane[D * D][32];  // (M^16)^lane, entry-major: conflict-free
  __shared__ double s_tot[kIirThreads / 32][D];
  __shared__ double s_carry[D];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t sig = blockIdx.x;
  const double* __restrict__ xs = x + sig * stride;
  double* ys = y + sig * stride;
  const bool shift_in = P.initial != 0.0 && P.shift_in, shift_out = P.initial != 0.0 && P.shift_out;
  for (int e = tid; e < D * D * 32; e += kIirThreads) s_lane[e / 32][e % 32] = TT.lane[e % 32][(e / 32) / D * kMaxJoint + (e / 32) % D];
  if (tid < D) s_carry[tid] = zi ? zi[sig * zs + tid] : 0.0;  // zi[sig][section][2] is the joint state in order
  Biquad q[S];
  double c1[S], c2[S];
#pragma unroll
  for (int k = 0; k < S; ++k) {
    q[k] = P.sec[k];
    c1[k] = TT.c1[k];
    c2[k] = TT.c2[k];
  }

  // tile `base` -> buffer: 2048 chunks of two samples, chunk c of the tile at row c / 8, column 2 (c % 8)
  auto load_tile = [&](double* buf, int64_t base) {
    const int cnt = (int)min((int64_t)kIirTile, n - base);
#pragma unroll
    for (int j = 0; j < kIirT / 2; ++j) {
      const int c = j * kIirThreads + tid;
      double* dst = buf + (c >> 3) * kJointPitch + 2 * (c & 7);
      if (kAligned) {
        const int left = cnt - 2 * c;  // samples of this chunk that exist: the rest is zero-filled
        cp_async16_zfill(dst, left > 0 ? xs + base + 2 * c : xs + base, left >= 2 ? 16 : (left == 1 ? 8 : 0));
      } else {
        dst[0] = 2 * c < cnt ? xs[base + 2 * c] : 0.0;
        dst[1] = 2 * c + 1 < cnt ? xs[base + 2 * c + 1] : 0.0;
      }
    }
  };
  if (n > 0) load_tile(s_buf0, 0);
  if (kAligned) cp_async_commit();
  __syncthreads();

  int it = 0;
  for (int64_t base = 0; base < n; base += kIirTile, ++it) {
    const int cnt = (int)min((int64_t)kIirTile, n - base);
    double* buf = (it & 1) ? s_buf1 : s_buf0;
    if (base + kIirTile < n) load_tile((it & 1) ? s_buf0 : s_buf1, base + kIirTile);
    if (kAligned) {
      cp_async_commit();
      cp_async_wait<1>();  // everything but the tile just requested has landed
    }
    __syncthreads();
    double v[kIirT];
    {
      const double2* row = reinterpret_cast<const double2*>(buf + tid * kJointPitch);
#pragma unroll
      for (int k = 0; k < kIirT / 2; ++k) {
        const double2 d = row[k];
        v[2 * k] = d.x;
        v[2 * k + 1] = d.y;
      }
    }
    const int valid = max(0, min(kIirT, cnt - tid * kIirT));  // my valid samples
    if (shift_in) {
#pragma unroll
      for (int k = 0; k < kIirT; ++k) v[k] = k < valid ? sub(v[k], P.initial) : 0.0;
    }

    double carry[D];
#pragma unroll
    for (int i = 0; i < D; ++i) carry[i] = s_carry[i];
    // ---- pass a: end state of my chunk from the zero state (thread 0: from the tile's carry-in), state-space form
    double f[D];
#pragma unroll
    for (int i = 0; i < D; ++i) f[i] = tid == 0 ? carry[i] : 0.0;
#pragma unroll
    for (int k = 0; k < kIirT; ++k) {
      double in = v[k];
#pragma unroll
      for (int sct = 0; sct < S; ++sct) {
        const double z0 = f[2 * sct], z1 = f[2 * sct + 1];
        const double out = fma(q[sct].b0, in, z0);
        f[2 * sct] = fma(-q[sct].a1, z0, fma(c1[sct], in, z1));
        f[2 * sct + 1] = fma(-q[sct].a2, z0, c2[sct] * in);
        in = out;
      }
    }
    // ---- scan over the threads: inclusive warp scan with constant matrices per level
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      double p[D], r[D];
#pragma unroll
      for (int i = 0; i < D; ++i) p[i] = __shfl_up_sync(0xffffffffu, f[i], 1 << d);
      if (lane >= (1 << d)) {
        double m[D * D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) m[i * D + j] = TT.lvl[d][i * kMaxJoint + j];
        matvec_d<D>(m, p, r);
#pragma unroll
        for (int i = 0; i < D; ++i) f[i] += r[i];
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int i = 0; i < D; ++i) s_tot[warp][i] = f[i];
    }
    double e[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      e[i] = __shfl_up_sync(0xffffffffu, f[i], 1);
      if (lane == 0) e[i] = 0.0;
    }
    __syncthreads();
    // carry entering my warp: C_w = Q C_{w-1} + tot_{w-1}, Q = (M^16)^32
    if (warp > 0) {
      double w[D], r[D], mq[D * D];
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) mq[i * D + j] = TT.warp[i * kMaxJoint + j];
      for (int k = 0; k < warp; ++k) {
        matvec_d<D>(mq, w, r);
#pragma unroll
        for (int i = 0; i < D; ++i) w[i] = r[i] + s_tot[k][i];
      }
      double ml[D * D];
#pragma unroll
      for (int i = 0; i < D * D; ++i) ml[i] = s_lane[i][lane];
      matvec_d<D>(ml, w, r);
#pragma unroll
      for (int i = 0; i < D; ++i) e[i] += r[i];
    }
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) e[i] = carry[i];
    }
    // ---- pass c: the faithful recurrence from the carried-in state; section 2 runs one sample behind section 1
    double z[D];
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = e[i];
    const bool tail_here = valid < kIirT && tid * kIirT + valid == cnt;  // the signal ends inside my chunk
    double fin[D];
#pragma unroll
    for (int i = 0; i < D; ++i) fin[i] = 0.0;
    if (S == 1) {
#pragma unroll
      for (int k = 0; k < kIirT; ++k) {
        if (tail_here && k == valid) { fin[0] = z[0]; fin[1] = z[1]; }
        double out;
        biquad_step(q[0], v[k], z[0], z[1], out);
        v[k] = out;
      }
    } else {
      double o1_prev;
      {
        if (tail_here && valid == 0) { fin[0] = z[0]; fin[1] = z[1]; }
        biquad_step(q[0], v[0], z[0], z[1], o1_prev);
      }
#pragma unroll
      for (int k = 1; k <= kIirT; ++k) {
        double o1 = 0.0, o2;
        if (k < kIirT) {
          if (tail_here && k == valid) { fin[0] = z[0]; fin[1] = z[1]; }
          biquad_step(q[0], v[k], z[0], z[1], o1);
        }
        if (tail_here && k - 1 == valid) { fin[2 % D] = z[2 % D]; fin[3 % D] = z[3 % D]; }
        biquad_step(q[S - 1], o1_prev, z[2 % D], z[3 % D], o2);
        v[k - 1] = o2;
        o1_prev = o1;
      }
    }
    if (shift_out) {
#pragma unroll
      for (int k = 0; k < kIirT; ++k) v[k] = add(v[k], P.initial);
    }
    // my row back into the buffer (nobody else reads it before the barrier), then coalesced 16-byte stores
    {
      double2* row = reinterpret_cast<double2*>(buf + tid * kJointPitch);
#pragma unroll
      for (int k = 0; k < kIirT / 2; ++k) row[k] = make_double2(v[2 * k], v[2 * k + 1]);
    }
    __syncthreads();  // rows complete; all reads of s_carry / s_tot are done
    if (tail_here) {
#pragma unroll
      for (int i = 0; i < D; ++i) s_carry[i] = fin[i];
    }
    if (cnt == kIirTile && tid == kIirThreads - 1) {
#pragma unroll
      for (int i = 0; i < D; ++i) s_carry[i] = z[i];
    }
#pragma unroll
    for (int j = 0; j < kIirT / 2; ++j) {
      const int c = j * kIirThreads + tid;
      const double2 d = *reinterpret_cast<const double2*>(buf + (c >> 3) * kJointPitch + 2 * (c & 7));
      if (kAligned && 2 * c + 1 < cnt) {
        *reinterpret_cast<double2*>(ys + base + 2 * c) = d;
      } else {
        if (2 * c < cnt) ys[base + 2 * c] = d.x;
        if (2 * c + 1 < cnt) ys[base + 2 * c + 1] = d.y;
      }
    }
    __syncthreads();  // the buffer may be refilled (tile t + 2) and s_carry read (tile t + 1)
  }
  if (kAligned) cp_async_wait<0>();
  if (zf && tid < D) zf[sig * zs + tid] = s_carry[tid];
}

// ---------------------------------------------------------------------------
// lfilter: single high-order section, scipy.signal.lfilter's DF2T loop
// (scipy/signal/_lfilter.c.src): y = z[0] + b[0]*x;
// z[k] = (z[k+1] + x*b[k+1]) - y*a[k+1];  z[M-1] = x*b[M] - y*a[M].
// One thread per signal, sequential in time, bit-faithful (note the different
// association from sosfilt's (b1*x - a1*y) + z1).
// ---------------------------------------------------------------------------
constexpr int kMaxOrder = 16;
struct LfilterParams {
  int order;  // M = max(len(a), len(b)) - 1
  double b[kMaxOrder + 1];
  double a[kMaxOrder + 1];
};

// one lfilter step on the replicated state (M = order)
template <int M>
__device__ __forceinline__ double lfilter_step(const LfilterParams& P, double xv, double (&z)[kMaxOrder]) {
  if (M == 0) return mul(P.b[0], xv);
  const double yv = add(z[0], mul(P.b[0], xv));
#pragma unroll
  for (int k = 0; k < M - 1; ++k) z[k] = sub(add(z[k + 1], mul(xv, P.b[k + 1])), mul(yv, P.a[k + 1]));
  z[M - 1] = sub(mul(xv, P.b[M]), mul(yv, P.a[M]));
  return yv;
}
__device__ __forceinline__ double lfilter_step_any(const LfilterParams& P, double xv, double (&z)[kMaxOrder]) {
  const int M = P.order;
  if (M == 0) return mul(P.b[0], xv);
  const double yv = add(z[0], mul(P.b[0], xv));
#pragma unroll
  for (int k = 0; k < kMaxOrder - 1; ++k)
    if (k < M - 1) z[k] = sub(add(z[k + 1], mul(xv, P.b[k + 1])), mul(yv, P.a[k + 1]));
#pragma unroll
  for (int k = 0; k < kMaxOrder; ++k)
    if (k == M - 1) z[k] = sub(mul(xv, P.b[M]), mul(yv, P.a[M]));
  return yv;
}

// ONE WARP per signal (see sosfilt_exact_kernel): coalesced block loads, the recurrence
// replicated in every lane, lane l keeps output l
__global__ void __launch_bounds__(32 * kExactWarps) lfilter_exact_kernel(const __grid_constant__ LfilterParams P,
                                                                         const double* __restrict__ x, double* y,
                                                                         int64_t n_sig, int64_t n, int64_t stride,
                                                                         const double* __restrict__ zi,
                                                                         double* __restrict__ zf) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * kExactWarps + (threadIdx.x >> 5);
  if (s >= n_sig) return;
  const int M = P.order;
  double z[kMaxOrder];
#pragma unroll
  for (int k = 0; k < kMaxOrder; ++k) z[k] = (zi && k < M) ? zi[s * M + k] : 0.0;
  const double* __restrict__ xs = x + s * stride;
  double* ys = y + s * stride;
  double nxt = lane < n ? xs[lane] : 0.0;
  for (int64_t base = 0; base < n; base += 32) {
    const double mine = nxt;
    if (base + 32 + lane < n) nxt = xs[base + 32 + lane];
    const int cnt = (int)min((int64_t)32, n - base);
    double out = 0.0;
    if (M == 1 && cnt == 32) {  // full blocks: no branch between the samples (see sosfilt_exact_kernel)
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        const double o = lfilter_step<1>(P, shfl_f64(mine, l), z);
        if (lane == l) out = o;
      }
    } else if (M == 2 && cnt == 32) {
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        const double o = lfilter_step<2>(P, shfl_f64(mine, l), z);
        if (lane == l) out = o;
      }
    } else {
#pragma unroll 1
      for (int l = 0; l < cnt; ++l) {
        const double o = lfilter_step_any(P, shfl_f64(mine, l), z);
        if (lane == l) out = o;
      }
    }
    if (lane < cnt) ys[base + lane] = out;
  }
  if (zf && lane == 0) {
#pragma unroll
    for (int k = 0; k < kMaxOrder; ++k)
      if (k < M) zf[s * M + k] = z[k];
  }
}


// ---------------------------------------------------------------------------
// lfilter, block-parallel (orders 1..4): the DF2T recurrence above is a linear system on its M state words,
// z' = A z + g x with A[i][0] = -a[i+1], A[i][i+1] = 1, so the machinery of the joint sosfilt scan applies unchanged:
//   pass a: every thread runs its 16 samples from the zero state in the state-space form
//           z_i' = -a_{i+1} z_0 + (z_{i+1} + c_{i+1} x), c_j = b_j - a_j b_0 (one FMA per step on the critical path);
//   scan  : carries with M x M powers of A^16 (host, long double);
//   pass c: scipy's own recurrence (lfilter_step: same operations, same association) from the carried-in state.
// predistort() (distortion.py:289-337) filters every flux channel with ONE combined high-order (b, a): the sequential
// kernel needs 13.4 ms for cfg4's 256 x 400 000 samples, this one the time of a read and a write.
// ---------------------------------------------------------------------------
struct LfScanParams {
  double b[kMaxJoint + 1], a[kMaxJoint + 1], c[kMaxJoint + 1];  // normalised by a[0]; c[j] = b[j] - a[j] b[0]
};
template <int M>
__device__ __forceinline__ double lfilter_step_scan(const LfScanParams& P, double xv, double (&z)[M]) {
  const double yv = add(z[0], mul(P.b[0], xv));
#pragma unroll
  for (int k = 0; k < M - 1; ++k) z[k] = sub(add(z[k + 1], mul(xv, P.b[k + 1])), mul(yv, P.a[k + 1]));
  z[M - 1] = sub(mul(xv, P.b[M]), mul(yv, P.a[M]));
  return yv;
}
template <int M, bool kAligned>
__global__ void __launch_bounds__(kIirThreads, WFM_IIR_MINB) lfilter_scan_kernel(
    const __grid_constant__ LfScanParams P, const __grid_constant__ IirJointTables TT, const double* __restrict__ x, double* y,
    int64_t n, int64_t stride, const double* __restrict__ zi, double* __restrict__ zf) {
  constexpr int D = M;
  double* s_buf0 = reinterpret_cast<double*>(iir_smem);
  double* s_buf1 = reinterpret_cast<double*>(iir_smem + kJointTileBytes);
  __shared__ double s_lane[D * D][32];  // (A^16)^lane, entry-major: conflict-free
  __shared__ double s_tot[kIirThreads / 32][D];
  __shared__ double s_carry[D];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t sig = blockIdx.x;
  const double* __restrict__ xs = x + sig * stride;
  double* ys = y + sig * stride;
  for (int e = tid; e < D * D * 32; e += kIirThreads) s_lane[e / 32][e % 32] = TT.lane[e % 32][(e / 32) / D * kMaxJoint + (e / 32) % D];
  if (tid < D) s_carry[tid] = zi ? zi[sig * D + tid] : 0.0;

  auto load_tile = [&](double* buf, int64_t base) {
    const int cnt = (int)min((int64_t)kIirTile, n - base);
#pragma unroll
    for (int j = 0; j < kIirT / 2; ++j) {
      const int c = j * kIirThreads + tid;
      double* dst = buf + (c >> 3) * kJointPitch + 2 * (c & 7);
      if (kAligned) {
        const int left = cnt - 2 * c;
        cp_async16_zfill(dst, left > 0 ? xs + base + 2 * c : xs + base, left >= 2 ? 16 : (left == 1 ? 8 : 0));
      } else {
        dst[0] = 2 * c < cnt ? xs[base + 2 * c] : 0.0;
        dst[1] = 2 * c + 1 < cnt ? xs[base + 2 * c + 1] : 0.0;
      }
    }
  };
  if (n > 0) load_tile(s_buf0, 0);
  if (kAligned) cp_async_commit();
  __syncthreads();

  int it = 0;
  for (int64_t base = 0; base < n; base += kIirTile, ++it) {
    const int cnt = (int)min((int64_t)kIirTile, n - base);
    double* buf = (it & 1) ? s_buf1 : s_buf0;
    if (base + kIirTile < n) load_tile((it & 1) ? s_buf0 : s_buf1, base + kIirTile);
    if (kAligned) {
      cp_async_commit();
      cp_async_wait<1>();
    }
    __syncthreads();
    double v[kIirT];
    {
      const double2* row = reinterpret_cast<const double2*>(buf + tid * kJointPitch);
#pragma unroll
      for (int k = 0; k < kIirT / 2; ++k) {
        const double2 d = row[k];
        v[2 * k] = d.x;
        v[2 * k + 1] = d.y;
      }
    }
    const int valid = max(0, min(kIirT, cnt - tid * kIirT));
    double carry[D];
#pragma unroll
    for (int i = 0; i < D; ++i) carry[i] = s_carry[i];
    // ---- pass a: end state of my chunk from the zero state (thread 0: from the tile's carry-in)
    double f[D];
#pragma unroll
    for (int i = 0; i < D; ++i) f[i] = tid == 0 ? carry[i] : 0.0;
#pragma unroll
    for (int k = 0; k < kIirT; ++k) {
      const double in = v[k], z0 = f[0];
#pragma unroll
      for (int i = 0; i < D - 1; ++i) f[i] = fma(-P.a[i + 1], z0, fma(P.c[i + 1], in, f[i + 1]));
      f[D - 1] = fma(-P.a[D], z0, P.c[D] * in);
    }
    // ---- scan over the threads
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      double p[D], r[D];
#pragma unroll
      for (int i = 0; i < D; ++i) p[i] = __shfl_up_sync(0xffffffffu, f[i], 1 << d);
      if (lane >= (1 << d)) {
        double m[D * D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
          for (int j = 0; j < D; ++j) m[i * D + j] = TT.lvl[d][i * kMaxJoint + j];
        matvec_d<D>(m, p, r);
#pragma unroll
        for (int i = 0; i < D; ++i) f[i] += r[i];
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int i = 0; i < D; ++i) s_tot[warp][i] = f[i];
    }
    double e[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      e[i] = __shfl_up_sync(0xffffffffu, f[i], 1);
      if (lane == 0) e[i] = 0.0;
    }
    __syncthreads();
    if (warp > 0) {
      double w[D], r[D], mq[D * D];
#pragma unroll
      for (int i = 0; i < D; ++i) w[i] = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) mq[i * D + j] = TT.warp[i * kMaxJoint + j];
      for (int k = 0; k < warp; ++k) {
        matvec_d<D>(mq, w, r);
#pragma unroll
        for (int i = 0; i < D; ++i) w[i] = r[i] + s_tot[k][i];
      }
      double ml[D * D];
#pragma unroll
      for (int i = 0; i < D * D; ++i) ml[i] = s_lane[i][lane];
      matvec_d<D>(ml, w, r);
#pragma unroll
      for (int i = 0; i < D; ++i) e[i] += r[i];
    }
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < D; ++i) e[i] = carry[i];
    }
    // ---- pass c: scipy's recurrence from the carried-in state
    double z[D];
#pragma unroll
    for (int i = 0; i < D; ++i) z[i] = e[i];
    const bool tail_here = valid < kIirT && tid * kIirT + valid == cnt;  // the signal ends inside my chunk
    double fin[D];
#pragma unroll
    for (int i = 0; i < D; ++i) fin[i] = 0.0;
#pragma unroll
    for (int k = 0; k < kIirT; ++k) {
      if (tail_here && k == valid) {
#pragma unroll
        for (int i = 0; i < D; ++i) fin[i] = z[i];
      }
      v[k] = lfilter_step_scan<M>(P, v[k], z);
    }
    {
      double2* row = reinterpret_cast<double2*>(buf + tid * kJointPitch);
#pragma unroll
      for (int k = 0; k < kIirT / 2; ++k) row[k] = make_double2(v[2 * k], v[2 * k + 1]);
    }
    __syncthreads();
    if (tail_here) {
#pragma unroll
      for (int i = 0; i < D; ++i) s_carry[i] = fin[i];
    }
    if (cnt == kIirTile && tid == kIirThreads - 1) {
#pragma unroll
      for (int i = 0; i < D; ++i) s_carry[i] = z[i];
    }
#pragma unroll
    for (int j = 0; j < kIirT / 2; ++j) {
      const int c = j * kIirThreads + tid;
      const double2 d = *reinterpret_cast<const double2*>(buf + (c >> 3) * kJointPitch + 2 * (c & 7));
      if (kAligned && 2 * c + 1 < cnt) {
        *reinterpret_cast<double2*>(ys + base + 2 * c) = d;
      } else {
        if (2 * c < cnt) ys[base + 2 * c] = d.x;
        if (2 * c + 1 < cnt) ys[base + 2 * c + 1] = d.y;
      }
    }
    __syncthreads();
  }
  if (kAligned) cp_async_wait<0>();
  if (zf && tid < D) zf[sig * D + tid] = s_carry[tid];
}

// host: 2x2 matrix powers in long double
struct M2 {
  long double a, b, c, d;
};
static M2 mm(const M2& x, const M2& y) {
  return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}
static M2 mpow(M2 base, long e) {
  M2 r{1, 0, 0, 1};
  while (e) {
    if (e & 1) r = mm(r, base);
    base = mm(base, base);
    e >>= 1;
  }
  return r;
}
static void put(double* dst, const M2& m) {
  dst[0] = (double)m.a; dst[1] = (double)m.b; dst[2] = (double)m.c; dst[3] = (double)m.d;
}

// D x D matrices in long double (joint state of the cascade)
struct MD {
  long double m[kMaxJoint][kMaxJoint];
};
static MD md_mul(const MD& x, const MD& y, int D) {
  MD r{};
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) {
      long double acc = 0;
      for (int k = 0; k < D; ++k) acc += x.m[i][k] * y.m[k][j];
      r.m[i][j] = acc;
    }
  return r;
}
static MD md_identity(int D) {
  MD r{};
  for (int i = 0; i < D; ++i) r.m[i][i] = 1;
  return r;
}
static void md_put(double* dst, const MD& a) {
  for (int i = 0; i < kMaxJoint; ++i)
    for (int j = 0; j < kMaxJoint; ++j) dst[i * kMaxJoint + j] = (double)a.m[i][j];
}
// (M^16)^l per lane, (M^16)^(2^d) per scan level, (M^16)^32 per warp
static void power_tables(const MD& M, int D, IirJointTables* tab) {
  MD MT = md_identity(D);
  for (int i = 0; i < kIirT; ++i) MT = md_mul(MT, M, D);
  MD pw = md_identity(D);
  for (int l = 0; l < 32; ++l) {
    md_put(tab->lane[l], pw);
    pw = md_mul(pw, MT, D);
  }
  md_put(tab->warp, pw);
  MD sq = MT;
  for (int d = 0; d < 5; ++d) {
    md_put(tab->lvl[d], sq);
    sq = md_mul(sq, sq, D);
  }
}
// homogeneous one-step matrix of the cascade: section k's input is y_{k-1} = z0_{k-1} + b0_{k-1} x_{k-1}
static void joint_tables(const IirParams& P, IirJointTables* tab) {
  const int S = P.n_sections, D = 2 * S;
  MD M{};
  for (int k = 0; k < S; ++k) {
    const long double a1 = P.sec[k].a1, a2 = P.sec[k].a2, b0 = P.sec[k].b0, b1 = P.sec[k].b1, b2 = P.sec[k].b2;
    const long double c1 = b1 - a1 * b0, c2 = b2 - a2 * b0;
    tab->c1[k] = (double)c1;
    tab->c2[k] = (double)c2;
    M.m[2 * k][2 * k] = -a1;
    M.m[2 * k][2 * k + 1] = 1;
    M.m[2 * k + 1][2 * k] = -a2;
    // the input of section k as a linear function of the earlier sections' states (x = 0): u_k = z0_{k-1} + b0_{k-1} u_{k-1}
    long double gain = 1;
    for (int j = k - 1; j >= 0; --j) {
      M.m[2 * k][2 * j] += c1 * gain;
      M.m[2 * k + 1][2 * j] += c2 * gain;
      gain *= P.sec[j].b0;
    }
  }
  power_tables(M, D, tab);
}
// one-step matrix of lfilter's DF2T state (order Mo): z_i' = -a_{i+1} z_0 + z_{i+1}
static void lfilter_tables(const LfScanParams& P, int Mo, IirJointTables* tab) {
  MD M{};
  for (int i = 0; i < Mo; ++i) {
    M.m[i][0] = -(long double)P.a[i + 1];
    if (i + 1 < Mo) M.m[i][i + 1] = 1;
  }
  power_tables(M, Mo, tab);
}

// Host -> device copy of a small table (the initial states) that does NOT stall the stream: cudaMemcpyAsync from
// PAGEABLE memory synchronises the stream before it stages the data, which serialises the host with every filter call
// (predistort in scan mode: 0.32 ms of kernel became 0.59 ms per call).  The bytes go through a small ring of pinned
// buffers per host thread instead; a slot is reused only after the copy that read it has completed (one event each).
static cudaError_t upload_small(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  constexpr int kSlots = 4;
  constexpr size_t kSlotBytes = 64 << 10;
  struct Ring {
    unsigned char* host[kSlots] = {};
    cudaEvent_t done[kSlots] = {};
    int next = 0;
  };
  static thread_local Ring rings[64];  // per host thread AND device: an event belongs to the device it was created on
  if (bytes == 0) return cudaSuccess;
  int dev = -1;
  if (bytes > kSlotBytes || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  Ring& ring = rings[dev];
  const int i = ring.next;
  cudaError_t e = cudaSuccess;
  if (!ring.host[i]) {
    e = cudaHostAlloc((void**)&ring.host[i], kSlotBytes, cudaHostAllocPortable);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ring.done[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      cudaGetLastError();
      ring.host[i] = nullptr;
      return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    }
  } else {
    e = cudaEventSynchronize(ring.done[i]);
    if (e != cudaSuccess) return e;
  }
  memcpy(ring.host[i], src, bytes);
  e = cudaMemcpyAsync(dst, ring.host[i], bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaEventRecord(ring.done[i], st);
  ring.next = (i + 1) % kSlots;
  return e;
}

}  // namespace wfm

extern "C" int wfm_sosfilt(const double* sos, int32_t n_sections, double initial, const double* x, double* y,
                           int64_t n_sig, int64_t n, int64_t stride, const double* zi, double* zf, int32_t mode,
                           void* stream) {
  using namespace wfm;
  if (!sos || n_sections < 1 || n_sections > kMaxSections || n_sig < 0 || n < 0 || (n_sig > 1 && stride < n))
    return WFM_EINVAL;
  if (mode != WFM_IIR_EXACT && mode != WFM_IIR_SCAN) return WFM_EINVAL;
  if (n_sig == 0) return WFM_OK;
  if ((!x || !y) && n > 0) return WFM_EINVAL;  // (an empty signal has no buffer to point at)
  cudaStream_t st = (cudaStream_t)stream;
  IirParams P{};
  P.n_sections = n_sections;
  P.initial = initial;
  for (int k = 0; k < n_sections; ++k) {
    const double* c = sos + 6 * k;
    const double a0 = c[3];
    // scipy normalises by a0 when it is not 1
    P.sec[k] = {c[0] / a0, c[1] / a0, c[2] / a0, c[4] / a0, c[5] / a0};
  }
  const size_t state_bytes = sizeof(double) * 2 * (size_t)n_sections * (size_t)n_sig;
  double *d_zi = nullptr, *d_zf = nullptr;
  cudaError_t e = cudaSuccess;
  // stream-ordered scratch for the states (pageable host sources are staged by the driver
  // before cudaMemcpyAsync returns, so no synchronisation is needed for them)
  if (zi) {
    e = cudaMallocAsync(&d_zi, state_bytes, st);
    if (e == cudaSuccess) e = upload_small(d_zi, zi, state_bytes, st);
  }
  if (e == cudaSuccess && zf) e = cudaMallocAsync(&d_zf, state_bytes, st);
  if (e == cudaSuccess && n > 0) {
    if (mode == WFM_IIR_EXACT) {
      const unsigned blocks = (unsigned)((n_sig + kExactWarps - 1) / kExactWarps);
      sosfilt_exact_kernel<<<blocks, 32 * kExactWarps, 0, st>>>(P, x, y, n_sig, n, stride, d_zi, d_zf);
      e = cudaGetLastError();
    } else if (!getenv("WFM_IIR_OLD_SCAN")) {
      // the joint kernel takes the sections two by two: a cascade of S sections is ceil(S / 2) launches, each a read and
      // a write of the signal; `initial` is subtracted by the first launch and added back by the last one only
      const bool aligned = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && (stride % 2 == 0 || n_sig == 1);
      const size_t smem = 2 * (size_t)kJointTileBytes;
      for (int k0 = 0; k0 < n_sections && e == cudaSuccess; k0 += 2) {
        IirParams Pk{};
        Pk.n_sections = std::min(2, n_sections - k0);
        Pk.initial = P.initial;
        Pk.shift_in = k0 == 0;
        Pk.shift_out = k0 + 2 >= n_sections;
        for (int k = 0; k < Pk.n_sections; ++k) Pk.sec[k] = P.sec[k0 + k];
        static thread_local IirJointTables jt;
        joint_tables(Pk, &jt);
        const double* xin = k0 == 0 ? x : y;
        auto launch = [&](auto kern) {
          cudaError_t ee = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (ee != cudaSuccess) return ee;
          kern<<<(unsigned)n_sig, kIirThreads, smem, st>>>(Pk, jt, xin, y, n, stride, d_zi ? d_zi + 2 * k0 : nullptr,
                                                           d_zf ? d_zf + 2 * k0 : nullptr, 2 * n_sections);
          return cudaGetLastError();
        };
        if (Pk.n_sections == 1) e = aligned ? launch(sosfilt_scan_joint_kernel<1, true>) : launch(sosfilt_scan_joint_kernel<1, false>);
        else e = aligned ? launch(sosfilt_scan_joint_kernel<2, true>) : launch(sosfilt_scan_joint_kernel<2, false>);
      }
    } else {
      static thread_local IirScanTables tab;
      for (int k = 0; k < n_sections; ++k) {
        const M2 A{-(long double)P.sec[k].a1, 1.0L, -(long double)P.sec[k].a2, 0.0L};
        const M2 AT = mpow(A, kIirT);
        M2 pw{1, 0, 0, 1};
        for (int l = 0; l < 32; ++l) {
          put(tab.lane[k][l], pw);
          pw = mm(pw, AT);
        }
        put(tab.warp[k], pw);  // AT^32
        M2 sq = AT;
        for (int d = 0; d < 5; ++d) {
          put(tab.lvl[k][d], sq);
          sq = mm(sq, sq);
        }
      }
      sosfilt_scan_kernel<<<(unsigned)n_sig, kIirThreads, 0, st>>>(P, tab, x, y, n, stride, d_zi, d_zf);
      e = cudaGetLastError();
    }
  } else if (e == cudaSuccess && zf && zi) {
    e = cudaMemcpyAsync(d_zf, d_zi, state_bytes, cudaMemcpyDeviceToDevice, st);
  } else if (e == cudaSuccess && zf) {
    e = cudaMemsetAsync(d_zf, 0, state_bytes, st);
  }
  if (e == cudaSuccess && zf) {
    e = cudaMemcpyAsync(zf, d_zf, state_bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the caller reads zf on return
  }
  if (d_zi) cudaFreeAsync(d_zi, st);
  if (d_zf) cudaFreeAsync(d_zf, st);
  return e == cudaSuccess ? WFM_OK : WFM_ECUDA;
}

extern "C" int wfm_lfilter(const double* b, int32_t nb, const double* a, int32_t na, const double* x, double* y,
                           int64_t n_sig, int64_t n, int64_t stride, const double* zi, double* zf, void* stream) {
  return wfm_lfilter_mode(b, nb, a, na, x, y, n_sig, n, stride, zi, zf, WFM_IIR_EXACT, stream);
}

extern "C" int wfm_lfilter_mode(const double* b, int32_t nb, const double* a, int32_t na, const double* x, double* y,
                                int64_t n_sig, int64_t n, int64_t stride, const double* zi, double* zf, int32_t mode,
                                void* stream) {
  using namespace wfm;
  if (mode != WFM_IIR_EXACT && mode != WFM_IIR_SCAN) return WFM_EINVAL;
  if (!b || !a || nb < 1 || na < 1 || a[0] == 0.0 || n_sig < 0 || n < 0 || (n_sig > 1 && stride < n)) return WFM_EINVAL;
  const int M = std::max(nb, na) - 1;
  if (M > kMaxOrder) return WFM_EUNSUPPORTED;
  if (n_sig == 0) return WFM_OK;
  if ((!x || !y) && n > 0) return WFM_EINVAL;  // (an empty signal has no buffer to point at)
  cudaStream_t st = (cudaStream_t)stream;
  LfilterParams P{};
  P.order = M;
  for (int k = 0; k <= M; ++k) {  // scipy normalises both polynomials by a[0]
    P.b[k] = k < nb ? b[k] / a[0] : 0.0;
    P.a[k] = k < na ? a[k] / a[0] : 0.0;
  }
  const size_t state_bytes = sizeof(double) * (size_t)std::max(M, 1) * (size_t)n_sig;
  double *d_zi = nullptr, *d_zf = nullptr;
  cudaError_t e = cudaSuccess;
  if (zi && M > 0) {
    e = cudaMallocAsync(&d_zi, state_bytes, st);
    if (e == cudaSuccess) e = upload_small(d_zi, zi, sizeof(double) * M * n_sig, st);
  }
  if (e == cudaSuccess && zf && M > 0) e = cudaMallocAsync(&d_zf, state_bytes, st);
  if (e == cudaSuccess && mode == WFM_IIR_SCAN && M >= 1 && M <= kMaxJoint && n > 0) {
    LfScanParams S{};
    for (int k = 0; k <= M; ++k) {
      S.b[k] = P.b[k];
      S.a[k] = P.a[k];
      S.c[k] = (double)((long double)P.b[k] - (long double)P.a[k] * (long double)P.b[0]);
    }
    static thread_local IirJointTables jt;
    lfilter_tables(S, M, &jt);
    const bool aligned = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && (stride % 2 == 0 || n_sig == 1);
    const size_t smem = 2 * (size_t)kJointTileBytes;
    auto launch = [&](auto kern) {
      cudaError_t ee = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (ee != cudaSuccess) return ee;
      kern<<<(unsigned)n_sig, kIirThreads, smem, st>>>(S, jt, x, y, n, stride, d_zi, d_zf);
      return cudaGetLastError();
    };
    switch (M) {
      case 1: e = aligned ? launch(lfilter_scan_kernel<1, true>) : launch(lfilter_scan_kernel<1, false>); break;
      case 2: e = aligned ? launch(lfilter_scan_kernel<2, true>) : launch(lfilter_scan_kernel<2, false>); break;
      case 3: e = aligned ? launch(lfilter_scan_kernel<3, true>) : launch(lfilter_scan_kernel<3, false>); break;
      default: e = aligned ? launch(lfilter_scan_kernel<4, true>) : launch(lfilter_scan_kernel<4, false>); break;
    }
  } else if (e == cudaSuccess) {
    // (orders above 4, order 0 and empty signals: the sequential kernel, whatever the mode)
    lfilter_exact_kernel<<<(unsigned)((n_sig + kExactWarps - 1) / kExactWarps), 32 * kExactWarps, 0, st>>>(
        P, x, y, n_sig, n, stride, d_zi, d_zf);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess && d_zf) {
    e = cudaMemcpyAsync(zf, d_zf, sizeof(double) * M * n_sig, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the caller reads zf on return
  }
  if (d_zi) cudaFreeAsync(d_zi, st);
  if (d_zf) cudaFreeAsync(d_zf, st);
  return e == cudaSuccess ? WFM_OK : WFM_ECUDA;
}

// wfm_sample.cu — K1: batched piecewise-waveform sampling kernel for sm_100a.
//
// Replaces, for a whole batch of channels in one launch, the reference's
//   np.arange -> calc_parts (np.searchsorted, per-segment _calc/_apply,
//   np.clip) -> zeros_like -> _fill_parts
// (/root/reference/waveforms/waveform.py:173-207, :529-563, :679-693;
//  /root/reference/waveforms/_waveform.pyx:130-169).
//
// Work decomposition: the output of the batch is cut into tiles of
// kTileSamples consecutive samples of ONE channel; one CTA per tile.
//   1. warps 0/1 locate the segments of the tile's first / last abscissa with
//      a 32-ary ballot search over the channel's bounds (global, L2-resident);
//   2. the CTA stages that slice of the segment table (bounds + factor/term
//      pointers) in shared memory;
//   3. every thread owns 16 bytes of output per row (2 fp64 / 4 fp32 samples),
//      finds its segment by binary search in shared memory, interprets the
//      segment's factor list / term list, and issues one 16-byte streaming
//      store.  A warp-row therefore writes 512 contiguous bytes.
// Tiles that lie inside a single segment skip the search; tiles inside a zero
// segment degenerate to pure stores (the HBM-write-bound case).
//
// Bounds are HBM-resident f64, the output is write-once: stores use
// st.global.cs (evict-first) so they do not displace the IR in L2.
#include <cuda_runtime.h>
#include <stdint.h>
#include "wfm_basis.cuh"
#include "wfm_multidrag.cuh"
#include "wfm_internal.h"

namespace wfm {

constexpr int kThreads = 256;
constexpr int kStageSegs = 1024;  // segment-table rows staged per tile (8 KB bounds + 8 KB ptrs)
constexpr int kMaxSlots = 12;     // distinct factor values cached per segment evaluation

template <typename T> struct OutVec;
template <> struct OutVec<double> { static constexpr int N = 2; };
template <> struct OutVec<float> { static constexpr int N = 4; };

__device__ __forceinline__ void store_vec(double* p, const double (&v)[2]) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v[0]), "d"(v[1]) : "memory");
}
__device__ __forceinline__ void store_vec(float* p, const double (&v)[4]) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"((float)v[0]), "f"((float)v[1]),
               "f"((float)v[2]), "f"((float)v[3])
               : "memory");
}
__device__ __forceinline__ void load_vec(const double* p, double (&v)[2]) {
  double2 t = *reinterpret_cast<const double2*>(p);
  v[0] = t.x; v[1] = t.y;
}
__device__ __forceinline__ void load_vec(const float* p, double (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

// abscissa of sample j of channel w: x[j] = t0 + j*delta — a multiply and an add,
// never fused (np.arange / np.linspace fill loops), minus the stack pre-shift.
__device__ __forceinline__ double abscissa(const WfmWave& w, const double* __restrict__ xs, int64_t j) {
  double x;
  if (w.flags & WFM_WAVE_EXPLICIT_X) {
    x = xs[w.x_off + j];
  } else {
    x = add(w.t0, mul((double)j, w.delta));
    if ((w.flags & WFM_WAVE_LAST_OVERRIDE) && j == w.n - 1) x = w.x_last;
  }
  if (w.flags & WFM_WAVE_PRESHIFT) x = sub(x, w.pre_shift);
  return x;
}

// Number of bounds <= x among b[0..n) (b sorted, b[n-1] = +inf): the index of
// the segment that owns x, i.e. np.searchsorted(bounds, x, side='right').
// Executed by one full warp; 32 pivots per round.
__device__ int warp_segment_search(const double* __restrict__ b, int n, double x, int lane) {
  int lo = 0, hi = n - 1;  // answer in [lo, hi]
  while (hi > lo) {
    int width = hi - lo;
    int stride = (width + 31) / 32;
    int idx = min(lo + (lane + 1) * stride - 1, hi);
    bool le = (idx < hi) ? (__ldg(b + idx) <= x) : false;  // b[hi] > x is known
    unsigned m = __ballot_sync(0xffffffffu, le);
    int c = __popc(m);  // bounds are sorted, so `le` is a prefix
    int new_lo = (c == 0) ? lo : min(lo + c * stride - 1, hi) + 1;
    int new_hi = (c == 32) ? hi : min(lo + (c + 1) * stride - 1, hi);
    lo = min(new_lo, hi);
    hi = max(new_hi, lo);
  }
  return lo;
}

struct Acc {
  double re, im;
};

// Evaluate one segment's program at abscissa x.
template <bool kComplex>
__device__ __forceinline__ Acc eval_segment(const DevProgram& P, const WfmWave& w, WfmSegPtr p0, WfmSegPtr p1,
                                            double x) {
  Acc total{w.offset, 0.0};
  const int nt = p1.term - p0.term;
  if (nt == 0) return total;  // zero segment: untouched by clip (calc_parts skips it)
  const int nf = p1.fac - p0.fac;
  double vals[kMaxSlots];
  const WfmFactor* __restrict__ facs = P.facs + p0.fac;
#pragma unroll 1
  for (int k = 0; k < nf && k < kMaxSlots; ++k) vals[k] = eval_factor(facs[k], x, P.args);

  double g_re = 0.0, g_im = 0.0;
  bool g_first = true;
#pragma unroll 1
  for (int it = 0; it < nt; ++it) {
    const WfmTerm tm = P.terms[p0.term + it];
    double prod = 1.0;
    bool p_first = true;
#pragma unroll 1
    for (int r = 0; r < tm.n_ref; ++r) {
      const WfmRef ref = P.refs[tm.ref_begin + r];
      double v = (ref.slot < kMaxSlots) ? vals[ref.slot] : eval_factor(facs[ref.slot], x, P.args);
      if (ref.kind == WFM_POW_INT) v = pow_small_int(v, (int)ref.expo);
      else if (ref.kind == WFM_POW_GEN) v = pow(v, ref.expo);
      prod = p_first ? v : mul(prod, v);  // 1 * v == v
      p_first = false;
    }
    const double t_re = mul(tm.amp_re, prod);
    g_re = g_first ? t_re : add(g_re, t_re);  // 0 + a == a
    if (kComplex) {
      const double t_im = mul(tm.amp_im, prod);
      g_im = g_first ? t_im : add(g_im, t_im);
    }
    g_first = false;
    if (tm.flags & WFM_TERM_GROUP_END) {
      total.re = add(total.re, g_re);
      if (kComplex) total.im = add(total.im, g_im);
      g_first = true;
    }
  }
  if (w.flags & WFM_WAVE_CLIP) total.re = fmin(fmax(total.re, w.clip_lo), w.clip_hi);
  return total;
}

template <typename OutT, bool kAccumulate>
__global__ void __launch_bounds__(kThreads) sample_kernel(DevProgram P, const TileDesc* __restrict__ tiles,
                                                          OutT* __restrict__ out) {
  constexpr int V = OutVec<OutT>::N;
  constexpr int kRowSamples = kThreads * V;
  __shared__ double s_bound[kStageSegs];
  __shared__ WfmSegPtr s_ptr[kStageSegs + 1];
  __shared__ int s_range[2];

  const TileDesc td = tiles[blockIdx.x];
  const WfmWave w = P.waves[td.wave];
  const int64_t j0 = td.j0;
  const int cnt = (int)min((int64_t)kTileSamples, w.n - j0);
  const double* __restrict__ gb = P.seg_bound + w.seg_begin;
  const WfmSegPtr* __restrict__ gp = P.seg_ptr + w.seg_begin;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp < 2) {
    const int64_t j = warp == 0 ? j0 : j0 + cnt - 1;
    const int s = warp_segment_search(gb, w.n_seg, abscissa(w, P.x, j), lane);
    if (lane == 0) s_range[warp] = s;
  }
  __syncthreads();
  const int seg_lo = s_range[0];
  const int nb = s_range[1] - seg_lo + 1;  // abscissae are non-decreasing => >= 1
  const bool staged = nb <= kStageSegs;
  if (staged) {
    for (int k = threadIdx.x; k < nb; k += kThreads) s_bound[k] = gb[seg_lo + k];
    for (int k = threadIdx.x; k <= nb; k += kThreads) s_ptr[k] = gp[seg_lo + k];
  }
  __syncthreads();

  OutT* __restrict__ dst = out + w.out_off + j0;
  const bool tile_zero = nb == 1 && s_ptr[0].term == s_ptr[1].term && w.offset == 0.0;

  for (int base = threadIdx.x * V; base < cnt; base += kRowSamples) {
    double v[V];
    if (tile_zero) {
#pragma unroll
      for (int e = 0; e < V; ++e) v[e] = 0.0;
    } else {
      int seg = 0;  // relative to seg_lo
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int jj = base + e;
        if (jj >= cnt) { v[e] = 0.0; continue; }
        const double x = abscissa(w, P.x, j0 + jj);
        WfmSegPtr p0, p1;
        if (staged) {
          if (e == 0) {
            int lo = 0, hi = nb - 1;  // first k with s_bound[k] > x
            while (lo < hi) {
              int mid = (lo + hi) >> 1;
              if (s_bound[mid] <= x) lo = mid + 1; else hi = mid;
            }
            seg = lo;
          } else {
            while (seg < nb - 1 && s_bound[seg] <= x) ++seg;
          }
          p0 = s_ptr[seg];
          p1 = s_ptr[seg + 1];
        } else {
          int lo = (e == 0) ? 0 : seg, hi = nb - 1;
          while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (__ldg(gb + seg_lo + mid) <= x) lo = mid + 1; else hi = mid;
          }
          seg = lo;
          p0 = gp[seg_lo + seg];
          p1 = gp[seg_lo + seg + 1];
        }
        v[e] = eval_segment<false>(P, w, p0, p1, x).re;
      }
    }
    if (base + V <= cnt) {
      if (kAccumulate) {
        double old[V];
        load_vec(dst + base, old);
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = add(old[e], v[e]);
      }
      store_vec(dst + base, v);
    } else {
      for (int e = 0; e < V && base + e < cnt; ++e)
        dst[base + e] = kAccumulate ? (OutT)add((double)dst[base + e], v[e]) : (OutT)v[e];
    }
  }
}

// complex128 output: interleaved (re, im); one sample per thread per row.
template <bool kAccumulate>
__global__ void __launch_bounds__(kThreads) sample_kernel_c128(DevProgram P, const TileDesc* __restrict__ tiles,
                                                               double2* __restrict__ out) {
  const TileDesc td = tiles[blockIdx.x];
  const WfmWave w = P.waves[td.wave];
  const int64_t j0 = td.j0;
  const int cnt = (int)min((int64_t)kTileSamples, w.n - j0);
  const double* __restrict__ gb = P.seg_bound + w.seg_begin;
  const WfmSegPtr* __restrict__ gp = P.seg_ptr + w.seg_begin;
  double2* __restrict__ dst = out + w.out_off + j0;
  int seg = 0;
  for (int jj = threadIdx.x; jj < cnt; jj += kThreads) {
    const double x = abscissa(w, P.x, j0 + jj);
    int lo = seg, hi = w.n_seg - 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (__ldg(gb + mid) <= x) lo = mid + 1; else hi = mid;
    }
    seg = lo;
    Acc a = eval_segment<true>(P, w, gp[seg], gp[seg + 1], x);
    if (kAccumulate) {
      double2 o = dst[jj];
      a.re = add(o.x, a.re);
      a.im = add(o.y, a.im);
    }
    dst[jj] = make_double2(a.re, a.im);
  }
}

cudaError_t launch_sample(const DevProgram& P, const TileDesc* tiles, int64_t n_tiles, int dtype, int accumulate,
                          void* out, cudaStream_t stream) {
  if (n_tiles == 0) return cudaSuccess;
  dim3 grid((unsigned)n_tiles), block(kThreads);
  if (dtype == WFM_F64) {
    if (accumulate) sample_kernel<double, true><<<grid, block, 0, stream>>>(P, tiles, (double*)out);
    else sample_kernel<double, false><<<grid, block, 0, stream>>>(P, tiles, (double*)out);
  } else if (dtype == WFM_F32) {
    if (accumulate) sample_kernel<float, true><<<grid, block, 0, stream>>>(P, tiles, (float*)out);
    else sample_kernel<float, false><<<grid, block, 0, stream>>>(P, tiles, (float*)out);
  } else {
    if (accumulate) sample_kernel_c128<true><<<grid, block, 0, stream>>>(P, tiles, (double2*)out);
    else sample_kernel_c128<false><<<grid, block, 0, stream>>>(P, tiles, (double2*)out);
  }
  return cudaGetLastError();
}

}  // namespace wfm

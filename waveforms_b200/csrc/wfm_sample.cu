// wfm_sample.cu — K1: batched piecewise-waveform sampling kernel for sm_100a.
//
// Replaces, for a whole batch of channels in one launch, the reference's
//   np.arange -> calc_parts (np.searchsorted, per-segment _calc/_apply,
//   np.clip) -> zeros_like -> _fill_parts
// (/root/reference/waveforms/waveform.py:173-207, :529-563, :679-693;
//  /root/reference/waveforms/_waveform.pyx:130-169).
//
// Work decomposition: the output of the batch is cut into tiles of kTileSamples
// consecutive samples of ONE channel; one CTA (256 threads) per tile.
//
//  0. (once per program, prepare_tiles_kernel) every tile learns the segments of
//     its first and last abscissa: one thread per tile, binary search over the
//     channel's bounds.
//  1. The CTA turns the tile's slice of the bound table into INTEGER sample
//     positions: thread k finds the first sample of the tile whose abscissa is
//     >= bound k — a division for the guess, then exact comparisons against the
//     rounded grid value x[j] = t0 + j*delta, so ownership is bit-identical to
//     np.searchsorted on the reference's grid.  Positions and the segments'
//     factor/term pointers live in shared memory.
//  2. A warp owns chunks of 32*V consecutive samples (V = 2 fp64 / 4 fp32 per
//     thread = one 16-byte store per thread, 512 contiguous bytes per warp).
//     The chunk's first segment comes from a per-chunk table (one LDS).  A chunk
//     that lies inside one segment is warp-uniform: an empty (zero) segment
//     costs a store and nothing else — no abscissa, no search; a non-empty one
//     runs the segment program without divergence.  Mixed chunks let every lane
//     advance from the chunk's first segment.
//  3. The segment program (distinct factors, then terms referencing factor
//     slots) is interpreted for the thread's V samples at once.
//
// Output is write-once: stores use st.global.cs (evict-first) so they do not
// displace the IR in L2.  Algorithmic traffic: 8 B (4 B) per sample, write-only.
#include <cuda_runtime.h>
#include <stdint.h>
#include "wfm_basis.cuh"
#include "wfm_multidrag.cuh"
#include "wfm_internal.h"

namespace wfm {

constexpr int kThreads = 256;
constexpr int kStageSegs = 1024;  // segment rows staged per tile
constexpr int kMaxSlots = 12;     // distinct factor values cached per segment evaluation
constexpr int kMaxChunks = kTileSamples / 64;

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

// number of bounds <= x among b[0..n) (b sorted, b[n-1] = +inf): the segment that
// owns x == np.searchsorted(bounds, x, side='right')
__device__ __forceinline__ int owning_segment(const double* __restrict__ b, int n, double x) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(b + mid) <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- pre-pass: segment range of every tile (one thread per tile) --------------------
__global__ void prepare_tiles_kernel(DevProgram P, TileDesc* __restrict__ tiles, int64_t n_tiles) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  TileDesc td = tiles[t];
  const WfmWave w = P.waves[td.wave];
  const int64_t last = min(td.j0 + (int64_t)kTileSamples, w.n) - 1;
  const double* b = P.seg_bound + w.seg_begin;
  td.seg_lo = owning_segment(b, w.n_seg, abscissa(w, P.x, td.j0));
  td.seg_hi = max(td.seg_lo, owning_segment(b, w.n_seg, abscissa(w, P.x, last)));
  tiles[t] = td;
}

// first sample jj in [0, cnt] of the tile with abscissa >= bound (cnt if none)
__device__ int first_sample_at_or_after(const WfmWave& w, const double* __restrict__ xs, int64_t j0, int cnt,
                                        double bound) {
  if (!(w.flags & WFM_WAVE_EXPLICIT_X) && w.delta > 0.0) {
    double b = bound;
    if (w.flags & WFM_WAVE_PRESHIFT) b = b + w.pre_shift;
    double g = ceil((b - w.t0) / w.delta) - (double)j0;
    int jj = g <= 0.0 ? 0 : (g >= (double)cnt ? cnt : (int)g);
    // exact fix-up against the rounded grid (monotone in j)
    while (jj > 0 && abscissa(w, xs, j0 + jj - 1) >= bound) --jj;
    while (jj < cnt && abscissa(w, xs, j0 + jj) < bound) ++jj;
    return jj;
  }
  int lo = 0, hi = cnt;  // generic: binary search over the tile's abscissae
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (abscissa(w, xs, j0 + mid) < bound) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <int V>
struct Acc {
  double re[V];
  double im[V];
};

// basis function on V samples at once; the hot ids are unrolled over V so the
// independent evaluations interleave, the rest fall back to the scalar code
template <int V>
__device__ __forceinline__ void eval_factor_v(const WfmFactor& f, const double (&x)[V], const double* __restrict__ args,
                                              double (&out)[V]) {
  switch (f.func) {
    case WFM_COS:
#pragma unroll
      for (int e = 0; e < V; ++e) out[e] = cos(mul(f.a0, sub(x[e], f.shift)));
      break;
    case WFM_GAUSSIAN:
#pragma unroll
      for (int e = 0; e < V; ++e) out[e] = f_gaussian(sub(x[e], f.shift), f.a0);
      break;
    case WFM_ERF:
#pragma unroll
      for (int e = 0; e < V; ++e) out[e] = erf(dvd(sub(x[e], f.shift), f.a0));
      break;
    case WFM_LINEAR:
#pragma unroll
      for (int e = 0; e < V; ++e) out[e] = sub(x[e], f.shift);
      break;
    default:
#pragma unroll 1
      for (int e = 0; e < V; ++e) out[e] = eval_factor(f, x[e], args);
      break;
  }
}

// Evaluate one segment's program at V abscissae.
template <int V, bool kComplex>
__device__ __forceinline__ void eval_segment(const DevProgram& P, const WfmWave& w, WfmSegPtr p0, WfmSegPtr p1,
                                             const double (&x)[V], Acc<V>& total) {
#pragma unroll
  for (int e = 0; e < V; ++e) { total.re[e] = w.offset; total.im[e] = 0.0; }
  const int nt = p1.term - p0.term;
  if (nt == 0) return;  // zero segment: untouched by clip (calc_parts skips it)
  const int nf = p1.fac - p0.fac;
  double vals[kMaxSlots][V];
  const WfmFactor* __restrict__ facs = P.facs + p0.fac;
#pragma unroll 1
  for (int k = 0; k < nf && k < kMaxSlots; ++k) eval_factor_v<V>(facs[k], x, P.args, vals[k]);

  double g_re[V], g_im[V];
  bool g_first = true;
#pragma unroll 1
  for (int it = 0; it < nt; ++it) {
    const WfmTerm tm = P.terms[p0.term + it];
    double prod[V];
#pragma unroll
    for (int e = 0; e < V; ++e) prod[e] = 1.0;
    bool p_first = true;
#pragma unroll 1
    for (int r = 0; r < tm.n_ref; ++r) {
      const WfmRef ref = P.refs[tm.ref_begin + r];
      double v[V];
      if (ref.slot < kMaxSlots) {
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = vals[ref.slot][e];
      } else {
        eval_factor_v<V>(facs[ref.slot], x, P.args, v);
      }
      if (ref.kind == WFM_POW_INT) {
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = pow_small_int(v[e], (int)ref.expo);
      } else if (ref.kind == WFM_POW_GEN) {
#pragma unroll 1
        for (int e = 0; e < V; ++e) v[e] = pow(v[e], ref.expo);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) prod[e] = p_first ? v[e] : mul(prod[e], v[e]);  // 1 * v == v
      p_first = false;
    }
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const double t_re = mul(tm.amp_re, prod[e]);
      g_re[e] = g_first ? t_re : add(g_re[e], t_re);  // 0 + a == a
      if (kComplex) {
        const double t_im = mul(tm.amp_im, prod[e]);
        g_im[e] = g_first ? t_im : add(g_im[e], t_im);
      }
    }
    g_first = false;
    if (tm.flags & WFM_TERM_GROUP_END) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        total.re[e] = add(total.re[e], g_re[e]);
        if (kComplex) total.im[e] = add(total.im[e], g_im[e]);
      }
      g_first = true;
    }
  }
  if (w.flags & WFM_WAVE_CLIP) {
#pragma unroll
    for (int e = 0; e < V; ++e) total.re[e] = fmin(fmax(total.re[e], w.clip_lo), w.clip_hi);
  }
}

template <typename OutT, bool kAccumulate>
__global__ void __launch_bounds__(kThreads) sample_kernel(DevProgram P, const TileDesc* __restrict__ tiles,
                                                          OutT* __restrict__ out) {
  constexpr int V = OutVec<OutT>::N;
  constexpr int kChunk = 32 * V;                      // samples per warp-chunk
  constexpr int kChunks = kTileSamples / kChunk;      // chunks per tile
  __shared__ int s_start[kStageSegs + 1];             // first tile-sample of staged segment k
  __shared__ WfmSegPtr s_ptr[kStageSegs + 1];
  __shared__ int s_chunk_seg[kMaxChunks];

  const TileDesc td = tiles[blockIdx.x];
  const WfmWave w = P.waves[td.wave];
  const int64_t j0 = td.j0;
  const int cnt = (int)min((int64_t)kTileSamples, w.n - j0);
  const double* __restrict__ gb = P.seg_bound + w.seg_begin;
  const WfmSegPtr* __restrict__ gp = P.seg_ptr + w.seg_begin;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg_lo = td.seg_lo;
  const int nb = td.seg_hi - seg_lo + 1;
  OutT* __restrict__ dst = out + w.out_off + j0;

  if (nb > kStageSegs) {
    // pathological density (> 1024 segments in one tile): per-sample search in global memory
    for (int base = threadIdx.x * V; base < cnt; base += kThreads * V) {
      double v[V];
      int seg = seg_lo;
#pragma unroll 1
      for (int e = 0; e < V; ++e) {
        v[e] = 0.0;
        if (base + e >= cnt) continue;
        double x1[1] = {abscissa(w, P.x, j0 + base + e)};
        while (seg < td.seg_hi && __ldg(gb + seg) <= x1[0]) ++seg;
        Acc<1> a;
        eval_segment<1, false>(P, w, gp[seg], gp[seg + 1], x1, a);
        v[e] = a.re[0];
      }
      for (int e = 0; e < V && base + e < cnt; ++e)
        dst[base + e] = kAccumulate ? (OutT)add((double)dst[base + e], v[e]) : (OutT)v[e];
    }
    return;
  }

  // ---- stage the tile's segment slice: sample positions + program pointers ----------
  for (int k = threadIdx.x; k <= nb; k += kThreads) {
    s_ptr[k] = gp[seg_lo + k];
    int pos;
    if (k == 0) pos = 0;
    else if (k == nb) pos = cnt;
    else pos = first_sample_at_or_after(w, P.x, j0, cnt, gb[seg_lo + k - 1]);
    s_start[k] = pos;
  }
  __syncthreads();
  if (threadIdx.x < kChunks) {
    // staged segment that owns the first sample of chunk c: last k with s_start[k] <= c*kChunk
    const int jj = threadIdx.x * kChunk;
    int lo = 0, hi = nb - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_start[mid] <= jj) lo = mid; else hi = mid - 1;
    }
    s_chunk_seg[threadIdx.x] = lo;
  }
  __syncthreads();

  for (int c = warp; c * kChunk < cnt; c += kThreads / 32) {
    const int cbeg = c * kChunk;
    const int cend = min(cbeg + kChunk, cnt);
    const int base = cbeg + lane * V;
    const int k0 = s_chunk_seg[c];
    const bool uniform = s_start[k0 + 1] >= cend;  // whole chunk inside staged segment k0
    double v[V];
    bool done = false;
    if (uniform) {
      const WfmSegPtr p0 = s_ptr[k0], p1 = s_ptr[k0 + 1];
      if (p0.term == p1.term) {
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = w.offset;  // empty segment: no abscissa needed
        done = true;
      } else if (base + V <= cend) {
        double x[V];
#pragma unroll
        for (int e = 0; e < V; ++e) x[e] = abscissa(w, P.x, j0 + base + e);
        Acc<V> a;
        eval_segment<V, false>(P, w, p0, p1, x, a);
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = a.re[e];
        done = true;
      }
    }
    if (!done) {
      int k = k0;
      // lanes advance from the chunk's first segment to their own
      while (k < nb - 1 && s_start[k + 1] <= base) ++k;
      if (base + V <= cend && s_start[k + 1] >= base + V) {
        const WfmSegPtr p0 = s_ptr[k], p1 = s_ptr[k + 1];
        if (p0.term == p1.term) {
#pragma unroll
          for (int e = 0; e < V; ++e) v[e] = w.offset;
        } else {
          double x[V];
#pragma unroll
          for (int e = 0; e < V; ++e) x[e] = abscissa(w, P.x, j0 + base + e);
          Acc<V> a;
          eval_segment<V, false>(P, w, p0, p1, x, a);
#pragma unroll
          for (int e = 0; e < V; ++e) v[e] = a.re[e];
        }
      } else {
        // a bound falls between this thread's samples (or the tile ends): one by one
#pragma unroll 1
        for (int e = 0; e < V; ++e) {
          v[e] = 0.0;
          if (base + e >= cend) continue;
          while (k < nb - 1 && s_start[k + 1] <= base + e) ++k;
          const WfmSegPtr p0 = s_ptr[k], p1 = s_ptr[k + 1];
          if (p0.term == p1.term) { v[e] = w.offset; continue; }
          double x1[1] = {abscissa(w, P.x, j0 + base + e)};
          Acc<1> a;
          eval_segment<1, false>(P, w, p0, p1, x1, a);
          v[e] = a.re[0];
        }
      }
    }
    if (base + V <= cend) {
      if (kAccumulate) {
        double old[V];
        load_vec(dst + base, old);
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = add(old[e], v[e]);
      }
      store_vec(dst + base, v);
    } else {
      for (int e = 0; e < V && base + e < cend; ++e)
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
  int seg = td.seg_lo;
  for (int jj = threadIdx.x; jj < cnt; jj += kThreads) {
    double x1[1] = {abscissa(w, P.x, j0 + jj)};
    int lo = seg, hi = td.seg_hi;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (__ldg(gb + mid) <= x1[0]) lo = mid + 1; else hi = mid;
    }
    seg = lo;
    Acc<1> a;
    eval_segment<1, true>(P, w, gp[seg], gp[seg + 1], x1, a);
    double re = a.re[0], im = a.im[0];
    if (kAccumulate) {
      double2 o = dst[jj];
      re = add(o.x, re);
      im = add(o.y, im);
    }
    dst[jj] = make_double2(re, im);
  }
}

cudaError_t launch_prepare_tiles(const DevProgram& P, TileDesc* tiles, int64_t n_tiles, cudaStream_t stream) {
  if (n_tiles == 0) return cudaSuccess;
  const int threads = 128;
  prepare_tiles_kernel<<<(unsigned)((n_tiles + threads - 1) / threads), threads, 0, stream>>>(P, tiles, n_tiles);
  return cudaGetLastError();
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

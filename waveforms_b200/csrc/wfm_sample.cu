// wfm_sample.cu — K1: batched piecewise-waveform sampling kernel for sm_100a.
//
// Replaces, for a whole batch of channels in one launch, the reference's
//   np.arange -> calc_parts (np.searchsorted, per-segment _calc/_apply,
//   np.clip) -> zeros_like -> _fill_parts
// (/root/reference/waveforms/waveform.py:173-207, :529-563, :679-693;
//  /root/reference/waveforms/_waveform.pyx:130-169).
//
// Work decomposition: the output of the batch is cut into tiles of tile_samples
// (2048..16384, chosen per program from its segment density) consecutive samples
// of ONE channel; one CTA (256 threads) per tile.
//
//  0. (once per program, prepare_tiles_kernel) every tile learns the segments of
//     its first and last abscissa: one thread per tile, binary search over the
//     channel's bounds.
//  1. Prologue.  Thread k turns bound k of the tile's slice of the segment table
//     into an INTEGER sample position: the first sample whose abscissa is >= the
//     bound — a division for the guess, then exact comparisons against the
//     rounded grid value x[j] = t0 + j*delta, so ownership is bit-identical to
//     np.searchsorted on the reference's grid.  It also classifies the segment:
//     FLAT (no basis factor: zero, or a constant evaluated once here) or ACTIVE.
//     Meanwhile one thread starts a TMA bulk copy (cp.async.bulk + mbarrier) of
//     the tile's slice of the factor / term / reference tables into shared
//     memory.
//  2. Phase 1 — stores.  A warp owns chunks of 32*V consecutive samples (V = 2
//     fp64 / 4 fp32 per thread = one 16-byte store per thread, 512 contiguous
//     bytes per warp).  Every sample of a FLAT segment is written here; no
//     abscissa is computed.  This is the HBM-write-bound part.
//  3. Phase 2 — compute.  The ACTIVE samples of the tile are enumerated through
//     a prefix sum over the segments and dealt round-robin to all 256 threads, so
//     a tile with one 40-sample pulse keeps 40 lanes of 2 warps busy once instead
//     of serialising inside one warp, and a dense tile keeps every lane busy.
//     Each thread interprets its sample's segment program (distinct factors,
//     then terms referencing factor slots) out of shared memory.
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
constexpr int kMaxChunks = kMaxTileSamples / 64;
constexpr int kIrBytes = 24576;   // shared-memory budget for the tile's factor/term/ref slice

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
__device__ __forceinline__ void store_one(double* p, double v) {
  asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void store_one(float* p, double v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"((float)v) : "memory");
}
__device__ __forceinline__ void load_vec(const double* p, double (&v)[2]) {
  double2 t = *reinterpret_cast<const double2*>(p);
  v[0] = t.x; v[1] = t.y;
}
__device__ __forceinline__ void load_vec(const float* p, double (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

// ---- mbarrier + TMA bulk copy (global -> shared) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// size and both addresses must be multiples of 16 bytes
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
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
  const int64_t last = min(td.j0 + (int64_t)P.tile_samples, w.n) - 1;
  const double* b = P.seg_bound + w.seg_begin;
  td.seg_lo = owning_segment(b, w.n_seg, abscissa(w, P.x, td.j0));
  td.seg_hi = max(td.seg_lo, owning_segment(b, w.n_seg, abscissa(w, P.x, last)));
  const WfmSegPtr a = P.seg_ptr[w.seg_begin + td.seg_lo], e = P.seg_ptr[w.seg_begin + td.seg_hi + 1];
  td.fac0 = a.fac;
  td.n_fac = e.fac - a.fac;
  td.term0 = a.term;
  td.n_term = e.term - a.term;
  td.ref0 = 0;
  td.n_ref = 0;
  if (td.n_term > 0) {
    td.ref0 = P.terms[a.term].ref_begin;
    const WfmTerm lt = P.terms[e.term - 1];
    td.n_ref = lt.ref_begin + lt.n_ref - td.ref0;
  }
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

// where the interpreter reads the program from: global tables, or the tile's
// slice staged in shared memory (pointers pre-biased so global indices work)
struct IrView {
  const WfmFactor* facs;
  const CTerm* cterms;  // compact terms (real-valued kernels)
  const WfmTerm* terms; // ABI terms / refs: global only (extended terms, complex kernel)
  const WfmRef* refs;
  const double* args;
};

// distinct factor values of one segment -> vals[0 .. min(nf, kMaxSlots))
__device__ __forceinline__ void eval_factors(const WfmFactor* facs, int nf, double x, const double* __restrict__ args,
                                             double (&vals)[kMaxSlots]) {
#pragma unroll 1
  for (int k = 0; k < nf && k < kMaxSlots; ++k) {
    const WfmFactor f = facs[k];
    if (f.func == WFM_COS_SINCOS) {
      // one range reduction serves every COS factor of this frequency
      double s, c;
      sincos(mul(f.a0, sub(x, f.shift)), &s, &c);
      vals[k] = c;
      vals[k + 1] = s;  // row k+1 is the NOP placeholder (validated at upload)
      ++k;
    } else if (f.func == WFM_COS_ROT) {
      // cos(a_t) with a_t = w*(x - shift) rounded exactly as the reference rounds
      // it, obtained from the base row's (cos, sin)(a_b):  a_t = a_b + D + eps with
      // D a host constant (cos D, sin D tabulated) and eps = (a_t - a_b) - D the
      // MEASURED residual (|eps| ~ ulp(a)), expanded to second order.
      const double* __restrict__ p = args + f.arg_off;
      const int base = (int)p[0];
      const double a_t = mul(f.a0, sub(x, f.shift));
      const double a_b = mul(f.a0, sub(x, p[1]));
      const double eps = sub(sub(a_t, a_b), p[2]);
      const double cb = vals[base], sb = vals[base + 1];
      const double C = fma(cb, p[3], -(sb * p[4]));
      const double S = fma(sb, p[3], cb * p[4]);
      vals[k] = fma(-0.5 * eps * eps, C, fma(-eps, S, C));
    } else {
      vals[k] = eval_factor(f, x, args);
    }
  }
}

// product of the referenced factor powers of an ABI term (general path)
__device__ __forceinline__ double term_product(const IrView& ir, const WfmFactor* facs, const WfmTerm& tm, double x,
                                               const double (&vals)[kMaxSlots]) {
  double prod = 1.0;
  bool first = true;
#pragma unroll 1
  for (int r = 0; r < tm.n_ref; ++r) {
    const WfmRef ref = ir.refs[tm.ref_begin + r];
    double v = (ref.slot < kMaxSlots) ? vals[ref.slot] : eval_factor(facs[ref.slot], x, ir.args);
    if (ref.kind == WFM_POW_INT) v = pow_small_int(v, (int)ref.expo);
    else if (ref.kind == WFM_POW_GEN) v = pow(v, ref.expo);
    prod = first ? v : mul(prod, v);  // 1 * v == v
    first = false;
  }
  return prod;
}

// Evaluate one segment's program at one abscissa (real-valued channels; compact terms).
__device__ __forceinline__ double eval_segment_real(const IrView& ir, const WfmWave& w, WfmSegPtr p0, WfmSegPtr p1,
                                                    double x) {
  double total = w.offset;
  const int nt = p1.term - p0.term;
  if (nt == 0) return total;  // zero segment: untouched by clip (calc_parts skips it)
  double vals[kMaxSlots];
  const WfmFactor* facs = ir.facs + p0.fac;
  eval_factors(facs, p1.fac - p0.fac, x, ir.args, vals);
  double g = 0.0;
  bool g_first = true;
#pragma unroll 1
  for (int it = 0; it < nt; ++it) {
    const CTerm ct = ir.cterms[p0.term + it];
    double prod;
    if (ct.flags & kCTermExt) {
      prod = term_product(ir, facs, ir.terms[p0.term + it], x, vals);
    } else {
      prod = 1.0;
#pragma unroll 1
      for (int r = 0; r < ct.n_ref; ++r) {
        const double v = vals[ct.slot[r]];
        prod = r == 0 ? v : mul(prod, v);
      }
    }
    const double t = mul(ct.amp, prod);
    g = g_first ? t : add(g, t);  // 0 + a == a
    g_first = false;
    if (ct.flags & kCTermGroupEnd) {
      total = add(total, g);
      g_first = true;
    }
  }
  if (w.flags & WFM_WAVE_CLIP) total = fmin(fmax(total, w.clip_lo), w.clip_hi);
  return total;
}

// complex amplitudes (WFM_C128 output): ABI terms from global memory
__device__ __forceinline__ void eval_segment_cplx(const IrView& ir, const WfmWave& w, WfmSegPtr p0, WfmSegPtr p1,
                                                  double x, double& out_re, double& out_im) {
  out_re = w.offset;
  out_im = 0.0;
  const int nt = p1.term - p0.term;
  if (nt == 0) return;
  double vals[kMaxSlots];
  const WfmFactor* facs = ir.facs + p0.fac;
  eval_factors(facs, p1.fac - p0.fac, x, ir.args, vals);
  double g_re = 0.0, g_im = 0.0;
  bool g_first = true;
#pragma unroll 1
  for (int it = 0; it < nt; ++it) {
    const WfmTerm tm = ir.terms[p0.term + it];
    const double prod = term_product(ir, facs, tm, x, vals);
    const double t_re = mul(tm.amp_re, prod), t_im = mul(tm.amp_im, prod);
    g_re = g_first ? t_re : add(g_re, t_re);
    g_im = g_first ? t_im : add(g_im, t_im);
    g_first = false;
    if (tm.flags & WFM_TERM_GROUP_END) {
      out_re = add(out_re, g_re);
      out_im = add(out_im, g_im);
      g_first = true;
    }
  }
  if (w.flags & WFM_WAVE_CLIP) out_re = fmin(fmax(out_re, w.clip_lo), w.clip_hi);
}

template <typename OutT, bool kAccumulate>
__global__ void __launch_bounds__(kThreads) sample_kernel(DevProgram P, const TileDesc* __restrict__ tiles,
                                                          OutT* __restrict__ out) {
  constexpr int V = OutVec<OutT>::N;
  constexpr int kChunk = 32 * V;                      // samples per warp-chunk
  __shared__ __align__(16) unsigned char s_ir[kIrBytes];
  __shared__ WfmSegPtr s_ptr[kStageSegs + 1];
  __shared__ double s_val[kStageSegs];                // value of a FLAT segment
  __shared__ uint16_t s_start[kStageSegs + 1];        // first tile-sample of staged segment k (<= 16384)
  __shared__ uint16_t s_act[kStageSegs + 1];          // active samples before staged segment k
  __shared__ int s_chunk_seg[kMaxChunks];
  __shared__ unsigned char s_active[kStageSegs];
  __shared__ uint64_t s_bar;

  const TileDesc td = tiles[blockIdx.x];
  const WfmWave w = P.waves[td.wave];
  const int64_t j0 = td.j0;
  const int cnt = (int)min((int64_t)P.tile_samples, w.n - j0);
  const int n_chunks = (cnt + kChunk - 1) / kChunk;
  const double* __restrict__ gb = P.seg_bound + w.seg_begin;
  const WfmSegPtr* __restrict__ gp = P.seg_ptr + w.seg_begin;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg_lo = td.seg_lo;
  const int nb = td.seg_hi - seg_lo + 1;
  OutT* __restrict__ dst = out + w.out_off + j0;
  IrView ir{P.facs, P.cterms, P.terms, P.refs, P.args};

  if (nb > kStageSegs) {
    // pathological density (> 1024 segments in one tile): per-sample search in global memory
    for (int jj = threadIdx.x; jj < cnt; jj += kThreads) {
      const double x = abscissa(w, P.x, j0 + jj);
      int lo = seg_lo, hi = td.seg_hi;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(gb + mid) <= x) lo = mid + 1; else hi = mid;
      }
      const double re = eval_segment_real(ir, w, gp[lo], gp[lo + 1], x);
      dst[jj] = kAccumulate ? (OutT)add((double)dst[jj], re) : (OutT)re;
    }
    return;
  }

  // ---- prologue ------------------------------------------------------------------------
  const uint32_t bf = (uint32_t)td.n_fac * sizeof(WfmFactor), bt = (uint32_t)td.n_term * sizeof(CTerm);
  const bool staged = td.n_fac > 0 && bf + bt <= (uint32_t)kIrBytes;
  if (threadIdx.x == 0 && staged) {
    // the tile's slice of the factor and compact-term tables: one TMA bulk copy each
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&s_bar, bf + bt);
    bulk_g2s(s_ir, P.facs + td.fac0, bf, &s_bar);
    bulk_g2s(s_ir + bf, P.cterms + td.term0, bt, &s_bar);
  }
  for (int k = threadIdx.x; k <= nb; k += kThreads) {
    const WfmSegPtr p0 = gp[seg_lo + k];
    s_ptr[k] = p0;
    int pos;
    if (k == 0) pos = 0;
    else if (k == nb) pos = cnt;
    else pos = first_sample_at_or_after(w, P.x, j0, cnt, gb[seg_lo + k - 1]);
    s_start[k] = (uint16_t)pos;
    if (k < nb) {
      const WfmSegPtr p1 = gp[seg_lo + k + 1];
      const bool active = p1.fac > p0.fac;
      s_active[k] = active ? 1 : 0;
      double re = w.offset;
      if (!active && p1.term > p0.term) re = eval_segment_real(ir, w, p0, p1, 0.0);  // constant segment
      s_val[k] = re;
    }
  }
  __syncthreads();
  if (warp == 0) {
    // exclusive prefix of active sample counts over the staged segments
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
      const int k = b0 + lane;
      int c = (k < nb && s_active[k]) ? ((int)s_start[k + 1] - (int)s_start[k]) : 0;
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (k < nb) s_act[k] = (uint16_t)(carry + incl - c);
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_act[nb] = (uint16_t)carry;
  } else {
    // staged segment that owns the first sample of chunk c: last k with s_start[k] <= c*kChunk
    for (int c = threadIdx.x - 32; c < n_chunks; c += kThreads - 32) {
      const int jj = c * kChunk;
      int lo = 0, hi = nb - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_start[mid] <= jj) lo = mid; else hi = mid - 1;
      }
      s_chunk_seg[c] = lo;
    }
  }
  __syncthreads();

  // ---- phase 1: every sample of a FLAT segment (store-bound, no abscissae) ------------
  constexpr int kSuper = 4;  // chunks handled per warp iteration
  for (int sc = warp * kSuper; sc < n_chunks; sc += (kThreads / 32) * kSuper) {
    const int sbeg = sc * kChunk;
    const int send = min(sbeg + kSuper * kChunk, cnt);
    const int ks = s_chunk_seg[sc];
    if (s_start[ks + 1] >= send) {
      // the whole super-chunk lies inside staged segment ks
      if (s_active[ks]) continue;
      const double val = s_val[ks];
#pragma unroll
      for (int q = 0; q < kSuper; ++q) {
        const int base = sbeg + q * kChunk + lane * V;
        double v[V];
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = val;
        if (base + V <= send) {
          if (kAccumulate) {
            double old[V];
            load_vec(dst + base, old);
#pragma unroll
            for (int e = 0; e < V; ++e) v[e] = add(old[e], v[e]);
          }
          store_vec(dst + base, v);
        } else {
          for (int e = 0; e < V && base + e < send; ++e)
            dst[base + e] = kAccumulate ? (OutT)add((double)dst[base + e], v[e]) : (OutT)v[e];
        }
      }
      continue;
    }
    for (int c = sc; c < sc + kSuper && c < n_chunks; ++c) {
      const int cbeg = c * kChunk;
      const int cend = min(cbeg + kChunk, cnt);
      const int base = cbeg + lane * V;
      const int k0 = s_chunk_seg[c];
      double v[V];
      if (s_start[k0 + 1] >= cend) {  // whole chunk inside staged segment k0
        if (s_active[k0]) continue;
        const double val = s_val[k0];
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = val;
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
        continue;
      }
      // a bound falls inside the chunk: lanes advance from the chunk's first segment
      int k = k0;
      bool flat[V];
      bool all_flat = base + V <= cend;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int jj = base + e;
        flat[e] = false;
        v[e] = 0.0;
        if (jj < cend) {
          while (k < nb - 1 && s_start[k + 1] <= jj) ++k;
          flat[e] = !s_active[k];
          v[e] = s_val[k];
        }
        all_flat = all_flat && flat[e];
      }
      if (all_flat) {
        if (kAccumulate) {
          double old[V];
          load_vec(dst + base, old);
#pragma unroll
          for (int e = 0; e < V; ++e) v[e] = add(old[e], v[e]);
        }
        store_vec(dst + base, v);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (flat[e]) {
            if (kAccumulate) dst[base + e] = (OutT)add((double)dst[base + e], v[e]);
            else store_one(dst + base + e, v[e]);
          }
      }
    }
  }

  // ---- phase 2: the tile's ACTIVE samples, dealt evenly to all threads ------------------
  const int n_active = s_act[nb];
  if (staged) mbar_wait(&s_bar, 0);  // also guarantees no copy is in flight when the CTA retires
  if (n_active == 0) return;
  if (staged) {
    ir.facs = reinterpret_cast<const WfmFactor*>(s_ir) - td.fac0;
    ir.cterms = reinterpret_cast<const CTerm*>(s_ir + bf) - td.term0;
  }
  for (int i = threadIdx.x; i < n_active; i += kThreads) {
    int lo = 0, hi = nb - 1;  // last k with s_act[k] <= i: the active segment holding sample i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_act[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const int jj = (int)s_start[lo] + (i - (int)s_act[lo]);
    const double x = abscissa(w, P.x, j0 + jj);
    const double re = eval_segment_real(ir, w, s_ptr[lo], s_ptr[lo + 1], x);
    if (kAccumulate) dst[jj] = (OutT)add((double)dst[jj], re);
    else store_one(dst + jj, re);
  }
}

// complex128 output: interleaved (re, im); one sample per thread per row.
template <bool kAccumulate>
__global__ void __launch_bounds__(kThreads) sample_kernel_c128(DevProgram P, const TileDesc* __restrict__ tiles,
                                                               double2* __restrict__ out) {
  const TileDesc td = tiles[blockIdx.x];
  const WfmWave w = P.waves[td.wave];
  const int64_t j0 = td.j0;
  const int cnt = (int)min((int64_t)P.tile_samples, w.n - j0);
  const double* __restrict__ gb = P.seg_bound + w.seg_begin;
  const WfmSegPtr* __restrict__ gp = P.seg_ptr + w.seg_begin;
  double2* __restrict__ dst = out + w.out_off + j0;
  const IrView ir{P.facs, P.cterms, P.terms, P.refs, P.args};
  int seg = td.seg_lo;
  for (int jj = threadIdx.x; jj < cnt; jj += kThreads) {
    const double x = abscissa(w, P.x, j0 + jj);
    int lo = seg, hi = td.seg_hi;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (__ldg(gb + mid) <= x) lo = mid + 1; else hi = mid;
    }
    seg = lo;
    double re, im;
    eval_segment_cplx(ir, w, gp[seg], gp[seg + 1], x, re, im);
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

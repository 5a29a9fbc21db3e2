// wfm_sample.cu — K1: batched piecewise-waveform sampling kernel for sm_100a.
//
// Replaces, for a whole batch of channels in one launch, the reference's
//   np.arange -> calc_parts (np.searchsorted, per-segment _calc/_apply,
//   np.clip) -> zeros_like -> _fill_parts
// (/root/reference/waveforms/waveform.py:173-207, :529-563, :679-693;
//  /root/reference/waveforms/_waveform.pyx:130-169).
//
// Once per program (prepare_segments_kernel / prepare_tiles_kernel):
//   * every segment learns the INTEGER sample position where it starts: the first
//     sample whose abscissa is >= its lower bound — a division for the guess,
//     then exact comparisons against the rounded grid value x[j] = t0 + j*delta,
//     so ownership is bit-identical to np.searchsorted on the reference's grid;
//   * every FLAT segment (no basis factor: zero, or a constant) gets its value;
//   * the ABI factor / term rows are rewritten into the device formats (64-byte
//     factor rows carrying their rotation constants, 16-byte compact terms);
//   * every tile (tile_samples consecutive samples of one channel) learns the
//     segment range it spans and its slice of the factor / term tables.
//
// Per launch, one CTA (256 threads) per tile; the tile is ASSEMBLED IN SHARED
// MEMORY and leaves the SM as ONE TMA bulk store (cp.async.bulk shared->global):
//   1. Prologue.  One thread starts a TMA bulk load (cp.async.bulk + mbarrier) of
//      the tile's slice of the factor / compact-term tables; every thread copies
//      one row of the segment tables (start position, flat value, pointers).
//   2. Flat fill.  The whole shared tile is first filled with the channel's
//      zero-segment value (16-byte shared stores, no table look-ups, overlapping the
//      table loads); warp 1 then compacts the list of flat segments whose value
//      differs (constant plateaus) and those runs are rewritten one warp per run.
//      No abscissa is computed for flat samples.
//   3. Active samples.  The ACTIVE samples of the tile are enumerated through a
//      prefix sum over the segments and dealt round-robin to all 256 threads, so
//      a tile with one 40-sample pulse keeps 40 lanes busy once and a dense tile
//      keeps every lane busy.  Each thread interprets its sample's segment
//      program (distinct factors into per-thread value slots in shared memory,
//      then terms referencing the slots) and writes the sample into the tile.
//   4. Store.  fence.proxy.async, barrier, one elected thread issues the bulk
//      store of the whole tile with an L2 evict-first policy (the output is
//      write-once; it must not displace the IR).  HBM sees full lines only, no
//      LSU store instructions are spent on the output, and the store drains
//      while the other resident CTAs of the SM compute.
//
// Algorithmic traffic: 8 B (4 B) per sample, write-only.
#include <cuda_runtime.h>
#include <stdint.h>
#include "wfm_basis.cuh"
#include "wfm_multidrag.cuh"
#include "wfm_internal.h"

namespace wfm {

#ifndef WFM_K1_THREADS
#define WFM_K1_THREADS 256
#endif
constexpr int kThreads = WFM_K1_THREADS;
#ifndef WFM_K1_MIN_BLOCKS
#define WFM_K1_MIN_BLOCKS 3  // resident CTAs per SM the register allocation is sized for
#endif
constexpr int kStageSegs = 256;  // segment rows staged per tile

template <typename T> struct OutVec;
template <> struct OutVec<double> { static constexpr int N = 2; };
template <> struct OutVec<float> { static constexpr int N = 4; };

// ---- shared-memory vector stores ---------------------------------------------------
__device__ __forceinline__ void fill_vec(double* p, double v) { *reinterpret_cast<double2*>(p) = make_double2(v, v); }
__device__ __forceinline__ void fill_vec(float* p, double v) {
  const float f = (float)v;
  *reinterpret_cast<float4*>(p) = make_float4(f, f, f, f);
}
// ---- mbarrier + TMA bulk copies ------------------------------------------------------
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
// shared -> global, write-once data: L2 evict-first
__device__ __forceinline__ void bulk_s2g_evict_first(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes), "l"(policy)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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

// first sample j in [0, n] of the channel with abscissa >= bound (n if none)
__device__ int first_sample_at_or_after(const WfmWave& w, const double* __restrict__ xs, int n, double bound) {
  if (!(w.flags & WFM_WAVE_EXPLICIT_X) && w.delta > 0.0) {
    double b = bound;
    if (w.flags & WFM_WAVE_PRESHIFT) b = b + w.pre_shift;
    const double g = ceil((b - w.t0) / w.delta);
    int j = g <= 0.0 ? 0 : (g >= (double)n ? n : (int)g);  // NaN -> n
    // exact fix-up against the rounded grid (monotone in j)
    while (j > 0 && abscissa(w, xs, j - 1) >= bound) --j;
    while (j < n && abscissa(w, xs, j) < bound) ++j;
    return j;
  }
  int lo = 0, hi = n;  // generic: binary search over the channel's abscissae
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (abscissa(w, xs, mid) < bound) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// global tables the interpreter may fall back to (extended terms, rows beyond the
// value cache, argument pool of the cold basis functions)
struct IrGlobal {
  const DFactor* dfacs;
  const WfmTerm* terms;
  const WfmRef* refs;
  const double* args;
};

// the few channel fields a sample evaluation needs
struct WaveEval {
  double offset, clip_lo, clip_hi;
  uint32_t flags;
};

// per-thread cache of the distinct factor values of one segment evaluation
struct LocalSlots {  // registers / local memory: fallback and complex kernels
  double v[kMaxSlots];
  __device__ __forceinline__ double get(int k) const { return v[k]; }
  __device__ __forceinline__ void set(int k, double x) { v[k] = x; }
};
struct SmemSlots {  // shared memory, slot-major: slot k of thread t at p[k*kThreads] (conflict-free)
  double* p;
  __device__ __forceinline__ double get(int k) const { return p[k * kThreads]; }
  __device__ __forceinline__ void set(int k, double x) const { p[k * kThreads] = x; }
};

__device__ __forceinline__ FacArgs fac_args(const DFactor& f) { return FacArgs{f.func, f.aux, f.shift, f.a0, f.a1}; }

// distinct factor values of one segment -> slots [0 .. min(nf, kMaxSlots));
// facs = the segment's first row (shared memory when the tile's slice is staged)
template <typename Slots>
__device__ __forceinline__ void eval_factors(const DFactor* __restrict__ facs, int nf, double x,
                                             const double* __restrict__ args, Slots& vals) {
#pragma unroll 1
  for (int k = 0; k < nf && k < kMaxSlots; ++k) {
    const int func = facs[k].func;
    const double shift = facs[k].shift, a0 = facs[k].a0;
    if (func == WFM_COS_SINCOS) {
      // one range reduction serves every COS factor of this frequency
      double s, c;
      sincos_cw(mul(a0, sub(x, shift)), &s, &c);
      vals.set(k, c);
      vals.set(k + 1, s);  // row k+1 is the NOP placeholder (validated at upload)
      ++k;
    } else if (func == WFM_COS_ROT) {
      // cos(a_t) with a_t = w*(x - shift) rounded exactly as the reference rounds
      // it, obtained from the base row's (cos, sin)(a_b):  a_t = a_b + D + eps with
      // D a host constant (cos D, sin D tabulated) and eps = (a_t - a_b) - D the
      // MEASURED residual (|eps| ~ ulp(a)), expanded to second order.
      const int base = facs[k].aux;
      const double bshift = facs[k].p[0], D = facs[k].p[1], cD = facs[k].p[2], sD = facs[k].p[3];
      const double a_t = mul(a0, sub(x, shift));
      const double a_b = mul(a0, sub(x, bshift));
      const double eps = sub(sub(a_t, a_b), D);
      const double cb = vals.get(base), sb = vals.get(base + 1);
      const double C = fma(cb, cD, -(sb * sD));
      const double S = fma(sb, cD, cb * sD);
      vals.set(k, fma(-0.5 * eps * eps, C, fma(-eps, S, C)));
    } else if (func == WFM_COS) {
      vals.set(k, cos_cw(mul(a0, sub(x, shift))));
    } else if (func == WFM_LINEAR) {
      vals.set(k, sub(x, shift));
    } else if (func == WFM_GAUSSIAN) {
      vals.set(k, f_gaussian(sub(x, shift), a0));
    } else if (func == WFM_ERF) {
      vals.set(k, erf(dvd(sub(x, shift), a0)));
    } else {
      vals.set(k, eval_factor(fac_args(facs[k]), x, args));
    }
  }
}

// product of the referenced factor powers of an ABI term (general path); gfac = the
// segment's first row in the GLOBAL factor table
template <typename Slots>
__device__ __forceinline__ double term_product(const IrGlobal& g, int gfac, const WfmTerm& tm, double x,
                                               const Slots& vals) {
  double prod = 1.0;
  bool first = true;
#pragma unroll 1
  for (int r = 0; r < tm.n_ref; ++r) {
    const WfmRef ref = g.refs[tm.ref_begin + r];
    double v = (ref.slot < kMaxSlots) ? vals.get(ref.slot) : eval_factor(fac_args(g.dfacs[gfac + ref.slot]), x, g.args);
    if (ref.kind == WFM_POW_INT) v = pow_small_int(v, (int)ref.expo);
    else if (ref.kind == WFM_POW_GEN) v = pow(v, ref.expo);
    prod = first ? v : mul(prod, v);  // 1 * v == v
    first = false;
  }
  return prod;
}

// Evaluate one segment's program at one abscissa (real-valued channels; compact terms).
// facs / cterms point at the segment's first rows (shared or global memory);
// gfac / gterm are the same rows' indices in the global tables.
template <typename Slots>
__device__ __forceinline__ double eval_segment_real(const DFactor* __restrict__ facs, const CTerm* __restrict__ cterms,
                                                    int nf, int nt, const IrGlobal& g, int gfac, int gterm,
                                                    const WaveEval& w, double x, Slots& vals) {
  double total = w.offset;
  if (nt == 0) return total;  // zero segment: untouched by clip (calc_parts skips it)
  eval_factors(facs, nf, x, g.args, vals);
  double grp = 0.0;
  bool g_first = true;
#pragma unroll 1
  for (int it = 0; it < nt; ++it) {
    const double amp = cterms[it].amp;
    const uint64_t pk = cterms[it].packed;
    const uint32_t flags = (uint32_t)(pk >> 8) & 0xffu;
    double prod;
    if (flags & kCTermExt) {
      prod = term_product(g, gfac, g.terms[gterm + it], x, vals);
    } else {
      const int n_ref = (int)(pk & 0xffu);
      uint32_t slots_lo = (uint32_t)(pk >> 16), slots_hi = (uint32_t)(pk >> 48);
      prod = 1.0;
#pragma unroll 1
      for (int r = 0; r < n_ref; ++r) {
        const double v = vals.get((int)(slots_lo & 0xffu));
        slots_lo = (slots_lo >> 8) | (slots_hi << 24);
        slots_hi >>= 8;
        prod = r == 0 ? v : mul(prod, v);
      }
    }
    const double t = mul(amp, prod);
    grp = g_first ? t : add(grp, t);  // 0 + a == a
    g_first = false;
    if (flags & kCTermGroupEnd) {
      total = add(total, grp);
      g_first = true;
    }
  }
  if (w.flags & WFM_WAVE_CLIP) total = fmin(fmax(total, w.clip_lo), w.clip_hi);
  return total;
}

// complex amplitudes (WFM_C128 output): ABI terms from global memory
__device__ __forceinline__ void eval_segment_cplx(const IrGlobal& g, const WaveEval& w, WfmSegPtr p0, WfmSegPtr p1,
                                                  double x, double& out_re, double& out_im) {
  out_re = w.offset;
  out_im = 0.0;
  const int nt = p1.term - p0.term;
  if (nt == 0) return;
  LocalSlots vals;
  eval_factors(g.dfacs + p0.fac, p1.fac - p0.fac, x, g.args, vals);
  double g_re = 0.0, g_im = 0.0;
  bool g_first = true;
#pragma unroll 1
  for (int it = 0; it < nt; ++it) {
    const WfmTerm tm = g.terms[p0.term + it];
    const double prod = term_product(g, p0.fac, tm, x, vals);
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

// ---- pre-pass (once per program) ----------------------------------------------------------
// one thread per segment: start position, value of a flat segment
__global__ void prepare_segments_kernel(DevProgram P, int32_t* __restrict__ seg_start, double* __restrict__ seg_val,
                                        int64_t n_segs) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segs) return;
  const WfmWave w = P.waves[P.seg_wave[s]];
  const int k = (int)(s - w.seg_begin);
  seg_start[s] = k == 0 ? 0 : first_sample_at_or_after(w, P.x, (int)w.n, P.seg_bound[s - 1]);
  const WfmSegPtr p0 = P.seg_ptr[s], p1 = P.seg_ptr[s + 1];
  double val = w.offset;
  if (p1.fac == p0.fac && p1.term > p0.term) {
    // constant segment: offset + sum over stack members of (sum of their constant terms)
    double grp = 0.0;
    bool g_first = true;
    for (int t = p0.term; t < p1.term; ++t) {
      const WfmTerm tm = P.terms[t];
      double c = tm.amp_re;
      if (tm.n_ref > 0) c = CUDART_NAN;  // cannot happen: a term with factors makes the segment active
      grp = g_first ? c : add(grp, c);
      g_first = false;
      if (tm.flags & WFM_TERM_GROUP_END) {
        val = add(val, grp);
        g_first = true;
      }
    }
    if (w.flags & WFM_WAVE_CLIP) val = fmin(fmax(val, w.clip_lo), w.clip_hi);
  }
  seg_val[s] = val;
}

// one thread per factor row: the 64-byte device row
__global__ void prepare_factors_kernel(DevProgram P, DFactor* __restrict__ dfacs, int64_t n_facs) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_facs) return;
  const WfmFactor f = P.facs[k];
  DFactor d;
  d.func = f.func;
  d.aux = f.arg_off;
  d.shift = f.shift;
  d.a0 = f.a0;
  d.a1 = f.a1;
  d.p[0] = d.p[1] = d.p[2] = d.p[3] = 0.0;
  if (f.func == WFM_COS_ROT) {
    const double* __restrict__ p = P.args + f.arg_off;  // [base_slot, base_shift, D, cos D, sin D]
    d.aux = (int)p[0];
    d.p[0] = p[1];
    d.p[1] = p[2];
    d.p[2] = p[3];
    d.p[3] = p[4];
  }
  dfacs[k] = d;
}

// one thread per term: the 16-byte compact term
__global__ void prepare_terms_kernel(DevProgram P, CTerm* __restrict__ cterms, int64_t n_terms) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_terms) return;
  const WfmTerm tm = P.terms[t];
  uint64_t packed = 0;
  bool ext = tm.n_ref > 6;
  for (int r = 0; r < tm.n_ref && !ext; ++r) {
    const WfmRef rf = P.refs[tm.ref_begin + r];
    if (rf.kind != WFM_POW_ONE || rf.slot >= kMaxSlots) ext = true;
    else packed |= (uint64_t)(uint32_t)rf.slot << (16 + 8 * r);
  }
  uint32_t flags = (tm.flags & WFM_TERM_GROUP_END) ? kCTermGroupEnd : 0u;
  if (ext) {
    flags |= kCTermExt;
    packed = 0;
  } else {
    packed |= (uint64_t)(uint32_t)tm.n_ref;
  }
  packed |= (uint64_t)flags << 8;
  cterms[t] = CTerm{tm.amp_re, packed};
}

// last segment k in [0, n) with start[k] <= j (start[0] == 0)
__device__ __forceinline__ int owning_segment(const int32_t* __restrict__ start, int n, int64_t j) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((int64_t)start[mid] <= j) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// one thread per tile: the segment rows it spans and its slice of the tables
__global__ void prepare_tiles_kernel(DevProgram P, TileDesc* __restrict__ tiles, int64_t n_tiles,
                                     int* __restrict__ max_ir_bytes) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  TileDesc td = tiles[t];
  const WfmWave w = P.waves[td.wave];
  const int32_t* st = P.seg_start + w.seg_begin;
  const int lo = owning_segment(st, w.n_seg, td.j0);
  const int hi = max(lo, owning_segment(st, w.n_seg, td.j0 + td.cnt - 1));
  td.seg0 = w.seg_begin + lo;
  td.nb = hi - lo + 1;
  const WfmSegPtr a = P.seg_ptr[w.seg_begin + lo], e = P.seg_ptr[w.seg_begin + hi + 1];
  td.fac0 = a.fac;
  td.n_fac = e.fac - a.fac;
  td.term0 = a.term;
  td.n_term = e.term - a.term;
  tiles[t] = td;
  if (td.n_fac > 0 && td.nb <= kStageSegs)
    atomicMax(max_ir_bytes, td.n_fac * (int)sizeof(DFactor) + td.n_term * (int)sizeof(CTerm));
}

// ---- the sampling kernel ------------------------------------------------------------------
// [a, b) of the shared tile <- val, executed by one warp
template <typename OutT>
__device__ __forceinline__ void fill_run(OutT* __restrict__ s_out, int a, int b, double val, int lane) {
  constexpr int V = OutVec<OutT>::N;
  const int a_al = (a + V - 1) & ~(V - 1), b_al = b & ~(V - 1);
  if (a_al >= b_al) {
    for (int p = a + lane; p < b; p += 32) s_out[p] = (OutT)val;
    return;
  }
  if (a + lane < a_al) s_out[a + lane] = (OutT)val;
  if (b_al + lane < b) s_out[b_al + lane] = (OutT)val;
#pragma unroll 1
  for (int p = a_al + lane * V; p < b_al; p += 32 * V) fill_vec(s_out + p, val);
}

extern __shared__ __align__(128) unsigned char k1_smem[];

template <typename OutT, bool kAccumulate>
__global__ void __launch_bounds__(kThreads, WFM_K1_MIN_BLOCKS) sample_kernel(DevProgram P, const TileDesc* __restrict__ tiles,
                                                          OutT* __restrict__ out) {
  constexpr int V = OutVec<OutT>::N;
  // dynamic: [tile: tile_samples x OutT][value slots: n_slots x 256 x f64][IR slice: ir_bytes]
  OutT* s_out = reinterpret_cast<OutT*>(k1_smem);
  double* s_slots = reinterpret_cast<double*>(k1_smem + (size_t)P.tile_samples * sizeof(OutT));
  unsigned char* s_ir = reinterpret_cast<unsigned char*>(s_slots + (size_t)P.n_slots * kThreads);
  __shared__ WfmSegPtr s_ptr[kStageSegs + 1];
  __shared__ double s_val[kStageSegs];           // value of a FLAT segment
  __shared__ uint16_t s_start[kStageSegs + 1];   // first tile-sample of staged segment k
  __shared__ uint16_t s_act[kStageSegs + 1];     // active samples before staged segment k
  __shared__ uint16_t s_patch[kStageSegs];       // flat, non-empty segments whose value differs from the base fill
  __shared__ int s_npatch;
  __shared__ WfmWave s_wave;
  __shared__ uint64_t s_bar;

  const TileDesc td = tiles[blockIdx.x];
  const int cnt = td.cnt;
  const int nb = td.nb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  OutT* __restrict__ dst = out + td.out0;
  const IrGlobal g{P.dfacs, P.terms, P.refs, P.args};

  const uint32_t bf = (uint32_t)td.n_fac * sizeof(DFactor), bt = (uint32_t)td.n_term * sizeof(CTerm);
  if (nb > kStageSegs || (td.n_fac > 0 && bf + bt > (uint32_t)P.ir_bytes)) {
    // pathological density (> 256 segment rows, or a table slice beyond the shared-memory
    // budget, in one tile): per-sample search, tables in global memory, direct stores
    const WfmWave w = P.waves[td.wave];
    const WaveEval we{w.offset, w.clip_lo, w.clip_hi, w.flags};
    const int32_t* __restrict__ st = P.seg_start + td.seg0;
    const WfmSegPtr* __restrict__ gp = P.seg_ptr + td.seg0;
    for (int jj = threadIdx.x; jj < cnt; jj += kThreads) {
      int lo = 0, hi = nb - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int64_t)st[mid] <= td.j0 + jj) lo = mid; else hi = mid - 1;
      }
      const WfmSegPtr p0 = gp[lo], p1 = gp[lo + 1];
      LocalSlots vals;
      const double re = eval_segment_real(P.dfacs + p0.fac, P.cterms + p0.term, p1.fac - p0.fac, p1.term - p0.term, g,
                                          p0.fac, p0.term, we, abscissa(w, P.x, td.j0 + jj), vals);
      dst[jj] = kAccumulate ? (OutT)add((double)dst[jj], re) : (OutT)re;
    }
    return;
  }

  // ---- prologue ------------------------------------------------------------------------
  const bool staged = td.n_fac > 0;
  if (threadIdx.x == 0 && staged) {
    // the tile's slice of the factor and compact-term tables: one TMA bulk copy each
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&s_bar, bf + bt);
    bulk_g2s(s_ir, P.dfacs + td.fac0, bf, &s_bar);
    bulk_g2s(s_ir + bf, P.cterms + td.term0, bt, &s_bar);
  }
  const double base = P.waves[td.wave].offset;  // value of every zero segment
  {
    // segment rows of the tile (nothing here depends on the channel record)
    const int32_t* __restrict__ gs = P.seg_start + td.seg0;
    const double* __restrict__ gv = P.seg_val + td.seg0;
    const WfmSegPtr* __restrict__ gp = P.seg_ptr + td.seg0;
    for (int k = threadIdx.x; k <= nb; k += kThreads) {
      s_ptr[k] = gp[k];
      int pos;
      if (k == 0) pos = 0;
      else if (k == nb) pos = cnt;
      else pos = (int)min((int64_t)cnt, max((int64_t)0, (int64_t)gs[k] - td.j0));
      s_start[k] = (uint16_t)pos;
      if (k < nb) s_val[k] = gv[k];
    }
    if (threadIdx.x >= kThreads - (int)(sizeof(WfmWave) / 8)) {
      const int q = threadIdx.x - (kThreads - (int)(sizeof(WfmWave) / 8));
      reinterpret_cast<uint64_t*>(&s_wave)[q] = reinterpret_cast<const uint64_t*>(P.waves + td.wave)[q];
    }
  }
  // base fill: the whole tile <- the zero-segment value; flat segments with another
  // value and the active samples overwrite it below
  {
    const int n_fill = (cnt + V - 1) & ~(V - 1);  // the tile buffer is a multiple of V
#pragma unroll 1
    for (int p = threadIdx.x * V; p < n_fill; p += kThreads * V) fill_vec(s_out + p, base);
  }
  __syncthreads();
  if (warp == 0) {
    // exclusive prefix of ACTIVE sample counts over the staged segments
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
      const int k = b0 + lane;
      int c = 0;
      if (k < nb && s_ptr[k + 1].fac > s_ptr[k].fac) c = (int)s_start[k + 1] - (int)s_start[k];
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
  } else if (warp == 1) {
    // compacted list of the flat segments the base fill did not already serve
    int n = 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
      const int k = b0 + lane;
      bool need = false;
      if (k < nb && s_start[k + 1] > s_start[k] && s_ptr[k + 1].fac == s_ptr[k].fac)
        need = __double_as_longlong(s_val[k]) != __double_as_longlong(base);
      const unsigned m = __ballot_sync(0xffffffffu, need);
      if (need) s_patch[n + __popc(m & ((1u << lane) - 1u))] = (uint16_t)k;
      n += __popc(m);
    }
    if (lane == 0) s_npatch = n;
  }
  __syncthreads();

  // ---- flat segments with their own value: one warp per run ------------------------------
  {
    const int n_patch = s_npatch;
    for (int i = warp; i < n_patch; i += kThreads / 32) {
      const int k = s_patch[i];
      fill_run(s_out, (int)s_start[k], (int)s_start[k + 1], s_val[k], lane);
    }
  }

  // ---- the tile's ACTIVE samples, dealt evenly to all threads ------------------------------
  const int n_active = s_act[nb];
  if (staged) mbar_wait(&s_bar, 0);  // also guarantees no copy is in flight when the CTA retires
  if (n_active > 0) {
    const WaveEval we{s_wave.offset, s_wave.clip_lo, s_wave.clip_hi, s_wave.flags};
    const uint32_t wflags = s_wave.flags;
    const double t0 = s_wave.t0, delta = s_wave.delta;
    const bool plain_grid = !(wflags & (WFM_WAVE_EXPLICIT_X | WFM_WAVE_LAST_OVERRIDE | WFM_WAVE_PRESHIFT));
    SmemSlots vals{s_slots + threadIdx.x};
    const DFactor* __restrict__ sf = reinterpret_cast<const DFactor*>(s_ir);
    const CTerm* __restrict__ sc = reinterpret_cast<const CTerm*>(s_ir + bf);
#pragma unroll 1
    for (int i = threadIdx.x; i < n_active; i += kThreads) {
      int lo = 0, hi = nb - 1;  // last k with s_act[k] <= i: the active segment holding sample i
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_act[mid] <= i) lo = mid; else hi = mid - 1;
      }
      const int jj = (int)s_start[lo] + (i - (int)s_act[lo]);
      const double x = plain_grid ? add(t0, mul((double)(td.j0 + jj), delta)) : abscissa(s_wave, P.x, td.j0 + jj);
      const WfmSegPtr p0 = s_ptr[lo], p1 = s_ptr[lo + 1];
      s_out[jj] = (OutT)eval_segment_real(sf + (p0.fac - td.fac0), sc + (p0.term - td.term0), p1.fac - p0.fac,
                                          p1.term - p0.term, g, p0.fac, p0.term, we, x, vals);
    }
  }

  // ---- store ---------------------------------------------------------------------------------
  if (kAccumulate) {
    // out += tile (Waveform.__call__(..., accumulate=True)): read-modify-write epilogue
    __syncthreads();
    for (int p = threadIdx.x; p < cnt; p += kThreads) dst[p] = (OutT)add((double)dst[p], (double)s_out[p]);
    return;
  }
  // the whole tile as one TMA bulk copy
  fence_proxy_async_smem();
  __syncthreads();
  const int n_bulk = cnt & ~(V - 1);  // 16-byte multiple; the ragged tail goes out as scalars
  if (threadIdx.x == 0 && n_bulk > 0) bulk_s2g_evict_first(dst, s_out, (uint32_t)n_bulk * sizeof(OutT));
  if (n_bulk + (int)threadIdx.x < cnt) dst[n_bulk + threadIdx.x] = s_out[n_bulk + threadIdx.x];
  if (threadIdx.x == 0) bulk_wait_read_all();  // shared memory must outlive the copy's reads
}

// complex128 output: interleaved (re, im); one sample per thread per row.
template <bool kAccumulate>
__global__ void __launch_bounds__(kThreads) sample_kernel_c128(DevProgram P, const TileDesc* __restrict__ tiles,
                                                               double2* __restrict__ out) {
  const TileDesc td = tiles[blockIdx.x];
  const WfmWave w = P.waves[td.wave];
  const WaveEval we{w.offset, w.clip_lo, w.clip_hi, w.flags};
  const int32_t* __restrict__ st = P.seg_start + td.seg0;
  const WfmSegPtr* __restrict__ gp = P.seg_ptr + td.seg0;
  double2* __restrict__ dst = out + td.out0;
  const IrGlobal g{P.dfacs, P.terms, P.refs, P.args};
  for (int jj = threadIdx.x; jj < td.cnt; jj += kThreads) {
    const double x = abscissa(w, P.x, td.j0 + jj);
    int lo = 0, hi = td.nb - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if ((int64_t)st[mid] <= td.j0 + jj) lo = mid; else hi = mid - 1;
    }
    double re, im;
    eval_segment_cplx(g, we, gp[lo], gp[lo + 1], x, re, im);
    if (kAccumulate) {
      double2 o = dst[jj];
      re = add(o.x, re);
      im = add(o.y, im);
    }
    dst[jj] = make_double2(re, im);
  }
}

cudaError_t launch_prepare(const DevProgram& P, const PrepareCounts& n, int32_t* seg_start, double* seg_val,
                           DFactor* dfacs, CTerm* cterms, TileDesc* tiles, int* max_ir_bytes, cudaStream_t stream) {
  const int threads = 128;
  auto blocks = [&](int64_t items) { return (unsigned)((items + threads - 1) / threads); };
  if (n.n_segs > 0) prepare_segments_kernel<<<blocks(n.n_segs), threads, 0, stream>>>(P, seg_start, seg_val, n.n_segs);
  if (n.n_facs > 0) prepare_factors_kernel<<<blocks(n.n_facs), threads, 0, stream>>>(P, dfacs, n.n_facs);
  if (n.n_terms > 0) prepare_terms_kernel<<<blocks(n.n_terms), threads, 0, stream>>>(P, cterms, n.n_terms);
  if (n.n_tiles > 0) prepare_tiles_kernel<<<blocks(n.n_tiles), threads, 0, stream>>>(P, tiles, n.n_tiles, max_ir_bytes);
  return cudaGetLastError();
}

size_t sample_smem_bytes(const DevProgram& P, int dtype) {
  const size_t esz = dtype == WFM_F32 ? 4 : 8;
  return (size_t)P.tile_samples * esz + (size_t)P.n_slots * kThreads * sizeof(double) + (size_t)P.ir_bytes;
}

cudaError_t launch_sample(const DevProgram& P, const TileDesc* tiles, int64_t n_tiles, int dtype, int accumulate,
                          void* out, cudaStream_t stream) {
  if (n_tiles == 0) return cudaSuccess;
  dim3 grid((unsigned)n_tiles), block(kThreads);
  const size_t smem = sample_smem_bytes(P, dtype);
  cudaError_t e = cudaSuccess;
  if (dtype == WFM_F64) {
    auto k = accumulate ? sample_kernel<double, true> : sample_kernel<double, false>;
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<grid, block, smem, stream>>>(P, tiles, (double*)out);
  } else if (dtype == WFM_F32) {
    auto k = accumulate ? sample_kernel<float, true> : sample_kernel<float, false>;
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    k<<<grid, block, smem, stream>>>(P, tiles, (float*)out);
  } else {
    if (accumulate) sample_kernel_c128<true><<<grid, block, 0, stream>>>(P, tiles, (double2*)out);
    else sample_kernel_c128<false><<<grid, block, 0, stream>>>(P, tiles, (double2*)out);
  }
  return cudaGetLastError();
}

}  // namespace wfm

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
// Per launch: a PERSISTENT grid (CTAs per SM x 148) of 8-warp CTAs in which every
// WARP is autonomous.  A warp owns a private slice of shared memory (output tile,
// factor-value slots, table slice, segment rows, one mbarrier) and loops over
// tiles (warp w of the grid takes tiles w, w+G, ...); there is NO block-level
// barrier anywhere, so a warp stalled on a table load or a long pulse never
// idles its neighbours.  The tile is ASSEMBLED IN SHARED MEMORY and leaves the
// SM as ONE TMA bulk store (cp.async.bulk shared->global):
//   1. Prologue.  Lane 0 starts a TMA bulk load (cp.async.bulk + mbarrier) of
//      the tile's slice of the factor / compact-term tables; the lanes copy the
//      tile's segment rows (start position, flat value, pointers); the NEXT tile's
//      descriptor is prefetched.
//   2. Flat fill.  After the previous tile's bulk store has finished reading the
//      buffer, the whole tile is filled with the channel's zero-segment value
//      (16-byte shared stores, no table look-ups); a ballot compacts the list of
//      flat segments whose value differs (constant plateaus) and those runs are
//      rewritten.  No abscissa is computed for flat samples.
//   3. Active samples.  The ACTIVE samples of the tile are enumerated through a
//      warp prefix sum over the segment rows and dealt round-robin to the 32
//      lanes.  Each lane interprets its sample's segment program (distinct
//      factors into per-lane value slots in shared memory, then terms referencing
//      the slots) and writes the sample into the tile.
//   4. Store.  fence.proxy.async, __syncwarp, lane 0 issues the bulk store of the
//      whole tile with an L2 evict-first policy (the output is write-once; it must
//      not displace the IR).  HBM sees full lines only, no LSU store instructions
//      are spent on the output, and the store drains while the warp is already in
//      the prologue of its next tile.
//
// Algorithmic traffic: 8 B (4 B) per sample, write-only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include "wfm_basis.cuh"
#include "wfm_multidrag.cuh"
#include "wfm_internal.h"

namespace wfm {

constexpr int kThreads = 256;  // 8 autonomous warps per CTA
constexpr int kWarpsPerCta = kThreads / 32;

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

// per-lane cache of the distinct factor values of one segment evaluation.  PHYSICAL
// slot 0 holds the constant 1.0; the value of factor row k lives in physical slot k+1.
struct LocalSlots {  // registers / local memory: cold path and complex kernel
  double v[kMaxSlots + 1];
  __device__ __forceinline__ LocalSlots() { v[0] = 1.0; }
  __device__ __forceinline__ double phys(int k) const { return v[k]; }
  __device__ __forceinline__ void setp(int k, double x) { v[k] = x; }
};
struct SmemSlots {  // the warp's shared slice, slot-major: physical slot k of lane l at p[k*32] (conflict-free)
  double* p;
  __device__ __forceinline__ double phys(int k) const { return p[k * 32]; }
  __device__ __forceinline__ void setp(int k, double x) const { p[k * 32] = x; }
};

__device__ __forceinline__ FacArgs fac_args(const DFactor& f) {
  return FacArgs{f.func & 0xffff, f.aux, f.shift, f.a0, f.a1};
}

// factor rows of one segment -> value slots.  Every row names its destination
// (physical) slot, so packets can drop the NOP placeholder rows; facs = the
// segment's first row (shared memory when the tile's packet is staged).
template <typename Slots>
__device__ __forceinline__ void eval_factors(const DFactor* __restrict__ facs, int n_rows, double x,
                                             const double* __restrict__ args, Slots& vals) {
#pragma unroll 1
  for (int k = 0; k < n_rows; ++k) {
    const uint32_t fo = (uint32_t)facs[k].func;
    const int op = (fo >> 16) & 0xff;
    const int dest = fo >> 24;
    const double a0 = facs[k].a0;
    const double t = sub(x, facs[k].shift);
    if (op == OP_ROT) {
      // cos(a_t) with a_t = w*(x - shift) rounded exactly as the reference rounds
      // it, obtained from the base row's (cos, sin)(a_b):  a_t = a_b + D + eps with
      // D a host constant (cos D, sin D tabulated) and eps = (a_t - a_b) - D the
      // MEASURED residual (|eps| ~ ulp(a)), expanded to second order.
      const int base = facs[k].aux;  // physical slot of the base row's cosine
      const double bshift = facs[k].p[0], D = facs[k].p[1], cD = facs[k].p[2], sD = facs[k].p[3];
      const double a_t = mul(a0, t);
      const double a_b = mul(a0, sub(x, bshift));
      const double eps = sub(sub(a_t, a_b), D);
      const double cb = vals.phys(base), sb = vals.phys(base + 1);
      const double C = fma(cb, cD, -(sb * sD));
      const double S = fma(sb, cD, cb * sD);
      vals.setp(dest, fma(-0.5 * eps * eps, C, fma(-eps, S, C)));
    } else if (op == OP_SINCOS) {
      // one range reduction serves every COS factor of this frequency
      const SinCos sc = sincos_cw(mul(a0, t));
      vals.setp(dest, sc.c);
      vals.setp(dest + 1, sc.s);
    } else if (op == OP_COS) {
      vals.setp(dest, cos_cw(mul(a0, t)));
    } else if (op == OP_LINEAR) {
      vals.setp(dest, t);
    } else if (op == OP_GAUSSIAN) {
      vals.setp(dest, f_gaussian(t, a0));
    } else if (op == OP_ERF) {
      vals.setp(dest, erf(dvd(t, a0)));
    } else if (op != OP_NOP) {
      vals.setp(dest, eval_factor(fac_args(facs[k]), x, args));
    }
  }
}

// product of the referenced factor powers of an ABI term (general path); gfac = the
// segment's first row in the GLOBAL factor table
template <typename Slots>
__device__ __noinline__ double term_product(const IrGlobal& g, int gfac, const WfmTerm& tm, double x,
                                            const Slots& vals) {
  double prod = 1.0;
#pragma unroll 1
  for (int r = 0; r < tm.n_ref; ++r) {
    const WfmRef ref = g.refs[tm.ref_begin + r];
    double v = (ref.slot < kMaxSlots) ? vals.phys(ref.slot + 1) : eval_factor(fac_args(g.dfacs[gfac + ref.slot]), x, g.args);
    if (ref.kind == WFM_POW_INT) v = pow_small_int(v, (int)ref.expo);
    else if (ref.kind == WFM_POW_GEN) v = pow(v, ref.expo);
    prod = mul(prod, v);  // the reference's product starts from 1; 1 * v is exact
  }
  return prod;
}

static __device__ __noinline__ double clip_value(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// Evaluate one segment's program at one abscissa (real-valued channels; compact terms).
// facs / cterms point at the segment's first rows (shared or global memory);
// gfac / gterm are the same rows' indices in the global tables.
// Order of operations = the reference's (_waveform.pyx:134-152, waveform.py:690-692):
// every term is amp * (((1 * f1) * f2) * ...), a member's terms are summed left to
// right starting from 0, the member sums are added to the accumulator (offset).
template <typename Slots>
__device__ __forceinline__ double eval_segment_real(const DFactor* __restrict__ facs, const CTerm* __restrict__ cterms,
                                                    int nf, int nt, const IrGlobal& g, int gfac, int gterm,
                                                    const WaveEval& w, double x, Slots& vals) {
  double total = w.offset;
  if (nt == 0) return total;  // zero segment: untouched by clip (calc_parts skips it)
  eval_factors(facs, nf, x, g.args, vals);  // global rows: nf <= kMaxSlots (the caller clamps)
  double grp = 0.0;
#pragma unroll 1
  for (int it = 0; it < nt; ++it) {
    const double amp = cterms[it].amp;
    const uint32_t pk = (uint32_t)cterms[it].packed;
    double prod;
    if (pk & kCTermExt) {
      prod = term_product(g, gfac, g.terms[gterm + it], x, vals);
    } else {
      prod = mul(mul(vals.phys((pk >> 8) & 0xffu), vals.phys((pk >> 16) & 0xffu)), vals.phys(pk >> 24));
    }
    grp = add(grp, mul(amp, prod));
    if (pk & kCTermGroupEnd) {
      total = add(total, grp);
      grp = 0.0;
    }
  }
  if (w.flags & WFM_WAVE_CLIP) total = clip_value(total, w.clip_lo, w.clip_hi);
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
  eval_factors(g.dfacs + p0.fac, min(p1.fac - p0.fac, kMaxSlots), x, g.args, vals);
  double g_re = 0.0, g_im = 0.0;
#pragma unroll 1
  for (int it = 0; it < nt; ++it) {
    const WfmTerm tm = g.terms[p0.term + it];
    const double prod = term_product(g, p0.fac, tm, x, vals);
    g_re = add(g_re, mul(tm.amp_re, prod));
    g_im = add(g_im, mul(tm.amp_im, prod));
    if (tm.flags & WFM_TERM_GROUP_END) {
      out_re = add(out_re, g_re);
      out_im = add(out_im, g_im);
      g_re = g_im = 0.0;
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
    // constant segment: offset + sum over stack members of (0 + sum of their constant terms)
    double grp = 0.0;
    for (int t = p0.term; t < p1.term; ++t) {
      const WfmTerm tm = P.terms[t];
      grp = add(grp, tm.n_ref > 0 ? CUDART_NAN : tm.amp_re);  // a term with factors would make the segment active
      if (tm.flags & WFM_TERM_GROUP_END) {
        val = add(val, grp);
        grp = 0.0;
      }
    }
    if (w.flags & WFM_WAVE_CLIP) val = fmin(fmax(val, w.clip_lo), w.clip_hi);
  }
  seg_val[s] = val;
}

// one thread per factor row: the 64-byte device row
__global__ void prepare_factors_kernel(DevProgram P, DFactor* __restrict__ dfacs, int64_t n_facs, int64_t n_segs) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_facs) return;
  const WfmFactor f = P.facs[k];
  int op = OP_GENERIC;
  switch (f.func) {
    case WFM_COS_ROT: op = OP_ROT; break;
    case WFM_COS_SINCOS: op = OP_SINCOS; break;
    case WFM_NOP: op = OP_NOP; break;
    case WFM_COS: op = OP_COS; break;
    case WFM_LINEAR: op = OP_LINEAR; break;
    case WFM_GAUSSIAN: op = OP_GAUSSIAN; break;
    case WFM_ERF: op = OP_ERF; break;
    default: break;
  }
  // destination (physical) value slot = row within its segment + 1; found from the
  // segment table by binary search over seg_ptr (rows of one segment are contiguous)
  int lo = 0, hi = (int)n_segs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if ((int64_t)P.seg_ptr[mid].fac <= k) lo = mid; else hi = mid - 1;
  }
  while (lo + 1 < (int)n_segs && P.seg_ptr[lo + 1].fac <= k) ++lo;  // skip factor-less segments sharing the offset
  const int row = (int)(k - P.seg_ptr[lo].fac);
  if (row >= kMaxSlots) op = OP_NOP;  // beyond the value cache: evaluated on demand by extended terms
  const int dest = row < kMaxSlots ? row + 1 : 0;
  DFactor d;
  d.func = f.func | (op << 16) | (dest << 24);
  d.aux = f.arg_off;
  d.shift = f.shift;
  d.a0 = f.a0;
  d.a1 = f.a1;
  d.p[0] = d.p[1] = d.p[2] = d.p[3] = 0.0;
  if (f.func == WFM_COS_ROT) {
    const double* __restrict__ p = P.args + f.arg_off;  // [base_slot, base_shift, D, cos D, sin D]
    d.aux = (int)p[0] + 1;  // physical slot of the base row's cosine
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
  bool ext = tm.n_ref > 3;
  for (int r = 0; r < tm.n_ref && !ext; ++r) {
    const WfmRef rf = P.refs[tm.ref_begin + r];
    if (rf.kind != WFM_POW_ONE || rf.slot >= kMaxSlots) ext = true;
    else packed |= (uint64_t)(uint32_t)(rf.slot + 1) << (8 + 8 * r);  // physical slot; 0 = the constant 1.0
  }
  uint32_t flags = (tm.flags & WFM_TERM_GROUP_END) ? kCTermGroupEnd : 0u;
  if (ext) {
    flags |= kCTermExt;
    packed = 0;
  }
  packed |= flags;
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

// what a tile's packet holds (shared by the measuring and the filling pass)
struct TileLayout {
  int n_arows, n_patch, n_fac, n_term, n_active;
  bool cold;
  __device__ __forceinline__ int bytes() const {
    if (cold) return (int)sizeof(PacketHeader);
    return (int)sizeof(PacketHeader) + (n_arows ? (n_arows + 1) * (int)sizeof(ARow) : 0) + n_patch * (int)sizeof(PatchRow) +
           n_fac * (int)sizeof(DFactor) + n_term * (int)sizeof(CTerm);
  }
};

// tile-relative [a, b) of segment k (channel-relative row) clipped to the tile
__device__ __forceinline__ void seg_span(const int32_t* __restrict__ st, int n_seg, int64_t n, int k, int64_t j0, int cnt,
                                         int& a, int& b) {
  const int64_t lo = st[k], hi = (k + 1 < n_seg) ? (int64_t)st[k + 1] : n;
  a = (int)(max(lo, j0) - j0);
  b = (int)(min(hi, j0 + cnt) - j0);
}

// rows of a segment the packet carries: no NOP placeholders, nothing beyond the value cache
__device__ __forceinline__ int packet_rows(const DFactor* __restrict__ dfacs, WfmSegPtr p0, WfmSegPtr p1) {
  int n = 0;
  for (int r = p0.fac; r < p1.fac; ++r) n += (((uint32_t)dfacs[r].func >> 16) & 0xff) != OP_NOP;
  return n;
}

__device__ TileLayout measure_tile(const DevProgram& P, const TileDesc& td, const WfmWave& w) {
  TileLayout L{0, 0, 0, 0, 0, false};
  const int32_t* __restrict__ st = P.seg_start + w.seg_begin;
  const int k0 = td.seg0 - w.seg_begin;
  for (int k = k0; k < k0 + td.nb; ++k) {
    int a, b;
    seg_span(st, w.n_seg, w.n, k, td.j0, td.cnt, a, b);
    if (b <= a) continue;
    const WfmSegPtr p0 = P.seg_ptr[w.seg_begin + k], p1 = P.seg_ptr[w.seg_begin + k + 1];
    if (p1.fac > p0.fac) {
      L.n_arows += 1;
      L.n_fac += packet_rows(P.dfacs, p0, p1);
      L.n_term += p1.term - p0.term;
      L.n_active += b - a;
    } else if (__double_as_longlong(P.seg_val[w.seg_begin + k]) != __double_as_longlong(w.offset)) {
      L.n_patch += 1;
    }
  }
  L.cold = L.bytes() > P.pkt_cap;
  return L;
}

// one thread per tile: the segment rows it spans, its slice of the tables, its packet size
__global__ void prepare_tiles_kernel(DevProgram P, TileDesc* __restrict__ tiles, int64_t n_tiles,
                                     uint32_t* __restrict__ pkt_size) {
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
  pkt_size[t] = (uint32_t)(measure_tile(P, td, w).bytes() / 16);
}

// ---- exclusive scan of the packet sizes (three small kernels) -----------------------------
constexpr int kScanBlock = 256, kScanItems = 16, kScanTile = kScanBlock * kScanItems;  // 4096 per block

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[kScanBlock / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  uint32_t before = 0, all = 0;
  for (int k = 0; k < kScanBlock / 32; ++k) {
    if (k < wid) before += warp_sums[k];
    all += warp_sums[k];
  }
  __syncthreads();
  *total = all;
  return before + incl - v;
}

__global__ void __launch_bounds__(kScanBlock) scan_reduce_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ sums,
                                                                 int64_t n) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v = 0;
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) v += in[base + i];
  uint32_t total;
  block_exclusive_scan(v, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of the block sums in place; sums[n_blocks] = grand total
__global__ void __launch_bounds__(kScanBlock) scan_sums_kernel(uint32_t* __restrict__ sums, int64_t n_blocks) {
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < n_blocks; b0 += kScanBlock) {
    const int64_t i = b0 + threadIdx.x;
    const uint32_t v = i < n_blocks ? sums[i] : 0;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, &total);
    if (i < n_blocks) sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[n_blocks] = carry;
}

__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ sums,
                                                                uint32_t* __restrict__ out, int64_t n, int64_t n_blocks) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t item[kScanItems];
  uint32_t v = 0;
  for (int i = 0; i < kScanItems; ++i) {
    item[i] = base + i < n ? in[base + i] : 0;
    v += item[i];
  }
  uint32_t total;
  uint32_t run = sums[blockIdx.x] + block_exclusive_scan(v, &total);
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += item[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[n_blocks];
}

// ---- pass 2: one warp per tile writes its packet --------------------------------------------
__global__ void __launch_bounds__(256) fill_packets_kernel(DevProgram P, const TileDesc* __restrict__ tiles,
                                                           int64_t n_tiles, unsigned char* __restrict__ packets) {
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= n_tiles) return;
  const TileDesc td = tiles[t];
  const WfmWave w = P.waves[td.wave];
  const TileLayout L = measure_tile(P, td, w);  // every lane, identical result
  unsigned char* pk = packets + (size_t)P.pkt_off[t] * 16;
  ARow* arows = reinterpret_cast<ARow*>(pk + sizeof(PacketHeader));
  PatchRow* patches = reinterpret_cast<PatchRow*>(arows + (L.n_arows ? L.n_arows + 1 : 0));
  DFactor* facs = reinterpret_cast<DFactor*>(patches + L.n_patch);
  CTerm* cterms = reinterpret_cast<CTerm*>(facs + L.n_fac);
  if (lane == 0) {
    PacketHeader h;
    h.out0 = td.out0;
    h.j0 = td.j0;
    h.base = w.offset;
    h.t0 = w.t0;
    h.delta = w.delta;
    h.wave = td.wave;
    h.flags = w.flags | (L.cold ? kPacketCold : 0u);
    h.cnt = (uint16_t)td.cnt;
    h.n_arows = L.cold ? 0 : (uint16_t)L.n_arows;
    h.n_patch = L.cold ? 0 : (uint16_t)L.n_patch;
    h.n_active = L.cold ? 0 : (uint16_t)L.n_active;
    h.n_fac = L.cold ? 0 : (uint16_t)L.n_fac;
    h.n_term = L.cold ? 0 : (uint16_t)L.n_term;
    h.reserved = 0;
    *reinterpret_cast<PacketHeader*>(pk) = h;
  }
  if (L.cold) return;
  // rows: one sequential walk over the tile's segments (lane 0), table rows copied by all lanes
  const int32_t* __restrict__ st = P.seg_start + w.seg_begin;
  const int k0 = td.seg0 - w.seg_begin;
  int ia = 0, ip = 0, fac_rel = 0, term_rel = 0, first = 0;
  for (int k = k0; k < k0 + td.nb; ++k) {
    int a, b;
    seg_span(st, w.n_seg, w.n, k, td.j0, td.cnt, a, b);
    if (b <= a) continue;
    const WfmSegPtr p0 = P.seg_ptr[w.seg_begin + k], p1 = P.seg_ptr[w.seg_begin + k + 1];
    if (p1.fac > p0.fac) {
      if (lane == 0)
        arows[ia] = ARow{(uint16_t)a, (uint16_t)first, (uint16_t)fac_rel, (uint16_t)term_rel, p0.fac, p0.term};
      // factor rows without the NOP placeholders (order kept), 16 bytes per lane and step
      int kept = 0;
      for (int r = p0.fac; r < p1.fac; ++r) {
        if ((((uint32_t)P.dfacs[r].func >> 16) & 0xff) == OP_NOP) continue;
        if (lane < 4)
          reinterpret_cast<uint4*>(facs + fac_rel + kept)[lane] = reinterpret_cast<const uint4*>(P.dfacs + r)[lane];
        ++kept;
      }
      for (int q = lane; q < p1.term - p0.term; q += 32)
        reinterpret_cast<uint4*>(cterms + term_rel)[q] = reinterpret_cast<const uint4*>(P.cterms + p0.term)[q];
      ia += 1;
      fac_rel += kept;
      term_rel += p1.term - p0.term;
      first += b - a;
    } else {
      const double v = P.seg_val[w.seg_begin + k];
      if (__double_as_longlong(v) != __double_as_longlong(w.offset)) {
        if (lane == 0) patches[ip] = PatchRow{(uint16_t)a, (uint16_t)b, 0u, v};
        ip += 1;
      }
    }
  }
  if (lane == 0 && L.n_arows)
    arows[ia] = ARow{(uint16_t)td.cnt, (uint16_t)first, (uint16_t)fac_rel, (uint16_t)term_rel, 0, 0};  // sentinel
}

// ---- the sampling kernel ------------------------------------------------------------------
// [a, b) of the warp's shared tile <- val
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

// per-warp shared-memory slice (dynamic shared memory; all sub-arrays 16-byte aligned):
//   [out: tile_samples x OutT][slots: n_slots x 32 x f64][packet buffer 0][packet buffer 1][2 mbarriers]
__host__ __device__ inline size_t warp_slice_bytes(int tile_samples, int n_slots, int pkt_cap, size_t esz) {
  size_t b = (size_t)tile_samples * esz + (size_t)n_slots * 32 * 8 + 2 * (size_t)pkt_cap + 16;
  return (b + 127) & ~(size_t)127;
}

// the cold path of a tile (its packet would not fit the warp's buffers): per-sample
// search, tables in global memory, direct stores
template <typename OutT, bool kAccumulate>
__device__ __noinline__ void sample_tile_cold(const DevProgram& P, const TileDesc& td, OutT* __restrict__ dst, int lane) {
  const WfmWave w = P.waves[td.wave];
  const WaveEval we{w.offset, w.clip_lo, w.clip_hi, w.flags};
  const IrGlobal g{P.dfacs, P.terms, P.refs, P.args};
  const int32_t* __restrict__ st = P.seg_start + td.seg0;
  const WfmSegPtr* __restrict__ gp = P.seg_ptr + td.seg0;
  for (int jj = lane; jj < td.cnt; jj += 32) {
    int lo = 0, hi = td.nb - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if ((int64_t)st[mid] <= td.j0 + jj) lo = mid; else hi = mid - 1;
    }
    const WfmSegPtr p0 = gp[lo], p1 = gp[lo + 1];
    LocalSlots vals;
    const double re = eval_segment_real(P.dfacs + p0.fac, P.cterms + p0.term, min(p1.fac - p0.fac, kMaxSlots),
                                        p1.term - p0.term, g, p0.fac, p0.term, we, abscissa(w, P.x, td.j0 + jj), vals);
    dst[jj] = kAccumulate ? (OutT)add((double)dst[jj], re) : (OutT)re;
  }
}

extern __shared__ __align__(128) unsigned char k1_smem[];

template <typename OutT, bool kAccumulate>
__global__ void __launch_bounds__(kThreads, WFM_K1_MIN_BLOCKS)
    sample_kernel(DevProgram P, const TileDesc* __restrict__ tiles, int tile_begin, int tile_end, OutT* __restrict__ out) {
  constexpr int V = OutVec<OutT>::N;
  const int lane = threadIdx.x & 31;
  const int warp_in_cta = threadIdx.x >> 5;
  // this warp's private slice
  unsigned char* slice = k1_smem + (size_t)warp_in_cta * warp_slice_bytes(P.tile_samples, P.n_slots, P.pkt_cap, sizeof(OutT));
  OutT* s_out = reinterpret_cast<OutT*>(slice);
  double* s_slots = reinterpret_cast<double*>(slice + (size_t)P.tile_samples * sizeof(OutT));
  unsigned char* s_pkt = reinterpret_cast<unsigned char*>(s_slots + (size_t)P.n_slots * 32);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_pkt + 2 * (size_t)P.pkt_cap);

  const IrGlobal g{P.dfacs, P.terms, P.refs, P.args};
  SmemSlots vals{s_slots + lane};
  const int n_warps = gridDim.x * kWarpsPerCta;
  int t = tile_begin + blockIdx.x * kWarpsPerCta + warp_in_cta;
  if (t >= tile_end) return;

  if (lane == 0) {
    mbar_init(s_bar, 1);
    mbar_init(s_bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  s_slots[lane] = 1.0;  // physical slot 0: the unit a missing term reference multiplies by
  __syncwarp();

  // packet of the first tile -> buffer 0; offsets of the second tile -> registers
  uint32_t off_next = 0, end_next = 0;  // packet of tile t + n_warps, 16-byte units
  {
    const uint32_t o0 = P.pkt_off[t], o1 = P.pkt_off[t + 1];
    if (lane == 0) {
      mbar_expect_tx(s_bar, (o1 - o0) * 16u);
      bulk_g2s(s_pkt, P.packets + (size_t)o0 * 16, (o1 - o0) * 16u, s_bar);
    }
    if (t + n_warps < tile_end) {
      off_next = P.pkt_off[t + n_warps];
      end_next = P.pkt_off[t + n_warps + 1];
    }
  }
  uint32_t phases = 0;         // bit b: parity to wait for on buffer b
  int buf = 0;
  bool store_pending = false;  // lane 0: a bulk store may still be reading s_out

#pragma unroll 1
  for (; t < tile_end; t += n_warps) {
    const unsigned char* pk = s_pkt + (size_t)buf * P.pkt_cap;
    // prefetch: the next tile's packet into the other buffer (its previous tile is done:
    // every lane passed the __syncwarp that ends an iteration), the offsets of the tile after
    if (t + n_warps < tile_end) {
      if (lane == 0) {
        mbar_expect_tx(s_bar + (buf ^ 1), (end_next - off_next) * 16u);
        bulk_g2s(s_pkt + (size_t)(buf ^ 1) * P.pkt_cap, P.packets + (size_t)off_next * 16, (end_next - off_next) * 16u,
                 s_bar + (buf ^ 1));
      }
      if (t + 2 * n_warps < tile_end) {
        off_next = P.pkt_off[t + 2 * n_warps];
        end_next = P.pkt_off[t + 2 * n_warps + 1];
      }
    }
    mbar_wait(s_bar + buf, (phases >> buf) & 1u);
    phases ^= 1u << buf;

    const PacketHeader* __restrict__ h = reinterpret_cast<const PacketHeader*>(pk);
    const int cnt = h->cnt;
    const uint32_t flags = h->flags;
    const int n_arows = h->n_arows, n_patch = h->n_patch, n_active = h->n_active;
    const double base = h->base;
    OutT* __restrict__ dst = out + h->out0;

    if (flags & kPacketCold) {
      sample_tile_cold<OutT, kAccumulate>(P, tiles[t], dst, lane);
      __syncwarp();
      buf ^= 1;
      continue;
    }

    // the previous tile's bulk store must have finished READING the tile buffer
    if (lane == 0 && store_pending) bulk_wait_read_all();
    __syncwarp();

    // base fill: the whole tile <- the zero-segment value; flat segments with another
    // value and the active samples overwrite it below
    {
      const int n_fill = (cnt + V - 1) & ~(V - 1);  // the tile buffer is a multiple of V
      int p = lane * V;
#pragma unroll 1
      for (; p + 3 * 32 * V < n_fill; p += 4 * 32 * V) {
        fill_vec(s_out + p, base);
        fill_vec(s_out + p + 32 * V, base);
        fill_vec(s_out + p + 2 * 32 * V, base);
        fill_vec(s_out + p + 3 * 32 * V, base);
      }
#pragma unroll 1
      for (; p < n_fill; p += 32 * V) fill_vec(s_out + p, base);
    }
    __syncwarp();

    const ARow* __restrict__ arows = reinterpret_cast<const ARow*>(pk + sizeof(PacketHeader));
    const PatchRow* __restrict__ patches = reinterpret_cast<const PatchRow*>(arows + (n_arows ? n_arows + 1 : 0));
    // ---- flat segments with their own value ----------------------------------------------
    for (int i = 0; i < n_patch; ++i) fill_run(s_out, (int)patches[i].a, (int)patches[i].b, patches[i].val, lane);

    // ---- the tile's ACTIVE samples, dealt round-robin to the lanes ------------------------
    if (n_active > 0) {
      const DFactor* __restrict__ sf = reinterpret_cast<const DFactor*>(patches + n_patch);
      const CTerm* __restrict__ sc = reinterpret_cast<const CTerm*>(sf + h->n_fac);
      WaveEval we{base, 0.0, 0.0, flags};
      if (flags & WFM_WAVE_CLIP) {
        we.clip_lo = P.waves[h->wave].clip_lo;
        we.clip_hi = P.waves[h->wave].clip_hi;
      }
      const double t0 = h->t0, delta = h->delta;
      const int64_t j0 = h->j0;
      const bool plain_grid = !(flags & (WFM_WAVE_EXPLICIT_X | WFM_WAVE_LAST_OVERRIDE | WFM_WAVE_PRESHIFT));
      int lo = 0;
#pragma unroll 1
      for (int i = lane; i < n_active; i += 32) {
        while ((int)arows[lo + 1].first <= i) ++lo;  // the active segment holding sample i (i grows monotonically)
        const ARow r0 = arows[lo];
        const int fac_end = arows[lo + 1].fac_rel, term_end = arows[lo + 1].term_rel;
        const int jj = (int)r0.start + (i - (int)r0.first);
        const double x = plain_grid ? add(t0, mul((double)(j0 + jj), delta)) : abscissa(P.waves[h->wave], P.x, j0 + jj);
        s_out[jj] = (OutT)eval_segment_real(sf + r0.fac_rel, sc + r0.term_rel, fac_end - (int)r0.fac_rel,
                                            term_end - (int)r0.term_rel, g, r0.gfac, r0.gterm, we, x, vals);
      }
    }

    // ---- store ---------------------------------------------------------------------------
    if (kAccumulate) {
      // out += tile (Waveform.__call__(..., accumulate=True)): read-modify-write epilogue
      __syncwarp();
      for (int p = lane; p < cnt; p += 32) dst[p] = (OutT)add((double)dst[p], (double)s_out[p]);
    } else {
      // the whole tile as one TMA bulk copy
      fence_proxy_async_smem();
      __syncwarp();
      const int n_bulk = cnt & ~(V - 1);  // 16-byte multiple; the ragged tail goes out as scalars
      if (lane == 0 && n_bulk > 0) {
        bulk_s2g_evict_first(dst, s_out, (uint32_t)n_bulk * sizeof(OutT));
        store_pending = true;
      }
      if (n_bulk + lane < cnt) dst[n_bulk + lane] = s_out[n_bulk + lane];
    }
    __syncwarp();  // all reads of the packet and of the tail of s_out are done
    buf ^= 1;
  }
  if (lane == 0 && store_pending) bulk_wait_read_all();  // shared memory must outlive the copy's reads
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
                           DFactor* dfacs, CTerm* cterms, TileDesc* tiles, uint32_t* pkt_size, cudaStream_t stream) {
  const int threads = 128;
  auto blocks = [&](int64_t items) { return (unsigned)((items + threads - 1) / threads); };
  if (n.n_segs > 0) prepare_segments_kernel<<<blocks(n.n_segs), threads, 0, stream>>>(P, seg_start, seg_val, n.n_segs);
  if (n.n_facs > 0) prepare_factors_kernel<<<blocks(n.n_facs), threads, 0, stream>>>(P, dfacs, n.n_facs, n.n_segs);
  if (n.n_terms > 0) prepare_terms_kernel<<<blocks(n.n_terms), threads, 0, stream>>>(P, cterms, n.n_terms);
  if (n.n_tiles > 0) prepare_tiles_kernel<<<blocks(n.n_tiles), threads, 0, stream>>>(P, tiles, n.n_tiles, pkt_size);
  return cudaGetLastError();
}

cudaError_t launch_scan(const uint32_t* pkt_size, uint32_t* pkt_off, uint32_t* scratch, int64_t n, cudaStream_t stream) {
  const int64_t n_blocks = (n + kScanTile - 1) / kScanTile;
  if (n_blocks == 0) return cudaMemsetAsync(pkt_off, 0, sizeof(uint32_t), stream);
  scan_reduce_kernel<<<(unsigned)n_blocks, kScanBlock, 0, stream>>>(pkt_size, scratch, n);
  scan_sums_kernel<<<1, kScanBlock, 0, stream>>>(scratch, n_blocks);
  scan_apply_kernel<<<(unsigned)n_blocks, kScanBlock, 0, stream>>>(pkt_size, scratch, pkt_off, n, n_blocks);
  return cudaGetLastError();
}

cudaError_t launch_fill_packets(const DevProgram& P, const TileDesc* tiles, int64_t n_tiles, unsigned char* packets,
                                cudaStream_t stream) {
  if (n_tiles == 0) return cudaSuccess;
  fill_packets_kernel<<<(unsigned)((n_tiles + 7) / 8), 256, 0, stream>>>(P, tiles, n_tiles, packets);
  return cudaGetLastError();
}

int warp_fixed_bytes(int n_slots) { return n_slots * 32 * 8 + 16 + 128; }

size_t sample_smem_bytes(const DevProgram& P, int dtype) {
  return kWarpsPerCta * warp_slice_bytes(P.tile_samples, P.n_slots, P.pkt_cap, dtype == WFM_F32 ? 4 : 8);
}

template <typename OutT, bool kAcc>
static cudaError_t launch_persistent(const DevProgram& P, const TileDesc* tiles, int64_t tile_begin, int64_t n_tiles,
                                     int dtype, void* out, cudaStream_t stream) {
  auto k = sample_kernel<OutT, kAcc>;
  const size_t smem = sample_smem_bytes(P, dtype);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, per_sm = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, smem)) != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidConfiguration;
  const int64_t want = (n_tiles + kWarpsPerCta - 1) / kWarpsPerCta;
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)sms * per_sm);
  k<<<grid, kThreads, smem, stream>>>(P, tiles, (int)tile_begin, (int)(tile_begin + n_tiles), (OutT*)out);
  return cudaGetLastError();
}

cudaError_t launch_sample(const DevProgram& P, const TileDesc* tiles, int64_t tile_begin, int64_t n_tiles, int dtype,
                          int accumulate, void* out, cudaStream_t stream) {
  if (n_tiles == 0) return cudaSuccess;
  if (tile_begin + n_tiles > INT32_MAX) return cudaErrorInvalidValue;
  if (dtype == WFM_F64)
    return accumulate ? launch_persistent<double, true>(P, tiles, tile_begin, n_tiles, dtype, out, stream)
                      : launch_persistent<double, false>(P, tiles, tile_begin, n_tiles, dtype, out, stream);
  if (dtype == WFM_F32)
    return accumulate ? launch_persistent<float, true>(P, tiles, tile_begin, n_tiles, dtype, out, stream)
                      : launch_persistent<float, false>(P, tiles, tile_begin, n_tiles, dtype, out, stream);
  dim3 grid((unsigned)n_tiles), block(kThreads);
  if (accumulate) sample_kernel_c128<true><<<grid, block, 0, stream>>>(P, tiles + tile_begin, (double2*)out);
  else sample_kernel_c128<false><<<grid, block, 0, stream>>>(P, tiles + tile_begin, (double2*)out);
  return cudaGetLastError();
}

}  // namespace wfm

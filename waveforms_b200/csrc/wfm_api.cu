// wfm_api.cu — extern "C" boundary of libwfmb200.so (see include/wfm_b200.h).
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "wfm_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define WFM_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return fail(WFM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <typename T>
cudaError_t upload(const T* host, int64_t n, const T** dev) {
  *dev = nullptr;
  // always allocate at least one element so kernels never see a null table
  size_t bytes = sizeof(T) * (size_t)std::max<int64_t>(n, 1);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return e;
  if (n > 0) {
    e = cudaMemcpy(p, host, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p); return e; }
  } else {
    cudaMemset(p, 0, bytes);
  }
  *dev = static_cast<const T*>(p);
  return cudaSuccess;
}

bool known_func(int f) { return (f >= WFM_LINEAR && f <= WFM_DRAG_SINX) || (f >= WFM_COS_SINCOS && f <= WFM_COS_ROT); }

}  // namespace

struct WfmProgram {
  int device = 0;
  wfm::DevProgram dev{};
  std::vector<WfmWave> waves;  // host copy (tile lists, output sizing)
  wfm::TileDesc* d_tiles = nullptr;
  std::vector<int64_t> tile_prefix;  // tiles before channel w
  int64_t n_tiles = 0;
  int64_t total_samples = 0;  // extent of the output buffer in samples
  int64_t launches = 0;
  void* d_stage = nullptr;  // device staging buffer for wfm_sample_host
  size_t stage_bytes = 0;
  bool any_complex = false;

  ~WfmProgram() {
    DeviceGuard g(device);
    cudaFree((void*)dev.waves);
    cudaFree((void*)dev.seg_bound);
    cudaFree((void*)dev.seg_ptr);
    cudaFree((void*)dev.facs);
    cudaFree((void*)dev.terms);
    cudaFree((void*)dev.cterms);
    cudaFree((void*)dev.seg_plan);
    cudaFree((void*)dev.row_slot);
    cudaFree((void*)dev.refs);
    cudaFree((void*)dev.args);
    cudaFree((void*)dev.x);
    cudaFree((void*)dev.seg_start);
    cudaFree((void*)dev.seg_val);
    cudaFree((void*)dev.seg_wave);
    cudaFree((void*)dev.pkt_off);
    cudaFree((void*)dev.packets);
    cudaFree(d_tiles);
    cudaFree(d_stage);
  }
};

static int validate(const WfmProgramDesc* d) {
  if (!d) return fail(WFM_EINVAL, "null program descriptor");
  if (d->n_waves < 0 || d->n_segs < 0 || d->n_facs < 0 || d->n_terms < 0 || d->n_refs < 0 || d->n_args < 0 || d->n_x < 0)
    return fail(WFM_EINVAL, "negative table size");
  if (d->n_segs > INT32_MAX - 1 || d->n_facs > INT32_MAX || d->n_terms > INT32_MAX || d->n_refs > INT32_MAX ||
      d->n_args > INT32_MAX)
    return fail(WFM_EINVAL, "table too large for 32-bit indices");
  if ((d->n_waves && !d->waves) || (d->n_segs && (!d->seg_bound || !d->seg_ptr)) || (d->n_facs && !d->facs) ||
      (d->n_terms && !d->terms) || (d->n_refs && !d->refs) || (d->n_args && !d->args) || (d->n_x && !d->x))
    return fail(WFM_EINVAL, "null table pointer with non-zero size");
  // segment table
  for (int64_t s = 0; s < d->n_segs; ++s) {
    const WfmSegPtr a = d->seg_ptr[s], b = d->seg_ptr[s + 1];
    if (a.fac < 0 || a.term < 0 || b.fac < a.fac || b.term < a.term)
      return fail(WFM_EINVAL, "segment %lld: pointer table not monotone", (long long)s);
  }
  if (d->n_segs) {
    const WfmSegPtr e = d->seg_ptr[d->n_segs];
    if (e.fac != d->n_facs || e.term != d->n_terms)
      return fail(WFM_EINVAL, "segment pointer table does not close (%d/%lld factors, %d/%lld terms)", e.fac,
                  (long long)d->n_facs, e.term, (long long)d->n_terms);
  }
  for (int64_t k = 0; k < d->n_facs; ++k) {
    const WfmFactor& f = d->facs[k];
    if (!known_func(f.func)) return fail(WFM_EUNSUPPORTED, "factor %lld: unknown basis id %d", (long long)k, f.func);
    if (f.arg_off < 0 || f.arg_off > d->n_args)
      return fail(WFM_EINVAL, "factor %lld: argument offset out of range", (long long)k);
    if (f.func == WFM_INTERP) {
      if (f.arg_off + 2 > d->n_args) return fail(WFM_EINVAL, "factor %lld: INTERP header out of range", (long long)k);
      double n = d->args[f.arg_off];
      if (!(n >= 1) || f.arg_off + 2 + (int64_t)n > d->n_args)
        return fail(WFM_EINVAL, "factor %lld: INTERP table out of range", (long long)k);
    }
  }
  // terms and refs, segment by segment (slot indices are segment-relative)
  for (int64_t s = 0; s < d->n_segs; ++s) {
    const WfmSegPtr a = d->seg_ptr[s], b = d->seg_ptr[s + 1];
    const int nf = b.fac - a.fac;
    for (int k = 0; k < nf; ++k) {
      const WfmFactor& f = d->facs[a.fac + k];
      if (f.func == WFM_COS_SINCOS) {
        if (k + 1 >= nf || d->facs[a.fac + k + 1].func != WFM_NOP)
          return fail(WFM_EINVAL, "segment %lld: COS_SINCOS row %d needs a NOP row after it", (long long)s, k);
      } else if (f.func == WFM_NOP) {
        if (k == 0 || d->facs[a.fac + k - 1].func != WFM_COS_SINCOS)
          return fail(WFM_EINVAL, "segment %lld: stray NOP row %d", (long long)s, k);
      } else if (f.func == WFM_COS_ROT) {
        if (f.arg_off + 5 > d->n_args) return fail(WFM_EINVAL, "segment %lld: COS_ROT row %d: pool out of range", (long long)s, k);
        const double bs = d->args[f.arg_off];
        const int base = (int)bs;
        if (!(bs >= 0) || base >= k || d->facs[a.fac + base].func != WFM_COS_SINCOS ||
            d->facs[a.fac + base].a0 != f.a0)
          return fail(WFM_EINVAL, "segment %lld: COS_ROT row %d: bad base row", (long long)s, k);
      }
    }
    for (int t = a.term; t < b.term; ++t) {
      const WfmTerm& tm = d->terms[t];
      if (tm.n_ref < 0 || tm.ref_begin < 0 || (int64_t)tm.ref_begin + tm.n_ref > d->n_refs)
        return fail(WFM_EINVAL, "term %d: reference range out of bounds", t);
      for (int r = tm.ref_begin; r < tm.ref_begin + tm.n_ref; ++r) {
        const WfmRef& rf = d->refs[r];
        if (rf.slot < 0 || rf.slot >= nf) return fail(WFM_EINVAL, "ref %d: slot %d outside segment (%d factors)", r, rf.slot, nf);
        if (d->facs[a.fac + rf.slot].func == WFM_NOP) return fail(WFM_EINVAL, "ref %d: refers to a NOP row", r);
        if (rf.kind < WFM_POW_ONE || rf.kind > WFM_POW_GEN) return fail(WFM_EINVAL, "ref %d: bad exponent kind", r);
      }
    }
    if (b.term > a.term && !(d->terms[b.term - 1].flags & WFM_TERM_GROUP_END))
      return fail(WFM_EINVAL, "segment %lld: last term does not close its group", (long long)s);
  }
  // the sampling kernel stages a tile's references as ONE contiguous slice: they
  // must be packed in term order
  for (int64_t t = 0; t + 1 < d->n_terms; ++t)
    if (d->terms[t + 1].ref_begin != d->terms[t].ref_begin + d->terms[t].n_ref)
      return fail(WFM_EINVAL, "term %lld: references must be packed in term order", (long long)(t + 1));
  for (int64_t w = 0; w < d->n_waves; ++w) {
    const WfmWave& wv = d->waves[w];
    if (wv.n < 0 || wv.out_off < 0) return fail(WFM_EINVAL, "channel %lld: negative extent", (long long)w);
    if (wv.n >= INT32_MAX) return fail(WFM_EINVAL, "channel %lld: more than 2^31-2 samples", (long long)w);
    if (wv.n_seg < 1 || wv.seg_begin < 0 || (int64_t)wv.seg_begin + wv.n_seg > d->n_segs)
      return fail(WFM_EINVAL, "channel %lld: segment range out of bounds", (long long)w);
    if (!(d->seg_bound[wv.seg_begin + wv.n_seg - 1] == INFINITY))
      return fail(WFM_EINVAL, "channel %lld: last bound must be +inf", (long long)w);
    if ((wv.flags & WFM_WAVE_EXPLICIT_X) && (wv.x_off < 0 || wv.x_off + wv.n > d->n_x))
      return fail(WFM_EINVAL, "channel %lld: explicit abscissae out of range", (long long)w);
    if (wv.out_off % 4) return fail(WFM_EINVAL, "channel %lld: out_off must be a multiple of 4 (16-byte stores)", (long long)w);
  }
  return WFM_OK;
}

extern "C" {

int wfm_abi_version(void) { return WFM_ABI_VERSION; }

const char* wfm_last_error(void) { return g_err; }

int wfm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int wfm_program_create(const WfmProgramDesc* d, int device, wfm_program_t* out) {
  if (!out) return fail(WFM_EINVAL, "null output handle");
  *out = nullptr;
  int rc = validate(d);
  if (rc != WFM_OK) return rc;
  DeviceGuard g(device);
  if (!g.ok) return fail(WFM_ECUDA, "cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
  WfmProgram* p = new (std::nothrow) WfmProgram();
  if (!p) return fail(WFM_ENOMEM, "out of host memory");
  p->device = device;
  p->waves.assign(d->waves, d->waves + d->n_waves);

  cudaError_t e = cudaSuccess;
  auto up = [&](auto host, int64_t n, auto dev) {
    if (e == cudaSuccess) e = upload(host, n, dev);
  };
  up(d->waves, d->n_waves, &p->dev.waves);
  up(d->seg_bound, d->n_segs, &p->dev.seg_bound);
  up(d->seg_ptr, d->n_segs ? d->n_segs + 1 : 0, &p->dev.seg_ptr);
  up(d->facs, d->n_facs, &p->dev.facs);
  up(d->terms, d->n_terms, &p->dev.terms);
  up(d->refs, d->n_refs, &p->dev.refs);
  up(d->args, d->n_args, &p->dev.args);
  up(d->x, d->n_x, &p->dev.x);

  // owning channel of every segment row (pre-pass only) and the widest segment
  int max_rows = 0;
  {
    std::vector<int32_t> seg_wave((size_t)d->n_segs, 0);
    for (int64_t w = 0; w < d->n_waves; ++w) {
      const WfmWave& wv = d->waves[w];
      std::fill(seg_wave.begin() + wv.seg_begin, seg_wave.begin() + wv.seg_begin + wv.n_seg, (int32_t)w);
    }
    for (int64_t sg = 0; sg < d->n_segs; ++sg)
      max_rows = std::max(max_rows, (int)(d->seg_ptr[sg + 1].fac - d->seg_ptr[sg].fac));
    up(seg_wave.data(), d->n_segs, &p->dev.seg_wave);
  }
  p->dev.n_slots = 1 + std::max(1, std::min(max_rows, wfm::kMaxSlots));
  int32_t* d_seg_start = nullptr;
  double* d_seg_val = nullptr;
  wfm::SegPlan* d_seg_plan = nullptr;
  uint8_t* d_row_slot = nullptr;
  wfm::CTerm* d_cterms = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&d_seg_start, sizeof(int32_t) * (size_t)std::max<int64_t>(d->n_segs, 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_seg_val, sizeof(double) * (size_t)std::max<int64_t>(d->n_segs, 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_seg_plan, sizeof(wfm::SegPlan) * (size_t)std::max<int64_t>(d->n_segs, 1));
  if (e == cudaSuccess) e = cudaMalloc(&d_row_slot, (size_t)std::max<int64_t>(d->n_facs, 16));
  if (e == cudaSuccess) e = cudaMalloc(&d_cterms, sizeof(wfm::CTerm) * (size_t)std::max<int64_t>(d->n_terms, 1));
  p->dev.seg_start = d_seg_start;
  p->dev.seg_val = d_seg_val;
  p->dev.seg_plan = d_seg_plan;
  p->dev.row_slot = d_row_slot;
  p->dev.cterms = d_cterms;

  // Tile size.  A warp's slice of shared memory holds the output tile (8 B per sample),
  // the value slots and two packet buffers (the tile's rows of the device tables,
  // double-buffered).  Take the largest tile (a multiple of 128 samples) whose AVERAGE
  // packet leaves a 4x margin in a packet buffer for tiles where pulses cluster; a tile
  // whose packet still does not fit takes the kernel's cold path.
  {
    int64_t samples = 0;
    for (int64_t w = 0; w < d->n_waves; ++w) samples += d->waves[w].n;
    const double per_sample = samples > 0 ? 1.0 / (double)samples : 0.0;
    const double ir = ((double)d->n_facs * 40.0 /* SRow 16, RRow 64, GRow 32; NOP rows vanish */ +
                       (double)d->n_terms * sizeof(wfm::CTerm) + (double)d->n_segs * sizeof(wfm::ARow)) * per_sample;
    const int fixed = wfm::warp_fixed_bytes(p->dev.n_slots);
    int ts = wfm::kMaxTileSamples, cap = 0;
    for (;; ts -= 128) {
      cap = ((wfm::kWarpSliceBytes - fixed - ts * 8) / 2) & ~15;
      if (ts <= wfm::kMinTileSamples || (cap >= 256 && 64.0 + 4.0 * ir * ts <= cap)) break;
    }
    p->dev.tile_samples = ts;
    p->dev.pkt_cap = std::max(cap, 64);
  }
  const int64_t tile_samples = p->dev.tile_samples;
  // tile list: tile_samples consecutive samples of one channel per tile
  std::vector<wfm::TileDesc> tiles;
  p->tile_prefix.resize(d->n_waves + 1);
  int64_t total = 0;
  for (int64_t w = 0; w < d->n_waves; ++w) {
    p->tile_prefix[w] = (int64_t)tiles.size();
    const WfmWave& wv = d->waves[w];
    for (int64_t j = 0; j < wv.n; j += tile_samples)
      tiles.push_back({j, wv.out_off + j, (int32_t)w, (int32_t)std::min<int64_t>(tile_samples, wv.n - j), 0, 0});
    total = std::max(total, wv.out_off + wv.n);
    if (wv.flags & WFM_WAVE_COMPLEX) p->any_complex = true;
  }
  p->tile_prefix[d->n_waves] = (int64_t)tiles.size();
  p->n_tiles = (int64_t)tiles.size();
  p->total_samples = total;
  if (e == cudaSuccess && p->n_tiles >= INT32_MAX) {
    const long long nt = (long long)p->n_tiles;
    delete p;
    return fail(WFM_EINVAL, "too many tiles (%lld)", nt);
  }
  if (e == cudaSuccess) {
    const wfm::TileDesc* dt = nullptr;
    e = upload(tiles.data(), (int64_t)tiles.size(), &dt);
    p->d_tiles = const_cast<wfm::TileDesc*>(dt);
  }
  // device pre-pass 1: segment start positions and flat values, device table formats,
  // every tile's segment range and packet size; then the packet offsets (exclusive scan)
  uint32_t *d_pkt_size = nullptr, *d_pkt_off = nullptr, *d_scratch = nullptr;
  const size_t nt1 = (size_t)p->n_tiles + 1;
  if (e == cudaSuccess) e = cudaMalloc(&d_pkt_size, sizeof(uint32_t) * nt1);
  if (e == cudaSuccess) e = cudaMalloc(&d_pkt_off, sizeof(uint32_t) * nt1);
  if (e == cudaSuccess) e = cudaMalloc(&d_scratch, sizeof(uint32_t) * (nt1 / 4096 + 2));
  p->dev.pkt_off = d_pkt_off;
  if (e == cudaSuccess)
    e = wfm::launch_prepare(p->dev, wfm::PrepareCounts{d->n_segs, d->n_facs, d->n_terms, p->n_tiles}, d_seg_start,
                            d_seg_val, d_seg_plan, d_row_slot, d_cterms, p->d_tiles, d_pkt_size, 0);
  if (e == cudaSuccess) e = wfm::launch_scan(d_pkt_size, d_pkt_off, d_scratch, p->n_tiles, 0);
  uint32_t total16 = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&total16, d_pkt_off + p->n_tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost);
  // pass 2: the packets themselves
  unsigned char* d_packets = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&d_packets, std::max<size_t>((size_t)total16 * 16, 16));
  p->dev.packets = d_packets;
  if (e == cudaSuccess) e = wfm::launch_fill_packets(p->dev, p->d_tiles, p->n_tiles, d_packets, 0);
  if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  cudaFree(d_pkt_size);
  cudaFree(d_scratch);
  if (e != cudaSuccess) {
    delete p;
    return fail(WFM_ECUDA, "uploading the program failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return WFM_OK;
}

int wfm_program_destroy(wfm_program_t prog) {
  delete prog;
  return WFM_OK;
}

int64_t wfm_program_total_samples(wfm_program_t prog) { return prog ? prog->total_samples : -1; }

int64_t wfm_program_launch_count(wfm_program_t prog) { return prog ? prog->launches : -1; }

static int check_launch(wfm_program_t prog, const WfmLaunch* l, int64_t* first, int64_t* count, int64_t* need) {
  if (!prog || !l) return fail(WFM_EINVAL, "null program or launch");
  const int64_t nw = (int64_t)prog->waves.size();
  *first = l->first_wave;
  *count = l->n_wave == 0 ? nw - l->first_wave : l->n_wave;
  if (*first < 0 || *count < 0 || *first + *count > nw) return fail(WFM_EINVAL, "channel range out of bounds");
  if (l->dtype != WFM_F64 && l->dtype != WFM_F32 && l->dtype != WFM_C128) return fail(WFM_EINVAL, "bad dtype");
  if (prog->any_complex && l->dtype != WFM_C128)
    for (int64_t w = *first; w < *first + *count; ++w)
      if (prog->waves[w].flags & WFM_WAVE_COMPLEX)
        return fail(WFM_EINVAL, "channel %lld has complex amplitudes: request WFM_C128 output", (long long)w);
  int64_t ext = 0;
  for (int64_t w = *first; w < *first + *count; ++w) ext = std::max(ext, prog->waves[w].out_off + prog->waves[w].n);
  *need = ext;
  if (!l->out && ext > 0) return fail(WFM_EINVAL, "null output buffer");
  if (l->out_elems < ext) return fail(WFM_EINVAL, "output buffer too small: %lld < %lld", (long long)l->out_elems, (long long)ext);
  return WFM_OK;
}

int wfm_sample(wfm_program_t prog, const WfmLaunch* l, void* stream) {
  int64_t first, count, need;
  int rc = check_launch(prog, l, &first, &count, &need);
  if (rc != WFM_OK) return rc;
  if ((uintptr_t)l->out % 16) return fail(WFM_EINVAL, "output buffer must be 16-byte aligned");
  DeviceGuard g(prog->device);
  const int64_t t0 = prog->tile_prefix[first], t1 = prog->tile_prefix[first + count];
  WFM_CUDA(wfm::launch_sample(prog->dev, prog->d_tiles, t0, t1 - t0, l->dtype, l->accumulate, l->out,
                              (cudaStream_t)stream));
  if (t1 > t0) prog->launches += 1;
  return WFM_OK;
}

int wfm_sample_host(wfm_program_t prog, const WfmLaunch* l) {
  int64_t first, count, need;
  int rc = check_launch(prog, l, &first, &count, &need);
  if (rc != WFM_OK) return rc;
  if (l->accumulate) return fail(WFM_EUNSUPPORTED, "accumulate is not available with host buffers");
  DeviceGuard g(prog->device);
  const size_t esz = l->dtype == WFM_F64 ? 8 : (l->dtype == WFM_F32 ? 4 : 16);
  // only the extent actually covered by the requested channels is staged/copied
  int64_t lo = INT64_MAX;
  for (int64_t w = first; w < first + count; ++w) lo = std::min(lo, prog->waves[w].out_off);
  if (count == 0 || need == 0) return WFM_OK;
  const size_t bytes = (size_t)need * esz;
  if (prog->stage_bytes < bytes) {
    cudaFree(prog->d_stage);
    prog->d_stage = nullptr;
    prog->stage_bytes = 0;
    WFM_CUDA(cudaMalloc(&prog->d_stage, bytes));
    prog->stage_bytes = bytes;
  }
  // padding between channels is never written by the kernel: keep it defined
  WFM_CUDA(cudaMemsetAsync((char*)prog->d_stage + (size_t)lo * esz, 0, (size_t)(need - lo) * esz, 0));
  const int64_t t0 = prog->tile_prefix[first], t1 = prog->tile_prefix[first + count];
  WFM_CUDA(wfm::launch_sample(prog->dev, prog->d_tiles, t0, t1 - t0, l->dtype, 0, prog->d_stage, 0));
  if (t1 > t0) prog->launches += 1;
  WFM_CUDA(cudaMemcpyAsync((char*)l->out + (size_t)lo * esz, (char*)prog->d_stage + (size_t)lo * esz,
                           (size_t)(need - lo) * esz, cudaMemcpyDeviceToHost, 0));
  WFM_CUDA(cudaStreamSynchronize(0));
  return WFM_OK;
}

}  // extern "C"

// wfm_expand.cu — pulse TEMPLATES expanded into the flat IR on the device.
//
// A gate sequence repeats a handful of pulse shapes thousands of times; its flat IR (WfmFactor / WfmTerm / WfmRef rows
// per pulse) is as large as its output (the randomized-benchmarking batch: 0.86 GB of tables for 0.69 GB of samples).
// What really differs from pulse to pulse is a few numbers: the start-time-dependent shifts, sometimes an amplitude,
// a phase-derived argument.  wfm_expand_templates takes
//   * the tables of every TEMPLATE once (what lowering produces for one pulse of that shape),
//   * per template a PATCH list (which table entries depend on the pulse's parameters) and its rotation rows,
//   * per pulse: its template, its four destination offsets and a short PAYLOAD of doubles (one per patch),
// and writes the per-pulse rows into device tables that wfm_program_create then takes as they are
// (WFM_DESC_DEVICE_TABLES).  Host -> device traffic and host build time drop by an order of magnitude; the kernels see
// the same IR as before.  The host-side producer is waveforms_b200.builder (pulse_train_batch(compact=True)); the
// reference has no counterpart (it builds one Python object per pulse: waveforms/_waveform.pyx:68-88, :216-235).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/wfm_b200.h"

namespace {

__global__ void __launch_bounds__(256) expand_kernel(WfmExpandDesc D, WfmFactor* __restrict__ facs, WfmTerm* __restrict__ terms,
                                                     WfmRef* __restrict__ refs, double* __restrict__ args) {
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);  // one warp per pulse
  const int lane = threadIdx.x & 31;
  if (p >= D.n_pulses) return;
  const WfmTemplateDesc T = D.templates[D.pulse_tmpl[p]];
  const int64_t f0 = D.pulse_fac[p], t0 = D.pulse_term[p], r0 = D.pulse_ref[p], a0 = D.pulse_arg[p];
  const double* __restrict__ pay = D.payload + p * (int64_t)D.payload_stride;
  for (int r = lane; r < T.n_fac; r += 32) {
    WfmFactor f = D.t_facs[T.fac0 + r];
    if (D.t_has_args[T.fac0 + r]) f.arg_off += (int32_t)a0;
    facs[f0 + r] = f;
  }
  for (int t = lane; t < T.n_term; t += 32) {
    WfmTerm tm = D.t_terms[T.term0 + t];
    tm.ref_begin += (int32_t)r0;
    terms[t0 + t] = tm;
  }
  for (int r = lane; r < T.n_ref; r += 32) refs[r0 + r] = D.t_refs[T.ref0 + r];
  for (int a = lane; a < T.n_arg; a += 32) args[a0 + a] = D.t_args[T.arg0 + a];
  __syncwarp();
  // the entries that depend on this pulse's parameters: payload slot j of the template = patch j
  for (int j = lane; j < T.n_patch; j += 32) {
    const WfmPatch pt = D.patches[T.patch0 + j];
    const double v = pay[j];
    switch (pt.kind) {
      case WFM_PATCH_SHIFT: facs[f0 + pt.index].shift = v; break;
      case WFM_PATCH_A0: facs[f0 + pt.index].a0 = v; break;
      case WFM_PATCH_A1: facs[f0 + pt.index].a1 = v; break;
      case WFM_PATCH_ARG: args[a0 + pt.index] = v; break;
      case WFM_PATCH_AMP: terms[t0 + pt.index].amp_re = v; break;
      default: break;  // WFM_PATCH_VALUE: only read by the rotation rows below
    }
  }
  __syncwarp();
  // rotation rows (WFM_COS_ROT): their pool block [base slot, base shift, D, cos D, sin D] follows from the row's own
  // shift, its base row's shift and w: D = w * (s_b - s_t), as lowering._emit_rows computes it
  for (int j = lane; j < T.n_rot; j += 32) {
    const WfmRotRow rr = D.rots[T.rot0 + j];
    const double w = rr.w_slot >= 0 ? pay[rr.w_slot] : rr.w;
    const double sb = rr.sb_slot >= 0 ? pay[rr.sb_slot] : rr.s_b;
    const double st = facs[f0 + rr.fac_row].shift;
    const double dl = __dmul_rn(w, __dsub_rn(sb, st));
    double sn, cs;
    sincos(dl, &sn, &cs);
    double* blk = args + a0 + rr.arg_off;
    blk[1] = sb;
    blk[2] = dl;
    blk[3] = cs;
    blk[4] = sn;
  }
}

}  // namespace

extern "C" int wfm_expand_templates(const WfmExpandDesc* d, WfmFactor* facs, WfmTerm* terms, WfmRef* refs, double* args,
                                    void* stream) {
  if (!d || d->n_pulses < 0 || d->payload_stride < 0) return WFM_EINVAL;
  if (d->n_pulses == 0) return WFM_OK;
  if (!d->templates || !d->pulse_tmpl || !d->pulse_fac || !d->pulse_term || !d->pulse_ref || !d->pulse_arg) return WFM_EINVAL;
  const int warps = 8;
  expand_kernel<<<(unsigned)((d->n_pulses + warps - 1) / warps), 32 * warps, 0, (cudaStream_t)stream>>>(*d, facs, terms, refs, args);
  return cudaGetLastError() == cudaSuccess ? WFM_OK : WFM_ECUDA;
}

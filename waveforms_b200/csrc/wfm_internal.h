// wfm_internal.h — structures shared by the C-ABI layer and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/wfm_b200.h"

namespace wfm {

// samples per CTA tile of the sampling kernel: chosen per program from the segment
// density (sparse programs amortise the tile prologue over more samples, dense
// ones keep their table slice small enough to stage in shared memory)
constexpr int kMinTileSamples = 2048;
constexpr int kMaxTileSamples = 16384;

// device-resident copy of a lowered batch (all DEVICE pointers)
struct DevProgram {
  const WfmWave* waves;
  const double* seg_bound;
  const WfmSegPtr* seg_ptr;
  const WfmFactor* facs;
  const WfmTerm* terms;
  const WfmRef* refs;
  const double* args;
  const double* x;
  int tile_samples;  // kMinTileSamples .. kMaxTileSamples, power of two
};

// one CTA's work item: DevProgram::tile_samples consecutive samples of channel `wave`
struct TileDesc {
  int64_t j0;
  int32_t wave;
  // filled on the device by prepare_tiles_kernel (once per program):
  int32_t seg_lo;   // channel-relative segment of the tile's first abscissa
  int32_t seg_hi;   // ... and of its last one
  int32_t fac0, n_fac;    // slice of the factor table the tile can touch
  int32_t term0, n_term;  // ... of the term table
  int32_t ref0, n_ref;    // ... of the reference table
  int32_t reserved;
};
static_assert(sizeof(TileDesc) == 48, "TileDesc layout");

cudaError_t launch_prepare_tiles(const DevProgram& P, TileDesc* tiles, int64_t n_tiles, cudaStream_t stream);

cudaError_t launch_sample(const DevProgram& P, const TileDesc* tiles, int64_t n_tiles, int dtype, int accumulate,
                          void* out, cudaStream_t stream);

}  // namespace wfm

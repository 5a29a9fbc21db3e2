// wfm_internal.h — structures shared by the C-ABI layer and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/wfm_b200.h"

namespace wfm {

// samples per CTA tile of the sampling kernel (both fp64 and fp32 outputs)
constexpr int kTileSamples = 4096;

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
};

// one CTA's work item: kTileSamples consecutive samples of channel `wave`
struct TileDesc {
  int64_t j0;
  int32_t wave;
  int32_t seg_lo;  // channel-relative segment of the tile's first abscissa  } filled on the device by
  int32_t seg_hi;  // ... and of its last one                                } prepare_tiles_kernel
  int32_t reserved;
};

cudaError_t launch_prepare_tiles(const DevProgram& P, TileDesc* tiles, int64_t n_tiles, cudaStream_t stream);

cudaError_t launch_sample(const DevProgram& P, const TileDesc* tiles, int64_t n_tiles, int dtype, int accumulate,
                          void* out, cudaStream_t stream);

}  // namespace wfm

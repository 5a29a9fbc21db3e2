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

// Compact term built at upload from WfmTerm + WfmRef: amplitude and up to six
// factor slots with exponent 1 in ONE 16-byte record (one load per term in the
// interpreter).  Terms that do not fit (more refs, an exponent != 1, a slot
// beyond the value cache) carry kCTermExt and are read from the ABI tables.
struct CTerm {
  double amp;
  uint8_t n_ref;
  uint8_t flags;
  uint8_t slot[6];
};
static_assert(sizeof(CTerm) == 16, "CTerm layout");
constexpr uint8_t kCTermGroupEnd = 1, kCTermExt = 2;
constexpr int kMaxSlots = 12;  // distinct factor values cached per segment evaluation

// device-resident copy of a lowered batch (all DEVICE pointers)
struct DevProgram {
  const WfmWave* waves;
  const double* seg_bound;
  const WfmSegPtr* seg_ptr;
  const WfmFactor* facs;
  const WfmTerm* terms;
  const CTerm* cterms;  // parallel to terms
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

// wfm_internal.h — structures shared by the C-ABI layer and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/wfm_b200.h"

namespace wfm {

// samples per CTA tile of the sampling kernel: the tile is assembled in shared
// memory (tile_samples x 8 B) and stored with one TMA bulk copy.  Chosen per
// program from the segment density: dense programs take smaller tiles so that the
// tile's segment rows (<= 256) and table slice still fit in shared memory.
constexpr int kMinTileSamples = 1024;
#ifndef WFM_K1_MAX_TILE
#define WFM_K1_MAX_TILE 4096
#endif
constexpr int kMaxTileSamples = WFM_K1_MAX_TILE;

// Compact term built on the device at upload from WfmTerm + WfmRef: amplitude and up
// to six factor slots with exponent 1 in ONE 16-byte record (one load per term in the
// interpreter).  packed: bits 0..7 n_ref, 8..15 flags, 16+8r..23+8r slot r.  Terms
// that do not fit (more refs, an exponent != 1, a slot beyond the value cache) carry
// kCTermExt and are read from the ABI tables.
struct CTerm {
  double amp;
  uint64_t packed;
};
static_assert(sizeof(CTerm) == 16, "CTerm layout");
constexpr uint32_t kCTermGroupEnd = 1, kCTermExt = 2;
constexpr int kMaxSlots = 12;  // distinct factor values cached per segment evaluation

// Device factor row built at upload from WfmFactor (+ its argument-pool block for
// the rotation rows): everything the hot basis functions need in ONE 64-byte row
// that is staged in shared memory with the tile.
//   WFM_COS_ROT: aux = base slot, p = {base_shift, D, cos D, sin D}
//   others     : aux = arg_off (argument pool stays in global memory)
struct DFactor {
  int32_t func;
  int32_t aux;
  double shift;
  double a0, a1;
  double p[4];
};
static_assert(sizeof(DFactor) == 64, "DFactor layout");

constexpr int kMaxIrBytes = 24576;  // cap of the per-tile table slice staged in shared memory

// device-resident copy of a lowered batch (all DEVICE pointers)
struct DevProgram {
  const WfmWave* waves;
  const double* seg_bound;
  const WfmSegPtr* seg_ptr;
  const WfmFactor* facs;   // ABI rows (pre-pass only)
  const WfmTerm* terms;
  const WfmRef* refs;
  const double* args;
  const double* x;
  // built on the device once per program (prepare kernels):
  const DFactor* dfacs;      // parallel to facs
  const CTerm* cterms;       // parallel to terms
  const int32_t* seg_start;  // [n_segs] first sample (channel-relative) owned by the segment
  const double* seg_val;     // [n_segs] value of a FLAT segment (offset + constant terms, clipped)
  const int32_t* seg_wave;   // [n_segs] owning channel (host-built; pre-pass only)
  int tile_samples;  // kMinTileSamples .. kMaxTileSamples, power of two
  int n_slots;       // factor-value slots per thread in shared memory (max rows per segment, <= kMaxSlots)
  int ir_bytes;      // shared-memory budget for a tile's factor / compact-term slice
};

// one CTA's work item: up to DevProgram::tile_samples consecutive samples of channel `wave`
struct TileDesc {
  int64_t j0;    // first sample of the tile in its channel
  int64_t out0;  // index of that sample in the output buffer
  int32_t wave;
  int32_t cnt;   // samples in the tile
  // filled on the device by prepare_tiles_kernel (once per program):
  int32_t seg0;           // ABSOLUTE segment row that owns the tile's first sample
  int32_t nb;             // segment rows the tile spans
  int32_t fac0, n_fac;    // slice of the factor table the tile can touch
  int32_t term0, n_term;  // ... of the term table
};
static_assert(sizeof(TileDesc) == 48, "TileDesc layout");

struct PrepareCounts {
  int64_t n_segs, n_facs, n_terms, n_tiles;
};

// segment start positions, flat values, device factor rows, compact terms, tile ranges
cudaError_t launch_prepare(const DevProgram& P, const PrepareCounts& n, int32_t* seg_start, double* seg_val,
                           DFactor* dfacs, CTerm* cterms, TileDesc* tiles, int* max_ir_bytes, cudaStream_t stream);
size_t sample_smem_bytes(const DevProgram& P, int dtype);

cudaError_t launch_sample(const DevProgram& P, const TileDesc* tiles, int64_t n_tiles, int dtype, int accumulate,
                          void* out, cudaStream_t stream);

}  // namespace wfm

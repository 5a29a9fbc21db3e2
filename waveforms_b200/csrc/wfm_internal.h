// wfm_internal.h — structures shared by the C-ABI layer and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/wfm_b200.h"

namespace wfm {

// samples per tile of the sampling kernel.  A tile is assembled by ONE WARP in its
// private slice of shared memory (tile_samples x 8 B) and stored with one TMA bulk
// copy.  Chosen per program from the table density: dense programs take smaller
// tiles so that the tile's segment rows (<= kStageSegs) and its slice of the
// factor / term tables fit the warp's slice.
#ifndef WFM_K1_MIN_BLOCKS
#define WFM_K1_MIN_BLOCKS 2  // resident 8-warp CTAs per SM the kernel is sized for
#endif
#ifndef WFM_K1_MAX_TILE
#define WFM_K1_MAX_TILE 1024
#endif
constexpr int kMinTileSamples = 128;
constexpr int kMaxTileSamples = WFM_K1_MAX_TILE;
// shared memory per warp: 8 * WFM_K1_MIN_BLOCKS warps share the SM's 227 KB (1 KB per CTA is reserved)
constexpr int kWarpSliceBytes = ((227 * 1024 / WFM_K1_MIN_BLOCKS - 1024) / 8) & ~127;

// Compact term built on the device at upload from WfmTerm + WfmRef: amplitude and up
// to three factor slots with exponent 1 in ONE 16-byte record (one load per term in
// the interpreter).  packed: bits 0..7 flags, 8..15 / 16..23 / 24..31 PHYSICAL value
// slots of the three references; a missing reference points at physical slot 0, which
// always holds 1.0 (x * 1.0 is exact), so the product needs no loop and no branch.
// Terms that do not fit (more refs, an exponent != 1, a slot beyond the value cache)
// carry kCTermExt and are read from the ABI tables.
struct CTerm {
  double amp;
  uint64_t packed;
};
static_assert(sizeof(CTerm) == 16, "CTerm layout");
constexpr uint32_t kCTermGroupEnd = 1, kCTermExt = 2;
constexpr int kMaxSlots = 12;  // distinct factor values cached per segment evaluation (physical slots 1..12)

// interpreter op of a factor row (upper 16 bits of DFactor::func), hottest first
enum : int { OP_ROT = 0, OP_SINCOS = 1, OP_NOP = 2, OP_COS = 3, OP_LINEAR = 4, OP_GAUSSIAN = 5, OP_ERF = 6, OP_GENERIC = 7 };

// Device factor row built at upload from WfmFactor (+ its argument-pool block for
// the rotation rows): everything the hot basis functions need in ONE 64-byte row
// that is staged in shared memory with the tile.
//   WFM_COS_ROT: aux = base slot, p = {base_shift, D, cos D, sin D}
//   others     : aux = arg_off (argument pool stays in global memory)
// func = WFM_* id | OP_* << 16
struct DFactor {
  int32_t func;
  int32_t aux;
  double shift;
  double a0, a1;
  double p[4];
};
static_assert(sizeof(DFactor) == 64, "DFactor layout");


// ---- tile packets: the device IR the sampling kernel executes ------------------------------
// Built once per program on the device.  A packet is everything ONE tile needs,
// contiguous in global memory (16-byte aligned, size a multiple of 16) so that a
// warp brings it into shared memory with ONE TMA bulk copy:
//   PacketHeader | ARow[n_arows + 1] | PatchRow[n_patch] | DFactor[n_fac] | CTerm[n_term]
// Zero segments do not appear at all: the kernel fills the tile with `base` first.
struct PacketHeader {
  int64_t out0;   // index of the tile's first sample in the output buffer
  int64_t j0;     // first sample of the tile in its channel
  double base;    // value of the channel's zero segments (its accumulator start)
  double t0, delta;  // affine grid of the channel: x[j] = t0 + j*delta
  int32_t wave;
  uint32_t flags;    // WfmWave::flags | kPacketCold
  uint16_t cnt;      // samples in the tile
  uint16_t n_arows;  // active segments intersecting the tile
  uint16_t n_patch;  // flat segments whose value differs from `base`
  uint16_t n_active; // active samples
  uint16_t n_fac;    // factor rows in the packet
  uint16_t n_term;   // compact terms in the packet
  uint32_t reserved;
};
static_assert(sizeof(PacketHeader) == 64, "PacketHeader layout");
constexpr uint32_t kPacketCold = 0x80000000u;  // header only: the tile takes the global-table path

// one ACTIVE segment of the tile; row n_arows is a sentinel closing the ranges
struct ARow {
  uint16_t start;     // first tile-sample of the segment
  uint16_t first;     // active samples of the tile before this segment
  uint16_t fac_rel;   // its first factor row within the packet
  uint16_t term_rel;  // its first compact term within the packet
  int32_t gfac;       // the same rows in the global tables (extended terms only)
  int32_t gterm;
};
static_assert(sizeof(ARow) == 16, "ARow layout");

// a flat run [a, b) of the tile with its own value
struct PatchRow {
  uint16_t a, b;
  uint32_t reserved;
  double val;
};
static_assert(sizeof(PatchRow) == 16, "PatchRow layout");

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
  const uint32_t* pkt_off;       // [n_tiles + 1] packet offsets in 16-byte units
  const unsigned char* packets;  // the tile packets
  int tile_samples;  // kMinTileSamples .. kMaxTileSamples, power of two
  int n_slots;       // PHYSICAL value slots per lane in shared memory: 1 (the constant 1.0) + max rows per segment
  int pkt_cap;       // bytes of ONE packet buffer in a warp's shared slice (two buffers per warp)
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

// pass 1: segment start positions, flat values, device factor rows, compact terms,
// every tile's segment range and packet size (16-byte units) -> pkt_size[n_tiles]
cudaError_t launch_prepare(const DevProgram& P, const PrepareCounts& n, int32_t* seg_start, double* seg_val,
                           DFactor* dfacs, CTerm* cterms, TileDesc* tiles, uint32_t* pkt_size, cudaStream_t stream);
// exclusive scan: pkt_size[n] -> pkt_off[n + 1] (pkt_off[n] = total); scratch holds ceil(n / 4096) + 1 words
cudaError_t launch_scan(const uint32_t* pkt_size, uint32_t* pkt_off, uint32_t* scratch, int64_t n, cudaStream_t stream);
// pass 2: write the packets
cudaError_t launch_fill_packets(const DevProgram& P, const TileDesc* tiles, int64_t n_tiles, unsigned char* packets,
                                cudaStream_t stream);
// shared memory per warp besides the output tile and the two packet buffers
int warp_fixed_bytes(int n_slots);
size_t sample_smem_bytes(const DevProgram& P, int dtype);

cudaError_t launch_sample(const DevProgram& P, const TileDesc* tiles, int64_t tile_begin, int64_t n_tiles, int dtype,
                          int accumulate, void* out, cudaStream_t stream);

}  // namespace wfm

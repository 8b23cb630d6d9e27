// Program format of the MASKED tiled kernel (blur_masked.cu): the round-1 kernel, kept for PSFs whose program fits
// kMaskedMaxChunks (taps.cu: 4) chunks of <= 18 rows x 20 columns -- every low-exposure PSF and most high-exposure ones --
// where it is the faster of the two tiled kernels (tools/exp/route_probe.py).  Larger PSFs get the dense sheared program
// of dib_common.cuh (blur_tiled.cu).  Names live in dib::mk and shadow the dense
// kernel's geometry constants of the same name.
#pragma once
#include "dib_common.cuh"

namespace dib {
namespace mk {

// Tiled-kernel program of one PSF (built by taps.cu, executed by blur_masked.cu).  The PSF support is cut into
// groups of kGroupW columns; a SEGMENT is one group's run of rows [dy0, dy0 + nsteps) with one kGroupW-wide weight
// vector per row ("step"; zero where the PSF has no tap -- the kernel skips those with uniform branches).  Segments
// are packed into CHUNKS whose tap extents are bounded (rows <= kChunkHaloRows, columns <= kChunkGroups groups) so that
// tile + halo of any chunk fits the kernel's fixed shared-memory stage, whatever the PSF's overall extent.
#ifndef DIB_GW
#define DIB_GW 4
#endif
constexpr int kGroupW = DIB_GW;         // PSF columns per group (kGroupW / 4 float4 of weights per step)
constexpr int kChunkGroups = 20 / kGroupW;   // groups per chunk  -> column halo <= 19 (GW 4) / 15 (GW 8)
static_assert(kGroupW == 4 || kGroupW == 8, "weight vectors are read as float4");
constexpr int kChunkHaloRows = 17;      // dy_hi - dy_lo per chunk
constexpr int kProgMaxChunks = 32;
struct SegRec {         // 8 bytes
    int16_t dx0;        // first tap column of the group, relative to the PSF centre (tap dx = x - centre)
    int16_t dy0;        // first tap row of the run, relative to the centre
    int16_t nsteps;     // rows in the run
    int16_t woff;       // index of the run's first weight vector inside the chunk's weight array
};
struct ChunkRec {       // 16 bytes
    int16_t dy_lo, dy_hi;   // tap row range of the chunk (relative to the centre)
    int16_t dx_lo, dx_hi;   // tap column range: first group's dx0 .. last group's dx0 + kGroupW - 1
    int16_t nseg;           // segments in the chunk (<= kChunkGroups)
    int16_t wsteps;         // weight vectors in the chunk
    int32_t data_off;       // byte offset of the chunk's data block inside the PSF's program section
};
// chunk data block: SegRec slots (48 B, fixed) then float4[wsteps] (+ one zero vector: the kernel prefetches one ahead)
constexpr int kChunkSegBytes = 48;
constexpr int kChunkMaxSteps = kChunkGroups * (kChunkHaloRows + 1);                       // 90
constexpr int kStepBytes = 4 * kGroupW;                                                    // one weight vector
constexpr int kChunkDataMax = kChunkSegBytes + kStepBytes * (kChunkMaxSteps + 1);
constexpr size_t kProgHeaderBytes = sizeof(ChunkRec) * kProgMaxChunks;                    // 512
constexpr size_t kProgDataBytes = 16384;
constexpr size_t kProgBytes = kProgHeaderBytes + kProgDataBytes;

}  // namespace mk
}  // namespace dib

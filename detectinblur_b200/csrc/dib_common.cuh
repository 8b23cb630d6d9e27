// Shared device/host helpers for libdib.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dib.h"

namespace dib {

// ---- error reporting (thread-local message, no global mutable state across threads) ----
void set_error(const char* fmt, ...);

#define DIB_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            dib::set_error(__VA_ARGS__);    \
            return DIB_ERR_INVALID;         \
        }                                   \
    } while (0)

#define DIB_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            dib::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return DIB_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

// ---- tap set layout (see include/dib.h) ----
// Geometry of the tiled kernel's register tiling (blur_tiled.cu); the program builder (taps.cu) shares it because the
// extents of a program chunk are bounded by what one shared-memory stage of the kernel holds.
#ifndef DIB_R
#define DIB_R 4
#endif
#ifndef DIB_WARP_ROWS
#define DIB_WARP_ROWS 4
#endif
#ifndef DIB_WARP_COLS
#define DIB_WARP_COLS 2
#endif
#ifndef DIB_STAGES
#define DIB_STAGES 2
#endif
#ifndef DIB_CTAS_PER_SM
#define DIB_CTAS_PER_SM 1
#endif
constexpr int kR = DIB_R;                   // row pairs per thread (rows r and r + kR share 64-bit accumulators: FFMA2)
constexpr int kRows = 2 * kR;               // output rows per thread = rotation period of the register window
constexpr int kCC = 7;                      // output columns per thread (odd lane stride: conflict-free scalar LDS)
constexpr int kWarpW = 32 * kCC;            // 224 output columns per warp
constexpr int kWarpRows = DIB_WARP_ROWS;    // compute warps are arranged kWarpRows x kWarpCols over a tile
constexpr int kWarpCols = DIB_WARP_COLS;
constexpr int kComputeWarps = kWarpRows * kWarpCols;
constexpr int kTH = kWarpRows * kRows;      // output rows per tile
constexpr int kTW = kWarpCols * kWarpW;     // output columns per tile
constexpr int kPitch = 512;                 // floats per staged row: two 256-element TMA boxes
constexpr int kOutPitch = kWarpW + 4;       // one staged output row of a warp (skew <= 3)
constexpr int kStageHdrBytes = 64;
// Tiled-kernel program of one PSF (built by taps.cu, executed by blur_tiled.cu).  The PSF support is SHEARED by `shear`
// columns per row (x' = x - shear * (y - ymin): a slanted motion streak becomes near-vertical) and cut into GROUPS of
// `group_w` sheared columns (2 or 4).  A SEGMENT is one group's run of rows [dy0, dy0 + nsteps) with one group_w-wide
// weight vector per row ("step"); the kernel executes every step DENSELY -- all group_w taps, zero weight where the PSF
// has none -- so the sweep has no data-dependent branches.  Segments are packed into CHUNKS (a band of neighbouring groups
// x a window of rows) whose true tap extents fit one shared-memory stage: rows <= kChunkTapRows, columns <= kPitch - kTW
// minus what the shear costs.
constexpr int kChunkTapRowsCap = 36;        // most PSF rows a chunk may span, whatever the stage could hold
constexpr int kChunkAuxBytes = 3776;        // segment records + weight vectors of one chunk (stage header + aux = 30 * 128 B)
constexpr int kStageBytesFor(int rows) { return kStageHdrBytes + kChunkAuxBytes + rows * kPitch * 4; }
constexpr int kStages = DIB_STAGES;         // shared-memory stages per CTA (1: the CTAs co-resident on an SM alternate instead)
constexpr int kCtasPerSm = DIB_CTAS_PER_SM;
// 228 KB of shared memory per SM, 1 KB of which the system reserves per resident CTA; at most 227 KB per CTA
constexpr int kSmemBudget = (233472 / kCtasPerSm - 1024 < 232448 ? 233472 / kCtasPerSm - 1024 : 232448) - 128;   // minus barriers / ticket slots
constexpr int kOutBufBytes = kComputeWarps * 2 * kOutPitch * 4;
constexpr int kRowsMaxRaw = (kSmemBudget - kOutBufBytes - kStages * (kStageHdrBytes + kChunkAuxBytes)) / (kStages * kPitch * 4);
constexpr int kRowsMax = kRowsMaxRaw > 60 ? 60 : kRowsMaxRaw;   // staged rows (one TMA box per producer thread and row half)
constexpr int kChunkTapRows = kRowsMax - kTH + 1 < kChunkTapRowsCap ? kRowsMax - kTH + 1 : kChunkTapRowsCap;   // PSF rows per chunk
static_assert(kChunkTapRows >= 8, "tile too tall for the shared-memory stage");
constexpr int kShearMax = (kRows <= 6) ? 2 : 1;       // |shear|: (kRows - 1) * |shear| extra columns per tile row band
constexpr int kChunkSegSlots = 12;          // segments (groups) per chunk
constexpr int kChunkSegBytes = kChunkSegSlots * 8;
constexpr int kChunkMaxCols = 24;           // sheared columns per band (kChunkSegSlots groups of 2, or 6 of 4)
constexpr int kChunkMaxWeightBytes = kChunkAuxBytes - kChunkSegBytes - 16;   // weight vectors (+ one zero vector past the end)
static_assert(kChunkMaxCols * kChunkTapRows * 4 <= kChunkMaxWeightBytes, "chunk weights must fit the stage's aux area");
constexpr int kProgMaxChunks = 48;
// columns of halo a chunk may need: true dx extent + the shear's per-band parallelogram + the <= 3 elements a staged row
// is skewed by (TMA boxes start 16-byte aligned in global memory) must fit kPitch - kTW
__host__ __device__ constexpr int chunk_col_span(int shear) { return kPitch - kTW - 3 - (kRows - 1) * (shear < 0 ? -shear : shear); }
// sheared columns per band such that any kChunkTapRows-row chunk of the band keeps its true dx extent within the stage
__host__ __device__ constexpr int band_cols(int shear) {
    const int a = shear < 0 ? -shear : shear;
    const int c = chunk_col_span(shear) + 1 - (kChunkTapRows - 1) * a;
    return c < kChunkMaxCols ? c : kChunkMaxCols;
}
static_assert(band_cols(kShearMax) >= 4 && band_cols(-kShearMax) >= 4, "shear leaves no room for a group");
struct SegRec {         // 8 bytes
    int16_t dx0;        // tap column of the group's first column at step 0, relative to the PSF centre (dx = x - centre);
                        // at step s the group covers dx0 + shear * s .. + group_w - 1
    int16_t dy0;        // first tap row of the run, relative to the centre
    int16_t nsteps;     // rows in the run
    int16_t woff;       // index of the run's first weight vector inside the chunk's weight array
};
struct ChunkRec {       // 16 bytes
    int16_t dy_lo, dy_hi;   // tap row range of the chunk (relative to the centre)
    int16_t dx_lo, dx_hi;   // true tap column range over all steps of all segments (including zero-weight group columns)
    int16_t wsteps;         // weight vectors in the chunk
    uint8_t nseg;           // segments in the chunk (<= kChunkSegSlots)
    int8_t shear;           // columns per row the groups move (same for every chunk of a PSF)
    uint8_t group_w;        // 2 or 4
    uint8_t pad;
    uint16_t data_off16;    // offset of the chunk's data block inside the PSF's program section, in 16-byte units
};
static_assert(sizeof(ChunkRec) == 16 && sizeof(SegRec) == 8, "program records are read as vectors");
// chunk data block: SegRec slots (kChunkSegBytes, fixed) then float[wsteps + 1][group_w] (the last vector is zero: the
// kernel prefetches one vector ahead)
constexpr size_t kProgHeaderBytes = sizeof(ChunkRec) * kProgMaxChunks;                    // 768
constexpr size_t kProgDataBytes = 32768;
constexpr size_t kProgBytes = kProgHeaderBytes + kProgDataBytes;

// Work-distribution words of the tiled kernel (self-resetting, see blur_tiled.cu)
struct SchedWords {
    unsigned int next_tile;     // next tile to hand out
    unsigned int done_ctas;     // CTAs that have stopped fetching
};
constexpr int kSchedSlots = 4;      // launches that may overlap (DIB_ALGO_OVERLAP) rotate through these (256 bytes are reserved)


// Device-side launch planning (DIB_ALGO_DEVICE_PLAN): which kernel takes a PSF is read from the tap set's summaries on the
// device instead of a host copy of them.  0: no tiled program; kMasked / kDense as in blur_api.cu.
__host__ __device__ inline int psf_program_kind(const dib_psf_meta& m) {
    if (m.count <= 0 || m.prog_chunks <= 0 || (m.flags & (DIB_META_NO_PROGRAM | DIB_META_TRUNCATED))) return 0;
    return m.prog_group_w == 0 ? 1 : 2;
}

inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

inline dib_tapset_layout tapset_layout(int n, int max_taps) {
    dib_tapset_layout L;
    L.meta_offset = 0;
    L.taps_offset = align256(sizeof(dib_psf_meta) * (size_t)n);
    L.prog_offset = L.taps_offset + align256(sizeof(dib_tap) * (size_t)n * (size_t)max_taps);
    L.prog_bytes_per_psf = kProgBytes;
    L.sched_offset = L.prog_offset + kProgBytes * (size_t)n;
    L.total_bytes = L.sched_offset + 256;
    return L;
}

// ---- index maps of manual_blur (models/blur_functions.py:17-69) ----
// Source row/column read by output position i for a tap at PSF coordinate `coord`; -1 = zero padding.
//   padded length L = n + lo + hi, out[i] = padded[(i + 2*lo - coord) mod L], padded[q] = img[map(q - lo)]
__host__ __device__ inline int src_index(int i, int n, int coord, int pad_mode) {
    const int lo = (pad_mode == DIB_PAD_REPLICATE256) ? 127 : 63;
    const int L = n + 2 * lo + 1;
    int q = (i + 2 * lo - coord) % L;
    if (q < 0) q += L;          // torch.roll wraps: a tap on the last PSF row/col reads padded index L-1 at i = 0
    int s = q - lo;
    if (pad_mode == DIB_PAD_REFLECT128) {
        if (s < 0) s = -s;
        if (s >= n) s = 2 * (n - 1) - s;
    } else if (pad_mode == DIB_PAD_REPLICATE256) {
        s = s < 0 ? 0 : (s >= n ? n - 1 : s);
    } else {
        if (s < 0 || s >= n) s = -1;
    }
    return s;
}

// ---- Philox4x32-10 (counter-based; same generator family torch's CUDA randn uses) ----
struct Philox {
    uint32_t key[2];
    __device__ Philox(uint64_t seed) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
    }
    __device__ uint4 operator()(uint64_t ctr_lo, uint64_t ctr_hi) const {
        uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// one N(0,1) draw for element `idx` of stream (seed, offset): Box-Muller on two Philox words
__device__ inline float philox_normal(uint64_t seed, uint64_t offset, uint64_t idx) {
    Philox ph(seed);
    const uint4 r = ph(idx >> 1, offset);
    const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.28318530717958647692f * u2, &s, &c);
    return (idx & 1) ? rad * s : rad * c;
}

// ---- epilogue shared by both blur kernels (blur_functions.py:72-74, net_transforms.py:135-139) ----
struct Epilogue {
    int flags;
    float noise_sd, gamma;
    float mean, std;      // of the channel being written
};

__device__ inline float apply_epilogue_f32(float v, const Epilogue& e, float noise) {
    if (e.flags & DIB_EPI_NOISE) v = __fadd_rn(v, __fmul_rn(noise, e.noise_sd));
    if (e.flags & DIB_EPI_CLAMP) v = fminf(fmaxf(v, 0.0f), 1.0f);
    if (e.flags & DIB_EPI_GAMMA) v = powf(v, e.gamma);
    if (e.flags & DIB_EPI_NORMALIZE) v = __fdiv_rn(__fsub_rn(v, e.mean), e.std);
    return v;
}

}  // namespace dib

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
// Tiled-kernel program of one PSF (built by taps.cu, executed by blur_tiled.cu).  The PSF support is cut into
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

// Work-distribution words of the tiled kernel (self-resetting, see blur_tiled.cu)
struct SchedWords {
    unsigned int next_tile;     // next tile to hand out
    unsigned int done_ctas;     // CTAs that have stopped fetching
};
constexpr int kSchedSlots = 4;      // launches that may overlap (DIB_ALGO_OVERLAP) rotate through these (256 bytes are reserved)


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

// Masked tiled sparse-PSF blur for sm_100a: the tiled kernel dib_blur_batch uses for PSFs whose support fits one chunk of
// its program (masked_common.cuh; every low-exposure PSF), float and half I/O.  Larger PSFs take blur_tiled.cu.
//
// Replaces the per-tap `output += torch.roll(pad(image), shift) * w` loop of manual_blur
// (models/blur_functions.py:59-69) -- O(taps) launches and ~7 passes over the padded tensor per tap -- with one
// persistent, warp-specialised launch per batch:
//   * work unit  = one 36 x 448 output tile of one channel of one image; the CTAs (one per SM) take tiles from a
//                  global ticket counter, images with the heaviest PSFs first;
//   * producers  = one warpgroup (4 warps, 56 registers after setmaxnreg.dec).  For every stage (tile x program chunk)
//                  thread t stages row t of tile + halo global -> shared: one TMA bulk copy (cp.async.bulk) for the
//                  16-byte-aligned interior of the row segment, 4-byte cp.async for the <= 3 unaligned floats at each
//                  row end and, shared out over all producer threads, for the reflect-101 border columns, all completing
//                  on the stage's "full" mbarrier.  Rows are independent copies, so reflected rows cost nothing extra
//                  and the reference's native unpitched CHW layout (row pitch 5332 B for W = 1333) needs no repacking;
//                  a row table records each row's 0-3 float skew.  Two stages are in flight; a refill waits on the
//                  stage's "empty" mbarrier.  Half-precision images: the rows land as halves in the second half of
//                  their own bytes and the producer warps widen them in place (issue_stage_half);
//   * consumers  = twelve warps (6 x 2 over the tile, three per SM sub-partition, 152 registers after
//                  setmaxnreg.inc).  Each thread owns a 6-row x 7-column output block (lanes sit 7 floats apart in a
//                  row: odd stride -> conflict-free scalar LDS); rows r and r + 3 share 64-bit accumulators and every
//                  multiply-add is a packed FFMA2.  Taps are consumed as the program built by taps.cu: groups of 4 PSF
//                  columns swept row by row; a rotating 6 x 10 register window of the input slides with the sweep
//                  (one new row per step, loaded into the slot the previous step freed, consumed last).  Absent taps
//                  inside a group are skipped with warp-uniform branches (21 FFMA2 each);
//   * epilogue   = noise / clamp / gamma / (x - mean) / std (blur_functions.py:72-74, net_transforms.py:135-139) fused
//                  on the way out, per warp and without block-level barriers: accumulators -> the warp's private
//                  two-row buffer -> 16-byte vector stores.  Destinations with 16-byte-aligned rows take a lean store;
//                  others stage each row skewed to its global address phase and store the <= 3 unaligned floats at
//                  each row end one by one.
// No tensor cores: the contraction is sparse and data dependent.  Results differ from the exact-order kernel only
// by FMA contraction and tap order (measured <= 4e-7 on [0,1] images; bound 1e-5).
#include "masked_common.cuh"

namespace dib {
namespace mk {

// Shape of the register tiling.  Every thread owns a 2*kR-row x kCC-column output block and treats rows r and r + kR
// as a PAIR: their accumulators share a 64-bit register and every multiply-add is a packed FFMA2 (fma.rn.f32x2: one
// issue slot, two FMAs -- measured 62 TFLOP/s at 8 warps/SM where 3-register FFMA reaches 50), which leaves issue
// slots for the window loads and the sweep control.  The pair's input rows are the same sliding window kR steps
// apart, so one new row per step feeds both halves.  The unrolled sweep body is 2*kR (rotations) x kGroupW (tap
// columns) x kR * kCC FFMA2 of 16 bytes = 8 KB: it has to stay in the instruction cache (a 29 KB body stalled on
// instruction fetch as often as it issued, profiles/round1_notes.md).  kR = 3 keeps a compute thread at ~130
// registers, so 12 compute warps (three per SM sub-partition) fit beside the producer warpgroup; kR = 4 needs 172
// registers, allows only 8 compute warps and measured 5 % slower; kR = 2 with 16 compute warps (104 registers) has
// more warps to hide latency with but a third fewer FMAs per window load and measured 7 % slower.  (Its first build
// hung: setmaxnreg.inc asked for more registers than the CTA's launch-time allocation holds -- see kLaunchRegs.)
constexpr int kDevicePlanShear = 0;      // this kernel's programs are not sheared
constexpr int kR = 3;                   // row pairs per thread
constexpr int kRows = 2 * kR;               // output rows per thread (= rotation period of the register window)
constexpr int kCC = 7;                      // output columns per thread (odd: conflict-free lane stride)
constexpr int kWarpW = 32 * kCC;            // 224 output columns per warp
constexpr int kWarpRows = 6;    // compute warps are arranged kWarpRows x kWarpCols over the tile
constexpr int kWarpCols = 2;
constexpr int kComputeWarps = kWarpRows * kWarpCols;       // a multiple of 4: equal load on the 4 SM sub-partitions
constexpr int kProducerWarps = 4;           // one more warpgroup: every thread stages at most one tile row
constexpr int kThreads = (kComputeWarps + kProducerWarps) * 32;
constexpr int kProducerRegs = 56;   // setmaxnreg budgets; together they must fit the 64K-register file
#ifndef DIB_MK_HALF_PRODUCER_REGS
#define DIB_MK_HALF_PRODUCER_REGS 96
#endif
// the half-precision producers also widen the rows and hold a row's unaligned ends in registers while the copies land: at
// 56 registers that spilled, and the spill stores made the issue phase wait for the global loads
constexpr int kProducerRegsHalf = DIB_MK_HALF_PRODUCER_REGS;
// setmaxnreg moves registers inside the CTA's launch-time allocation (threads x the per-thread count the launch bound
// allows, a multiple of 8); asking for more than the producers hand back blocks forever.
constexpr int kLaunchRegs = (65536 / kThreads) / 8 * 8 > 255 ? 248 : (65536 / kThreads) / 8 * 8;
constexpr int kComputeRegsRaw = (kThreads * kLaunchRegs - kProducerWarps * 32 * kProducerRegs) / (kComputeWarps * 32) / 8 * 8;
constexpr int kComputeRegs = kComputeRegsRaw > 232 ? 232 : kComputeRegsRaw;
constexpr int kComputeRegsHalf = (kThreads * kLaunchRegs - kProducerWarps * 32 * kProducerRegsHalf) / (kComputeWarps * 32) / 8 * 8;
static_assert(kComputeWarps % 4 == 0, "warpgroup-aligned compute warps");
constexpr int kTH = kWarpRows * kRows;      // 36 output rows per tile
constexpr int kTW = kWarpCols * kWarpW;     // 448 output columns per tile
constexpr int kWinW = kCC + kGroupW - 1;    // 10 input columns feed one group
constexpr int kRowsMax = kTH + kChunkHaloRows;                     // 56 staged rows
// Staged row: tile + halo + the row's skew (<= 3 floats for fp32 rows, <= 7 for half rows), rounded to 8 floats so that the
// second half of a row's bytes -- where half-precision rows land before they are widened in place -- is 16-byte aligned.
constexpr int kPitch = ((kTW + kChunkGroups * kGroupW - 1 + 7) + 7) / 8 * 8;   // 480 floats
constexpr int kOutPitch = kWarpW + 4;       // one staged output row of a warp (skew <= 3)
constexpr int kHdrBytes = 64;
constexpr int kAuxBytes = (kChunkDataMax + 15) / 16 * 16;          // segment records + weights of one chunk
constexpr int kRowTabBytes = ((kRowsMax * 4) + 15) / 16 * 16;
constexpr int kTileBytes = kRowsMax * kPitch * 4;
constexpr int kStageBytes = kHdrBytes + kAuxBytes + kRowTabBytes + kTileBytes;
constexpr int kOutBufBytes = kComputeWarps * 2 * kOutPitch * 4;    // two staged rows per compute warp
constexpr int kSmemBytes = 2 * kStageBytes + kOutBufBytes + 64;    // + 4 mbarriers + 2 tile-ticket slots
static_assert(kRowsMax <= kProducerWarps * 32, "one staged row per producer thread");
static_assert(kPitch % 8 == 0 && kPitch >= kTW + kChunkGroups * kGroupW - 1 + 7, "pitch must hold tile + halo + skew");
static_assert(kStageBytes % 16 == 0, "stage must keep 16-byte alignment");
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");

struct TiledImage {
    const void* src;      // float or __half (kernel template), pitches in elements
    void* dst;
    const float* noise;
    int64_t src_rp, src_cp, dst_rp, dst_cp;
    int C, H, W;
    int tiles_x, tiles_y;
    int first_tile;       // tiles of the images before this one
    int psf_index, nchunks;
    int epilogue;
    int zero_pad;         // DIB_PAD_ZERO128: pixels outside the image read as 0 instead of being mirrored
    int aligned_out;      // every destination row starts 16-byte aligned (base, row pitch and channel pitch): lean row store
    int rec0_valid;       // single-chunk PSF: its one chunk record, rebuilt on the host from the PSF summary, rides in the
    ChunkRec rec0;        // kernel parameters, so a tile's first stage need not wait for a load from the program section
    int philox_slot;      // position in the caller's batch (Philox stream id)
    float noise_sd, gamma;
    float mean[4], std[4];
};

struct TiledParams {
    TiledImage img[DIB_MAX_BATCH];
    const uint8_t* prog;          // program sections of the tap set
    int n_images;
    int total_tiles;
    uint64_t philox_seed, philox_offset;
    SchedWords* sched;            // dynamic tile scheduler (tap set buffer): tiles are handed out in index order
    int overlap_prev;             // DIB_ALGO_OVERLAP: do not wait for the grid launched before this one
    const dib_psf_meta* meta_dev; // DIB_ALGO_DEVICE_PLAN: per-PSF summaries on the device decide which images are this kernel's
};

// ---------------------------------------------------------------- optional timeline trace (kernel experiments only)
#ifdef DIB_TRACE
constexpr int kTracePerWarp = 4096;
__device__ unsigned long long g_trace[16 * kTracePerWarp];
__device__ unsigned int g_trace_n[16];
__device__ __forceinline__ void trace_event(int ev, int arg) {
    __shared__ unsigned int pos[16];
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        unsigned int i = pos[w];
        if (i >= (unsigned)kTracePerWarp) i = 0;
        i = (ev == 63) ? 0 : i;
        if (i < (unsigned)kTracePerWarp - 1) {
            g_trace[w * kTracePerWarp + i] = ((unsigned long long)clock64() << 24) | ((unsigned long long)(arg & 0xfff) << 12) | (w << 6) | (unsigned)ev;
            pos[w] = i + 1;
            g_trace_n[w] = i + 1;
        }
    }
}
#define DIB_TRACE_EVENT(ev, arg) trace_event(ev, arg)
#else
#define DIB_TRACE_EVENT(ev, arg)
#endif

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
#ifdef DIB_NO_HINT
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
#endif
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep in hardware, do not spin
}
// Producer-group wait: ONE thread polls the mbarrier (with a back-off), the other 127 block in a named barrier, which costs
// no issue slots -- four producer warps spinning on try_wait took a seventh of the SM's issue slots from the compute warps.
__device__ __forceinline__ void producer_wait(uint64_t* bar, uint32_t parity, int pt) {
    if (pt == 0) {
        uint32_t done = 0;
        while (true) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
            if (done) break;
            __nanosleep(96);
        }
    }
    asm volatile("bar.sync 2, %0;" ::"n"(4 * 32) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// arrive on the mbarrier once all cp.async issued so far by this thread have landed (counts as a normal arrival)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared-memory accesses by 32-bit shared address (no generic-pointer arithmetic in the hot loops)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_entry(uint32_t addr, float& w, int& code) {
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(w), "=r"(code) : "r"(addr));
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }
__device__ __forceinline__ uint2 lds_v2u(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
// four halves (8 bytes) -> four floats at addr .. addr + 15
__device__ __forceinline__ void sts_widened(uint32_t addr, const uint2& h4) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h4.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h4.y));
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
}

__device__ __forceinline__ int reflect101(int v, int n) {
    v = v < 0 ? -v : v;
    v = v >= n ? 2 * (n - 1) - v : v;
    return min(max(v, 0), n - 1);     // clamp only guards rows/cols that feed masked outputs
}

// ---------------------------------------------------------------- stage bookkeeping
// What the producer tells the consumers about a stage (first 64 bytes of the stage's shared memory).
struct __align__(16) StageHdr {
    int tile;           // global tile index; -1 = no more work
    int img, ch, i0, j0;
    int first_chunk, last_chunk;
    int dy_hi, dx_hi, nseg, wsteps;
};
static_assert(sizeof(StageHdr) <= kHdrBytes, "stage header too large");

struct Stage {
    int tile;       // global tile index, -1: none
    int chunk;
    int nchunks;    // chunks of the image's program; 0: the image belongs to another kernel (device-planned launches)
    int img, ch, i0, j0;
    ChunkRec rec;
};

struct StageSmem {
    StageHdr* hdr;
    uint8_t* aux;       // SegRec slots + weight vectors of the chunk
    int* rowtab;        // float offset of image column `cl` inside each staged row
    float* tile;
};

__device__ __forceinline__ StageSmem stage_smem(uint8_t* base, int b) {
    StageSmem s;
    uint8_t* p = base + (size_t)b * kStageBytes;
    s.hdr = reinterpret_cast<StageHdr*>(p);
    s.aux = p + kHdrBytes;
    s.rowtab = reinterpret_cast<int*>(p + kHdrBytes + kAuxBytes);
    s.tile = reinterpret_cast<float*>(p + kHdrBytes + kAuxBytes + kRowTabBytes);
    return s;
}

// Device-planned launches (DIB_ALGO_DEVICE_PLAN): first tile of every image in the ticket sequence, decided from the PSF
// summaries (warp 0, one lane per image, before the kernel's first block barrier).  Images of the other kernels get an
// empty range, so a launch that owns nothing ends after this prologue instead of drawing a ticket per foreign tile.
// Host-planned launches keep reading the host's first_tile from the kernel parameters, image by image and only as far as
// the tile at hand needs (a lane-per-image read of the parameters costs 32 serialised constant-cache misses up front).
__device__ __forceinline__ void tile_table(const TiledParams& p, int my_kind, int* first) {
    if (p.meta_dev != nullptr && threadIdx.x < 32) {
        const int n = threadIdx.x;
        int cnt = 0;
        if (n < p.n_images) {
            const TiledImage& im = p.img[n];
            cnt = im.tiles_x * im.tiles_y * im.C;
            if (p.meta_dev != nullptr) {
                const dib_psf_meta* m = p.meta_dev + im.psf_index;
                const int4 a = __ldg(reinterpret_cast<const int4*>(m));
                const int4 b = __ldg(reinterpret_cast<const int4*>(m) + 1);
                const int2 c = __ldg(reinterpret_cast<const int2*>(m) + 4);
                dib_psf_meta mm;
                mm.count = a.x; mm.prog_chunks = b.y; mm.flags = b.w; mm.prog_group_w = (int16_t)(c.y & 0xffff);
                if (psf_program_kind(mm) != my_kind) cnt = 0;
            }
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (n >= o) incl += t;
        }
        first[n] = incl - cnt;
        if (n == 31) first[DIB_MAX_BATCH] = incl;
    }
}

// `first`: shared-memory table of the first tile of every image, first[DIB_MAX_BATCH] = all tiles (tile_table above), or null
__device__ __forceinline__ void decode_tile(const TiledParams& p, const int* first, int tile, Stage& st) {
    int n = 0;
    if (first != nullptr) {
        while (n + 1 < p.n_images && tile >= first[n + 1]) ++n;
    } else {
        while (n + 1 < p.n_images && tile >= p.img[n + 1].first_tile) ++n;
    }
    const TiledImage& im = p.img[n];
    const int local = tile - (first != nullptr ? first[n] : im.first_tile);
    const int per_ch = im.tiles_x * im.tiles_y;
    st.img = n;
    st.ch = local / per_ch;
    const int rem = local - st.ch * per_ch;
    const int ty = rem / im.tiles_x;
    st.i0 = ty * kTH;
    st.j0 = (rem - ty * im.tiles_x) * kTW;
    st.nchunks = im.nchunks;
    if (p.meta_dev != nullptr) {          // planned on the device: the program's kind and length come from the PSF summary
        const dib_psf_meta* m = p.meta_dev + im.psf_index;
        const int4 a = __ldg(reinterpret_cast<const int4*>(m));            // count | ymin ymax | xmin xmax | sum
        const int4 b = __ldg(reinterpret_cast<const int4*>(m) + 1);        // support | prog_chunks | prog_steps | flags
        const int2 c = __ldg(reinterpret_cast<const int2*>(m) + 4);        // prog_segs | prog_group_w, prog_shear
        dib_psf_meta mm;
        mm.count = a.x; mm.prog_chunks = b.y; mm.flags = b.w; mm.prog_group_w = (int16_t)(c.y & 0xffff);
        st.nchunks = psf_program_kind(mm) == 1 ? mm.prog_chunks : 0;
    }
}

__device__ __forceinline__ ChunkRec load_chunk_rec(const TiledParams& p, int img, int chunk) {
    if (p.img[img].rec0_valid) return p.img[img].rec0;          // single-chunk PSF (chunk == 0)
    const uint8_t* prog = p.prog + (size_t)p.img[img].psf_index * dib::kProgBytes;
    const int4 v = __ldg(reinterpret_cast<const int4*>(prog) + chunk);
    ChunkRec r;
    r.dy_lo = (int16_t)(v.x & 0xffff);
    r.dy_hi = (int16_t)(v.x >> 16);
    r.dx_lo = (int16_t)(v.y & 0xffff);
    r.dx_hi = (int16_t)(v.y >> 16);
    r.nseg = (int16_t)(v.z & 0xffff);
    r.wsteps = (int16_t)(v.z >> 16);
    r.data_off = v.w;
    return r;
}

// Hand the producer group its next tile.  Thread 0 of the group takes a ticket from the global counter and shares
// it through shared memory; the two slots alternate so one named barrier per fetch is enough.
constexpr unsigned kNoTicket = 0xffffffffu;
__device__ __forceinline__ int fetch_tile(const TiledParams& p, const int* first, int* slots, int& nfetch, int pt, unsigned pre = kNoTicket) {
    int* slot = slots + (nfetch & 1);
    if (pt == 0) {
        // the first tile of a CTA is its block index (no round trip to the counter before the first load can go out);
        // tickets from the counter follow after the gridDim.x tiles handed out that way.  `pre`: a ticket thread 0 drew
        // earlier (the atomic's round trip then overlaps the work in between)
        const unsigned t = nfetch == 0 ? blockIdx.x : gridDim.x + (pre != kNoTicket ? pre : atomicAdd(&p.sched->next_tile, 1u));
        *slot = t < (unsigned)(first != nullptr ? first[DIB_MAX_BATCH] : p.total_tiles) ? (int)t : -1;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory");
    ++nfetch;
    return *slot;
}

// successor of a stage: next chunk of the same tile, else chunk 0 of the next tile the scheduler hands out
__device__ __forceinline__ void next_stage(const TiledParams& p, const int* first, const Stage& cur, Stage& nx, int* slots, int& nfetch,
                                           int pt, unsigned pre = kNoTicket) {
    if (cur.tile < 0) {
        nx.tile = -1;
        return;
    }
    if (cur.chunk + 1 < cur.nchunks) {
        nx = cur;
        nx.chunk = cur.chunk + 1;
    } else {
        do {        // device-planned launches: tiles of images that belong to another kernel are skipped
            nx.tile = fetch_tile(p, first, slots, nfetch, pt, pre);
            pre = kNoTicket;
            if (nx.tile < 0) return;
            nx.chunk = 0;
            decode_tile(p, first, nx.tile, nx);
        } while (nx.nchunks == 0);
    }
    nx.rec = load_chunk_rec(p, nx.img, nx.chunk);
}

// Producer group (4 warps): issue every load of one stage.  Thread t owns staged row t: it places the row (skewed so
// that the 16-byte-aligned interior of the in-image segment lands 16-byte aligned), moves that interior with one TMA
// bulk copy and fetches the <= 3 + 3 unaligned end floats with 4-byte cp.async.  Tiles that reach past the left /
// right image border then get their reflect-101 columns, each warp covering the rows its own lanes placed.  In
// zero-padding mode rows and columns outside the image are stored as zeros instead (plain shared-memory stores, which
// the thread's own arrive on the stage barrier publishes to the consumers).
__device__ __forceinline__ void issue_stage(const TiledParams& p, const Stage& st, const StageSmem& sm, uint64_t* bar, int pt,
                                            uint64_t* empty_bar, uint32_t empty_parity, bool wait_empty) {
    const TiledImage& im = p.img[st.img];
    const int lane = pt & 31, pw = pt >> 5;
    const int sr = lane * kProducerWarps + pw;                             // rows interleave over the producer warps
    const int rt = st.i0 - st.rec.dy_hi;                                  // image row of staged row 0
    const int nrows = kTH + st.rec.dy_hi - st.rec.dy_lo;
    const int cl = st.j0 - st.rec.dx_hi;                                  // image column of staged column 0
    const int cr = min(st.j0 + kTW, im.W) - 1 - st.rec.dx_lo;            // last staged image column
    const int xa = max(cl, 0), xb1 = min(cr, im.W - 1) + 1;              // in-image part [xa, xb1)
    const float* plane = static_cast<const float*>(im.src) + (int64_t)st.ch * im.src_cp;
    // Where this thread's row comes from and where it goes: pure arithmetic, done before the wait for the buffer so that
    // the copies go out as soon as the consumers release it.
    const float* gp = plane;
    int ro = sr * kPitch;
    int xa_al = cl, xb_al = cl;              // nothing inside the image: every column is mirrored
    const bool in_rows = sr < nrows;
    const bool zero_row = im.zero_pad && (rt + sr < 0 || rt + sr >= im.H);
    if (in_rows && !zero_row) {
        gp = plane + (int64_t)reflect101(rt + sr, im.H) * im.src_rp;
        if (xb1 > xa) {
            const uint32_t rowphase = (uint32_t)(reinterpret_cast<uintptr_t>(gp) >> 2);
            xa_al = xa + (int)((0u - (rowphase + (uint32_t)xa)) & 3u);
            xb_al = xb1 - (int)((rowphase + (uint32_t)xb1) & 3u);
            if (xb_al <= xa_al) xa_al = xb_al = xa;      // segment shorter than one aligned quad
        }
        ro += (int)((uint32_t)(cl - xa_al) & 3u);       // skew: makes (xa_al - cl + skew) a multiple of 4
    }
    float* drow = sm.tile + ro - cl;                     // drow[col] addresses image column col
    const uint32_t nb_row = (in_rows && !zero_row) ? (uint32_t)(xb_al - xa_al) * 4u : 0u;
    if (wait_empty) producer_wait(empty_bar, empty_parity, pt);  // consumers released the stage that used this buffer
    fence_proxy_async();   // order the consumers' generic-proxy reads of this buffer before the async-proxy writes
    if (pt == 0) {
        StageHdr h;
        h.tile = st.tile; h.img = st.img; h.ch = st.ch; h.i0 = st.i0; h.j0 = st.j0;
        h.first_chunk = (st.chunk == 0);
        h.last_chunk = (st.chunk + 1 == st.nchunks);
        h.dy_hi = st.rec.dy_hi; h.dx_hi = st.rec.dx_hi; h.nseg = st.rec.nseg; h.wsteps = st.rec.wsteps;
        *sm.hdr = h;
    }
    uint32_t bytes = nb_row;
    if (in_rows && zero_row) {
        for (int col = cl; col <= cr; ++col) drow[col] = 0.0f;
    } else if (in_rows) {
        if (nb_row) tma_bulk_g2s(drow + xa_al, gp + xa_al, nb_row, bar);
        for (int col = xa; col < xa_al; ++col) cp_async_4(drow + col, gp + col);      // unaligned head
        for (int col = xb_al; col < xb1; ++col) cp_async_4(drow + col, gp + col);     // unaligned tail
    }
    if (sr < kRowsMax) sm.rowtab[sr] = ro;   // rows past a partial tile are read (results discarded): offsets stay in range
    if (pt == 32) {
        const uint32_t nb = (uint32_t)(kChunkSegBytes + kStepBytes * (st.rec.wsteps + 1));
        tma_bulk_g2s(sm.aux, p.prog + (size_t)im.psf_index * dib::kProgBytes + st.rec.data_off, nb, bar);
        bytes += nb;
    }
    if (cl < 0 || cr >= im.W) {
        // border columns [cl, xa) and [xb1, cr] of every staged row: mirrored pixels (4-byte cp.async each), or zeros in
        // zero-padding mode.  All producer threads share the (row, column) pairs; the row table tells where a row sits.
        asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory");
        // (a chunk whose taps all point past the right edge of a last column tile stages no image column at all: cl >= W;
        //  the mirrored columns then start at cl, not at the first column beyond the image)
        const int rb = max(xb1, cl);
        const int nleft = min(xa, cr + 1) - cl, ncols = nleft + cr + 1 - rb;
        const int total = nrows * ncols;
        for (int idx = pt; idx < total; idx += kProducerWarps * 32) {
            const int r2 = idx / ncols, k = idx - r2 * ncols;
            const int col = k < nleft ? cl + k : rb + (k - nleft);
            const int irow = rt + r2;
            float* d = sm.tile + sm.rowtab[r2] - cl + col;
            if (im.zero_pad) {
                if (irow >= 0 && irow < im.H) *d = 0.0f;           // rows outside the image are already all zeros
            } else {
                cp_async_4(d, plane + (int64_t)reflect101(irow, im.H) * im.src_rp + reflect101(col, im.W));
            }
        }
    }
    mbar_arrive_expect_tx(bar, bytes);
    cp_async_mbar_arrive(bar);
}

// Half-precision I/O: the same stage built from __half rows.  Thread t places row t's 16-byte-aligned interior, as
// halves, in the SECOND half of the row's bytes with one TMA bulk copy (skewed so that global and shared addresses
// agree mod 16 bytes); when the stage's copies have landed (`landed` barrier) the producer warps widen the rows in place,
// front to back -- the float written for element p ends at byte 4p + 4, never past a half still to be read at byte
// 2 * kPitch + 2p' of a later group -- and then fill in what TMA cannot move: the <= 7 + 7 unaligned end elements and the mirrored (or zero) border columns,
// read straight from global memory.  The consumers see exactly the fp32 tile of the float path.
__device__ __forceinline__ void issue_stage_half(const TiledParams& p, const Stage& st, const StageSmem& sm, uint64_t* landed,
                                                 uint32_t landed_parity, uint64_t* full, int pt) {
    const TiledImage& im = p.img[st.img];
    const int lane = pt & 31, pw = pt >> 5;
    const int sr = lane * kProducerWarps + pw;
    const int rt = st.i0 - st.rec.dy_hi;
    const int nrows = kTH + st.rec.dy_hi - st.rec.dy_lo;
    const int cl = st.j0 - st.rec.dx_hi;
    const int cr = min(st.j0 + kTW, im.W) - 1 - st.rec.dx_lo;
    const int xa = max(cl, 0), xb1 = min(cr, im.W - 1) + 1;
    const __half* plane = static_cast<const __half*>(im.src) + (int64_t)st.ch * im.src_cp;
    fence_proxy_async();
    if (pt == 0) {
        StageHdr h;
        h.tile = st.tile; h.img = st.img; h.ch = st.ch; h.i0 = st.i0; h.j0 = st.j0;
        h.first_chunk = (st.chunk == 0);
        h.last_chunk = (st.chunk + 1 == st.nchunks);
        h.dy_hi = st.rec.dy_hi; h.dx_hi = st.rec.dx_hi; h.nseg = st.rec.nseg; h.wsteps = st.rec.wsteps;
        *sm.hdr = h;
    }
    uint32_t bytes = 0;
    const __half* gp = plane;
    int ro = sr * kPitch;
    int xa_al = cl, xb_al = cl;
    const bool in_rows = sr < nrows;
    const bool zero_row = im.zero_pad && (rt + sr < 0 || rt + sr >= im.H);
    if (in_rows && !zero_row) {
        gp = plane + (int64_t)reflect101(rt + sr, im.H) * im.src_rp;
        if (xb1 > xa) {
            const uint32_t rowphase = (uint32_t)(reinterpret_cast<uintptr_t>(gp) >> 1);       // in halves
            xa_al = xa + (int)((0u - (rowphase + (uint32_t)xa)) & 7u);
            xb_al = xb1 - (int)((rowphase + (uint32_t)xb1) & 7u);
            // no whole aligned group of 8 inside (a last column tile that holds few image columns: the span is then at most
            // 7 + 7 elements): the first seven travel as the "head", the rest as the "tail" -- both are register arrays of 7
            if (xb_al <= xa_al) xa_al = xb_al = min(xa + 7, xb1);
        }
        ro += (int)((uint32_t)(cl - xa_al) & 7u);       // skew: (xa_al - cl + skew) is a multiple of 8
        const uint32_t nb = (uint32_t)(xb_al - xa_al) * 2u;
        if (nb) {
            __half* hrow = reinterpret_cast<__half*>(sm.tile + sr * kPitch) + kPitch;      // the row's second half
            tma_bulk_g2s(hrow + (ro - sr * kPitch) + (xa_al - cl), gp + xa_al, nb, landed);
            bytes += nb;
        }
    }
    if (sr < kRowsMax) sm.rowtab[sr] = ro;
    if (pt == 32) {
        const uint32_t nb = (uint32_t)(kChunkSegBytes + kStepBytes * (st.rec.wsteps + 1));
        tma_bulk_g2s(sm.aux, p.prog + (size_t)im.psf_index * dib::kProgBytes + st.rec.data_off, nb, landed);
        bytes += nb;
    }
    mbar_arrive_expect_tx(landed, bytes);
    // the <= 7 + 7 unaligned end elements of the row: independent global loads, in flight while the bulk copies land
    __half head[7], tail[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        head[j] = (in_rows && !zero_row && xa + j < xa_al) ? gp[xa + j] : __half(0.0f);
        tail[j] = (in_rows && !zero_row && xb_al + j < xb1) ? gp[xb_al + j] : __half(0.0f);
    }
    DIB_TRACE_EVENT(10, 0);
    producer_wait(landed, landed_parity, pt);
    DIB_TRACE_EVENT(11, 0);
    float* drow = sm.tile + ro - cl;                  // drow[col] addresses image column col
    // Widen the aligned interiors in place.  A warp takes the rows its own lanes placed, one row at a time, a lane per
    // group of 8 elements: every lane first reads its 16 bytes of halves, then -- after a warp barrier -- writes its 32
    // bytes of floats.  Within a round of 32 groups the floats land on halves that the round has already read (a float
    // group ends at byte 32 g + 32, the half group it may reach starts at 2 * kPitch + 16 g); rounds ascend along the row.
    {
        static_assert(kPitch / 4 <= 128, "four chunks per lane cover a staged row");
        static_assert(kRowsMax * kPitch < (1 << 20), "row offset and group count share one shuffled word");
        const int my_n8 = (in_rows && !zero_row) ? (xb_al - xa_al) >> 3 : 0;
        const int packed = (ro - cl + xa_al) | (my_n8 << 20);             // float index of the first aligned element | groups
        // A lane widens four 4-element chunks per row, 32 chunks apart: every load (8 bytes per lane) and every store (16
        // bytes per lane) of the warp covers one contiguous span, so neither has bank conflicts.
        const uint32_t tile_b = smem_u32(sm.tile);
        uint32_t hb = tile_b + 2u * kPitch * (uint32_t)(pw + 1) + 8u * (uint32_t)lane;      // row pw's halves; + 2 * off
        const uint32_t fb = tile_b + 16u * (uint32_t)lane;                                    // + 4 * off
        const int nl = (nrows - pw + kProducerWarps - 1) / kProducerWarps;                   // rows this warp placed
        // two rows per round: the second row's loads are in flight while the first row's latency passes
        for (int l = 0; l < nl; l += 2, hb += 4u * kPitch * kProducerWarps) {
            uint2 hv[2][4];
            uint32_t f0[2];
            int n4[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int pk = __shfl_sync(0xffffffffu, packed, min(l + q, 31));
                const uint32_t off = (uint32_t)(pk & 0xfffff);
                n4[q] = (l + q < nl) ? (pk >> 20) * 2 : 0;                   // 4-element chunks of the row (<= 120)
                const uint32_t h0 = hb + 2u * off + (uint32_t)q * (2u * kPitch * kProducerWarps);
                f0[q] = fb + 4u * off;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    hv[q][j] = make_uint2(0u, 0u);
                    if (lane + 32 * j < n4[q]) hv[q][j] = lds_v2u(h0 + 256u * j);
                }
            }
            __syncwarp();                                                  // both rows are read before any of them is written
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (lane + 32 * j < n4[q]) sts_widened(f0[q] + 512u * j, hv[q][j]);
        }
        __syncwarp();
    }
    DIB_TRACE_EVENT(12, 0);
    if (in_rows && zero_row) {
        for (int col = cl; col <= cr; ++col) drow[col] = 0.0f;
    } else if (in_rows) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (xa + j < xa_al) drow[xa + j] = __half2float(head[j]);
            if (xb_al + j < xb1) drow[xb_al + j] = __half2float(tail[j]);
        }
    }
    if (cl < 0 || cr >= im.W) {
        // border columns [cl, xa) and [xb1, cr] of every staged row: mirrored pixels, or zeros in zero-padding mode.  All
        // producer threads share the (row, column) pairs; the row table written above tells where each row sits.
        asm volatile("bar.sync 1, %0;" ::"n"(kProducerWarps * 32) : "memory");
        // (a chunk whose taps all point past the right edge of a last column tile stages no image column at all: cl >= W;
        //  the mirrored columns then start at cl, not at the first column beyond the image)
        const int rb = max(xb1, cl);
        const int nleft = min(xa, cr + 1) - cl, ncols = nleft + cr + 1 - rb;
        const int total = nrows * ncols;
#pragma unroll 4
        for (int idx = pt; idx < total; idx += kProducerWarps * 32) {
            const int r2 = idx / ncols, k = idx - r2 * ncols;
            const int col = k < nleft ? cl + k : rb + (k - nleft);
            const int irow = rt + r2;
            float v = 0.0f;
            bool write = true;
            float* rowp = sm.tile + sm.rowtab[r2] - cl;        // rowp[c] is image column c of the staged (mirrored) row
            if (im.zero_pad) {
                write = irow >= 0 && irow < im.H;            // rows outside the image are already all zeros
            } else {
                // a reflect-101 pixel beside the image is another pixel of the same staged row (widened above): no trip to
                // global memory -- unless the tile holds so few image columns that the mirror lies left of what was staged
                const int mc = reflect101(col, im.W);
                v = (mc >= xa && mc < xb1) ? rowp[mc] : __half2float(plane[(int64_t)reflect101(irow, im.H) * im.src_rp + mc]);
            }
            if (write) rowp[col] = v;
        }
    }
    DIB_TRACE_EVENT(16, 0);
    mbar_arrive(full);          // release: this thread's shared-memory writes are visible to whoever observes the phase
}

// ---------------------------------------------------------------- compute
// Packed pair of fp32 values: .x belongs to output row r of the thread's block, .y to row r + kR.
__device__ __forceinline__ float2 ffma2(float w, float2 x, float2 a) {
    unsigned long long d, ww, xx, aa;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x.x), "f"(x.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(ww), "l"(xx), "l"(aa));     // SASS: FFMA2 with a scalar weight operand
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(d));
    return r;
}

// The window holds the kRows input rows of the current step.  Logical row q sits in slot (q - U) mod kRows at
// rotation U; slot j < kR is win[j].x, slot j + kR is win[j].y, so the two rows of a pair (q and q + kR) always share
// one 64-bit register -- in swapped halves for half of the rotations, which FFMA2's operand swizzle absorbs.
template <int U, int Q>
__device__ __forceinline__ float2 window_pair(const float2 (&win)[kR][kWinW], int k) {
    constexpr int slot = ((Q - U) % kRows + kRows) % kRows;
    if constexpr (slot < kR)
        return win[slot][k];
    else
        return make_float2(win[slot - kR][k].y, win[slot - kR][k].x);
}

// One tap of a sweep step: weight w multiplies the window shifted by E columns.  Pair 0 contains the row loaded at
// the start of this step and is consumed last, so the FMAs on the older rows cover that load's latency.
template <int U, int E>
__device__ __forceinline__ void fma_tap(float2 (&acc)[kR][kCC], const float2 (&win)[kR][kWinW], const float w) {
#pragma unroll
    for (int rr = 1; rr <= kR; ++rr) {
        const int r = rr % kR;
#pragma unroll
        for (int c = 0; c < kCC; ++c) {
            float2 x;
            // constexpr dispatch on the pair index (r is a compile-time constant after unrolling)
            if (r == 0) x = window_pair<U, 0>(win, c - E + kGroupW - 1);
            else if (r == 1) x = window_pair<U, 1>(win, c - E + kGroupW - 1);
            else if (r == 2) x = window_pair<U, 2>(win, c - E + kGroupW - 1);
            else x = window_pair<U, 3>(win, c - E + kGroupW - 1);
            acc[r][c] = ffma2(w, x, acc[r][c]);
        }
    }
}
static_assert(kR <= 4, "fma_tap dispatches on at most 4 row pairs");

// load one input row into window slot SLOT
template <int SLOT>
__device__ __forceinline__ void load_row(float2 (&win)[kR][kWinW], uint32_t addr) {
#pragma unroll
    for (int k = 0; k < kWinW; ++k) {
        if constexpr (SLOT < kR)
            win[SLOT][k].x = lds_f32(addr + 4 * k);
        else
            win[SLOT - kR][k].y = lds_f32(addr + 4 * k);
    }
}

// Step s of a segment sweep, s mod kRows == U: fetch the new top row into the slot the previous step freed and the
// NEXT step's weight vector (kRows is even, so the two weight registers simply alternate), then accumulate the taps
// present in this step's vector; absent taps are skipped with warp-uniform branches.
// Step s of a segment sweep, s mod kRows == U: fetch the new top row into the slot the previous step freed and the
// NEXT step's weight vector (kRows is even, so the two weight registers simply alternate), then accumulate the taps
// present in this step's vector; absent taps are skipped with warp-uniform branches.  (A fall-through chain driven by
// per-step first/last codes was tried: the compiler's nested reconvergence scaffolding made it slower.)
// the kGroupW weights of one step
struct WeightVec {
    float4 q[kGroupW / 4];
};
__device__ __forceinline__ WeightVec lds_weights(uint32_t addr) {
    WeightVec w;
#pragma unroll
    for (int i = 0; i < kGroupW / 4; ++i) w.q[i] = lds_v4(addr + 16u * i);
    return w;
}

template <int U>
__device__ __forceinline__ bool sweep_step(float2 (&acc)[kR][kCC], float2 (&win)[kR][kWinW], uint32_t tile_cb, uint32_t rowtab,
                                           int sr0, int& s, int nsteps, uint32_t& wp, int& ro_next, WeightVec (&wv)[2]) {
    if (s > 0) load_row<(kRows - U) % kRows>(win, tile_cb + 4u * (uint32_t)ro_next);
    ro_next = lds_s32(rowtab + 4u * (uint32_t)max(sr0 - (s + 1), 0));    // row offset of the next step, one step ahead
    wp += kStepBytes;
    wv[(U + 1) & 1] = lds_weights(wp);                                    // weights of step s + 1 (zero vector past the end)
    const WeightVec& w = wv[U & 1];
    if (w.q[0].x != 0.0f) fma_tap<U, 0>(acc, win, w.q[0].x);
    if (w.q[0].y != 0.0f) fma_tap<U, 1>(acc, win, w.q[0].y);
    if (w.q[0].z != 0.0f) fma_tap<U, 2>(acc, win, w.q[0].z);
    if (w.q[0].w != 0.0f) fma_tap<U, 3>(acc, win, w.q[0].w);
    if constexpr (kGroupW == 8) {
        if (w.q[kGroupW / 4 - 1].x != 0.0f) fma_tap<U, 4>(acc, win, w.q[kGroupW / 4 - 1].x);
        if (w.q[kGroupW / 4 - 1].y != 0.0f) fma_tap<U, 5>(acc, win, w.q[kGroupW / 4 - 1].y);
        if (w.q[kGroupW / 4 - 1].z != 0.0f) fma_tap<U, 6>(acc, win, w.q[kGroupW / 4 - 1].z);
        if (w.q[kGroupW / 4 - 1].w != 0.0f) fma_tap<U, 7>(acc, win, w.q[kGroupW / 4 - 1].w);
    }
    ++s;
    return s < nsteps;
}

// step 0 needs all kRows rows: logical row q -> slot q
template <int Q>
__device__ __forceinline__ void fill_window(float2 (&win)[kR][kWinW], uint32_t tile_cb, uint32_t rowtab, int sr0) {
    load_row<Q>(win, tile_cb + 4u * (uint32_t)lds_s32(rowtab + 4u * (uint32_t)(sr0 + Q)));
    if constexpr (Q + 1 < kRows) fill_window<Q + 1>(win, tile_cb, rowtab, sr0);
}

// kRows consecutive steps = one full rotation of the window registers
template <int U>
struct SweepRound {
    __device__ __forceinline__ static bool run(float2 (&acc)[kR][kCC], float2 (&win)[kR][kWinW], uint32_t tile_cb, uint32_t rowtab,
                                               int sr0, int& s, int nsteps, uint32_t& wp, int& ro_next, WeightVec (&wv)[2]) {
        if (!sweep_step<U>(acc, win, tile_cb, rowtab, sr0, s, nsteps, wp, ro_next, wv)) return false;
        if constexpr (U + 1 < kRows)
            return SweepRound<U + 1>::run(acc, win, tile_cb, rowtab, sr0, s, nsteps, wp, ro_next, wv);
        else
            return true;
    }
};

__device__ __forceinline__ void compute_chunk(float2 (&acc)[kR][kCC], uint32_t stage_addr, int nseg, int dy_hi, int dx_hi,
                                              int wrow, int wcol) {
    const int lane = threadIdx.x & 31;
    const uint32_t aux = stage_addr + kHdrBytes;
    const uint32_t rowtab = aux + kAuxBytes;
    const uint32_t tile = rowtab + kRowTabBytes;
    // The two compute warps that share an SM sub-partition (warp ids w and w + 4) walk the segments in opposite
    // orders: otherwise they run the same instruction sequence in lockstep and their load phases coincide.
    const bool reverse = (((threadIdx.x >> 5) - kProducerWarps) & 4) != 0;
#pragma unroll 1
    for (int sgi = 0; sgi < nseg; ++sgi) {
        const int sg = reverse ? nseg - 1 - sgi : sgi;
        int raw0, raw1;         // SegRec {dx0, dy0 | nsteps, woff} as two words
        lds_entry(aux + 8u * (uint32_t)sg, reinterpret_cast<float&>(raw0), raw1);
        DIB_TRACE_EVENT(2, raw1 & 0xffff);
        const int seg_dx0 = (int)(short)(raw0 & 0xffff), seg_dy0 = raw0 >> 16;
        const int nsteps = (int)(short)(raw1 & 0xffff), seg_woff = raw1 >> 16;
        const int colbase = wcol * kWarpW + kCC * lane - seg_dx0 - (kGroupW - 1) + dx_hi;
        const uint32_t tile_cb = tile + 4u * (uint32_t)colbase;
        const int sr0 = wrow * kRows - seg_dy0 + dy_hi;   // staged row of output row 0 at step 0
        uint32_t wp = aux + kChunkSegBytes + (uint32_t)kStepBytes * (uint32_t)seg_woff;
        WeightVec wv[2];
        wv[0] = lds_weights(wp);
        float2 win[kR][kWinW];
        fill_window<0>(win, tile_cb, rowtab, sr0);
        int ro_next = 0;
        int s = 0;
        DIB_TRACE_EVENT(6, sg);
#pragma unroll 1
        while (SweepRound<0>::run(acc, win, tile_cb, rowtab, sr0, s, nsteps, wp, ro_next, wv)) {
        }
        DIB_TRACE_EVENT(3, sg);
    }
}

// ---------------------------------------------------------------- epilogue + store (per warp, no block barrier)
// One staged row -> global with the fused epilogue (noise / clamp / gamma / normalize).  Kept out of line: the
// common no-epilogue path below stays small and the register-tile code is not replicated around powf / Philox.
// `srow` is the 16-byte-aligned start of the staged row; element x of the row sits at srow[skew + x].
__device__ __noinline__ void store_row_epilogue(const float* srow, float* g, const float* nz_row, int wv, int skew, Epilogue ep,
                                                uint64_t seed, uint64_t stream, uint64_t pbase) {
    const int lane = threadIdx.x & 31;
    for (int k = lane; 4 * k - skew < wv; k += 32) {
        const int x0 = 4 * k - skew;
        const float4 v = *reinterpret_cast<const float4*>(srow + 4 * k);
        float o[4] = {v.x, v.y, v.z, v.w};
        const bool whole = (x0 >= 0) && (x0 + 3 < wv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + j;
            if (x >= 0 && x < wv) {
                float nz = 0.f;
                if (ep.flags & DIB_EPI_NOISE) nz = nz_row ? nz_row[x] : philox_normal(seed, stream, pbase + x);
                o[j] = apply_epilogue_f32(o[j], ep, nz);
                if (!whole) g[x] = o[j];
            }
        }
        if (whole) *reinterpret_cast<float4*>(g + x0) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// One staged row -> global without epilogue.  Element x of the row sits at srow + 4 * (skew + x).  Aligned quads
// k0 .. k1-1 lie wholly inside [0, wv) and go out as 16-byte stores; the <= 3 elements before the first and after the
// last whole quad are stored one per lane.  (A specialised path for full-width rows with lane-dependent roles was
// measured 4 % slower: the divergence costs more than the saved arithmetic.)
template <bool kAffine>
__device__ __forceinline__ void store_row_plain(float* g, uint32_t srow, int skew, int wv, int lane, float scale, float shift) {
    const int k0 = (skew + 3) >> 2, k1 = (wv + skew) >> 2;
    const int head = min(4 * k0 - skew, wv), tail0 = max(4 * k1 - skew, head);
#pragma unroll
    for (int it = 0; it < (kWarpW + 3 + 127) / 128; ++it) {
        const int k = lane + 32 * it;
        if (k >= k0 && k < k1) {
            float4 v = lds_v4(srow + 16u * (uint32_t)k);
            if (kAffine) {
                v.x = fmaf(v.x, scale, shift);
                v.y = fmaf(v.y, scale, shift);
                v.z = fmaf(v.z, scale, shift);
                v.w = fmaf(v.w, scale, shift);
            }
            *reinterpret_cast<float4*>(g + (4 * k - skew)) = v;
        }
    }
    const int x = lane < head ? lane : tail0 + (lane - head);
    if (x < wv && (lane < head || x >= tail0)) {
        float v = lds_f32(srow + 4u * (uint32_t)(skew + x));
        if (kAffine) v = fmaf(v, scale, shift);
        g[x] = v;
    }
}

// Destination rows that all start 16-byte aligned (a pitched output, e.g. the padded batch or the wrapper's own
// allocations) need no per-row skew, no head elements and -- except in the last column tile -- no tail: a third of the
// general store's instructions.  Row q of the pass sits unskewed at obuf + q * kOutPitch.
template <bool kAffine>
__device__ __forceinline__ void store_rows_aligned(const TiledImage& im, int ch, int row0, int col0, float2 (&acc)[kR][kCC],
                                                   uint32_t obuf, float scale, float shift) {
    const int lane = threadIdx.x & 31;
    const int wv = min(kWarpW, im.W - col0);
    const int nrows = min(kRows, im.H - row0);
    const int nq = wv >> 2, tail = wv & 3;                     // whole quads; leftover elements of the last column tile
    const bool q0 = lane < nq, q1 = lane + 32 < nq, qt = lane < tail;
    float* g = static_cast<float*>(im.dst) + (int64_t)ch * im.dst_cp + (int64_t)row0 * im.dst_rp + col0;
    const uint32_t sts0 = obuf + 4u * (uint32_t)(kCC * lane), lds0 = obuf + 16u * (uint32_t)lane;
#pragma unroll
    for (int r = 0; r < kRows; r += 2) {
        if (r >= nrows) break;                                  // warp-uniform
#pragma unroll
        for (int c = 0; c < kCC; ++c) {
            sts_f32(sts0 + 4u * c, r < kR ? acc[r % kR][c].x : acc[r % kR][c].y);
            sts_f32(sts0 + 4u * (kOutPitch + c), r + 1 < kR ? acc[(r + 1) % kR][c].x : acc[(r + 1) % kR][c].y);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (r + q < nrows) {
                float* grow = g + (int64_t)q * im.dst_rp;
                const uint32_t l = lds0 + 4u * (uint32_t)(q * kOutPitch);
                if (q0) {
                    float4 v = lds_v4(l);
                    if (kAffine) { v.x = fmaf(v.x, scale, shift); v.y = fmaf(v.y, scale, shift); v.z = fmaf(v.z, scale, shift); v.w = fmaf(v.w, scale, shift); }
                    *reinterpret_cast<float4*>(grow + 4 * lane) = v;
                }
                if (q1) {
                    float4 v = lds_v4(l + 512u);
                    if (kAffine) { v.x = fmaf(v.x, scale, shift); v.y = fmaf(v.y, scale, shift); v.z = fmaf(v.z, scale, shift); v.w = fmaf(v.w, scale, shift); }
                    *reinterpret_cast<float4*>(grow + 4 * lane + 128) = v;
                }
                if (tail != 0 && qt) {
                    float v = lds_f32(obuf + 4u * (uint32_t)(q * kOutPitch + 4 * nq + lane));
                    if (kAffine) v = fmaf(v, scale, shift);
                    grow[4 * nq + lane] = v;
                }
            }
        }
        __syncwarp();
        g += 2 * im.dst_rp;
    }
}

// Half-precision destination with 16-byte-aligned rows: eight values per lane and row, rounded to half once.
template <bool kAffine>
__device__ __forceinline__ void store_rows_aligned_half(const TiledImage& im, int ch, int row0, int col0, float2 (&acc)[kR][kCC],
                                                        uint32_t obuf, float scale, float shift) {
    const int lane = threadIdx.x & 31;
    const int wv = min(kWarpW, im.W - col0);
    const int nrows = min(kRows, im.H - row0);
    const int n8 = wv >> 3, tail = wv & 7;
    const bool q0 = lane < n8, qt = lane < tail;
    __half* g = static_cast<__half*>(im.dst) + (int64_t)ch * im.dst_cp + (int64_t)row0 * im.dst_rp + col0;
    const uint32_t sts0 = obuf + 4u * (uint32_t)(kCC * lane), lds0 = obuf + 32u * (uint32_t)lane;
#pragma unroll
    for (int r = 0; r < kRows; r += 2) {
        if (r >= nrows) break;
#pragma unroll
        for (int c = 0; c < kCC; ++c) {
            sts_f32(sts0 + 4u * c, r < kR ? acc[r % kR][c].x : acc[r % kR][c].y);
            sts_f32(sts0 + 4u * (kOutPitch + c), r + 1 < kR ? acc[(r + 1) % kR][c].x : acc[(r + 1) % kR][c].y);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (r + q < nrows) {
                __half* grow = g + (int64_t)q * im.dst_rp;
                const uint32_t l = lds0 + 4u * (uint32_t)(q * kOutPitch);
                if (q0) {
                    float4 a = lds_v4(l), b = lds_v4(l + 16u);
                    if (kAffine) {
                        a.x = fmaf(a.x, scale, shift); a.y = fmaf(a.y, scale, shift); a.z = fmaf(a.z, scale, shift); a.w = fmaf(a.w, scale, shift);
                        b.x = fmaf(b.x, scale, shift); b.y = fmaf(b.y, scale, shift); b.z = fmaf(b.z, scale, shift); b.w = fmaf(b.w, scale, shift);
                    }
                    uint4 o;
                    *reinterpret_cast<__half2*>(&o.x) = __floats2half2_rn(a.x, a.y);
                    *reinterpret_cast<__half2*>(&o.y) = __floats2half2_rn(a.z, a.w);
                    *reinterpret_cast<__half2*>(&o.z) = __floats2half2_rn(b.x, b.y);
                    *reinterpret_cast<__half2*>(&o.w) = __floats2half2_rn(b.z, b.w);
                    *reinterpret_cast<uint4*>(grow + 8 * lane) = o;
                }
                if (tail != 0 && qt) {
                    float v = lds_f32(obuf + 4u * (uint32_t)(q * kOutPitch + 8 * n8 + lane));
                    if (kAffine) v = fmaf(v, scale, shift);
                    grow[8 * n8 + lane] = __float2half_rn(v);
                }
            }
        }
        __syncwarp();
        g += 2 * im.dst_rp;
    }
}

// Epilogue variants of the kernel: none; normalize only, applied as one FMA per pixel, x * (1/std) - mean/std (the
// tiled kernel is not the bit-exact path, and an IEEE division per pixel would double its store cost); everything else.
constexpr int kEpiNone = 0, kEpiAffine = 1, kEpiGeneral = 2;

template <int kEpi, bool kHalf>
__device__ __forceinline__ void store_rows(const TiledParams& p, const TiledImage& im, int ch, int row0, int col0,
                                           float2 (&acc)[kR][kCC], uint32_t obuf) {
    const int lane = threadIdx.x & 31;
    const int wv = min(kWarpW, im.W - col0);
    Epilogue ep;
    ep.flags = im.epilogue;
    ep.noise_sd = im.noise_sd;
    ep.gamma = im.gamma;
    ep.mean = im.mean[ch & 3];
    ep.std = im.std[ch & 3];
    const bool norm = (im.epilogue & DIB_EPI_NORMALIZE) != 0;
    const float aff_scale = norm ? 1.0f / ep.std : 1.0f, aff_shift = norm ? -ep.mean / ep.std : 0.0f;
    if constexpr (kHalf) {
        // the launcher only admits half images whose destination rows are 16-byte aligned and whose epilogue is affine
        store_rows_aligned_half<kEpi == kEpiAffine>(im, ch, row0, col0, acc, obuf, aff_scale, aff_shift);
        return;
    }
    if (kEpi != kEpiGeneral && im.aligned_out) {
        store_rows_aligned<kEpi == kEpiAffine>(im, ch, row0, col0, acc, obuf, aff_scale, aff_shift);
        return;
    }
    float* g = static_cast<float*>(im.dst) + (int64_t)ch * im.dst_cp + (int64_t)row0 * im.dst_rp + col0;
    const float* nz_row = im.noise ? im.noise + (int64_t)ch * im.dst_cp + (int64_t)row0 * im.dst_rp + col0 : nullptr;
    uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(g) >> 2);
    const uint32_t rp_lo = (uint32_t)im.dst_rp;
    const int nrows = min(kRows, im.H - row0);
    // two rows per pass: accumulators -> the warp's two row buffers (each skewed so that shared and global addresses
    // agree mod 16 bytes: element x of a row sits at buffer[skew + x]), then 16-byte stores of both rows.
    // Output row q is the .x half of pair q for q < kR and the .y half of pair q - kR otherwise.
#pragma unroll
    for (int r = 0; r < kRows; r += 2) {
        const int skew0 = (int)(phase & 3u), skew1 = (int)((phase + rp_lo) & 3u);
        const uint32_t b0 = obuf, b1 = obuf + 4u * kOutPitch;
        if (r < nrows) {
#pragma unroll
            for (int c = 0; c < kCC; ++c)
                sts_f32(b0 + 4u * (uint32_t)(skew0 + kCC * lane + c), r < kR ? acc[r % kR][c].x : acc[r % kR][c].y);
        }
        if (r + 1 < nrows) {
#pragma unroll
            for (int c = 0; c < kCC; ++c)
                sts_f32(b1 + 4u * (uint32_t)(skew1 + kCC * lane + c), r + 1 < kR ? acc[(r + 1) % kR][c].x : acc[(r + 1) % kR][c].y);
        }
        __syncwarp();
        float* g1 = g + im.dst_rp;
        if (kEpi != kEpiGeneral) {
            if (r < nrows) store_row_plain<kEpi == kEpiAffine>(g, b0, skew0, wv, lane, aff_scale, aff_shift);
            if (r + 1 < nrows) store_row_plain<kEpi == kEpiAffine>(g1, b1, skew1, wv, lane, aff_scale, aff_shift);
        } else {
            const uint64_t stream = p.philox_offset + (uint64_t)im.philox_slot;
            const float* sm0 = reinterpret_cast<const float*>(__cvta_shared_to_generic(b0));
            if (r < nrows)
                store_row_epilogue(sm0, g, nz_row, wv, skew0, ep, p.philox_seed, stream,
                                   ((uint64_t)ch * im.H + (row0 + r)) * im.W + col0);
            if (r + 1 < nrows)
                store_row_epilogue(sm0 + kOutPitch, g1, nz_row ? nz_row + im.dst_rp : nullptr, wv, skew1, ep, p.philox_seed, stream,
                                   ((uint64_t)ch * im.H + (row0 + r + 1)) * im.W + col0);
        }
        __syncwarp();
        g += 2 * im.dst_rp;
        if (nz_row) nz_row += 2 * im.dst_rp;
        phase += 2u * rp_lo;
    }
}

// ---------------------------------------------------------------- kernel
// kEpi selects the epilogue variant (kEpiNone / kEpiAffine / kEpiGeneral); batches without one run the leanest.
template <int kEpi, bool kHalf>
__global__ void __launch_bounds__(kThreads, 1) blur_masked_kernel(const __grid_constant__ TiledParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* obuf_all = reinterpret_cast<float*>(smem + 2 * kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytes + kOutBufBytes);
    uint64_t* full = bars;          // [2] producers -> consumers: stage loaded
    uint64_t* empty = bars + 2;     // [2] consumers -> producers: stage may be refilled
    uint64_t* landed = bars + 4;    // [2] half I/O only: the stage's bulk copies have arrived, rows may be widened
    int* tile_slots = reinterpret_cast<int*>(bars + 6);
    const int warp = threadIdx.x >> 5;

    // Programmatic dependent launch: let the next launch on the stream start filling SMs as this grid's CTAs retire, and
    // -- unless the caller declared this batch independent of the previous launch -- wait for that launch to complete
    // (and flush) before touching global memory.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!p.overlap_prev) asm volatile("griddepcontrol.wait;" ::: "memory");
    DIB_TRACE_EVENT(63, 0);

    __shared__ int s_first_tab[DIB_MAX_BATCH + 1];
    const int* s_first = p.meta_dev != nullptr ? s_first_tab : nullptr;
    tile_table(p, 1, s_first_tab);
    if (threadIdx.x == 0) {
        // float: per producer thread one arrive.expect_tx + one cp.async arrive; half: one plain arrive after widening
        mbar_init(&full[0], (kHalf ? 1 : 2) * kProducerWarps * 32);
        mbar_init(&full[1], (kHalf ? 1 : 2) * kProducerWarps * 32);
        mbar_init(&landed[0], kProducerWarps * 32);
        mbar_init(&landed[1], kProducerWarps * 32);
        mbar_init(&empty[0], kComputeWarps);
        mbar_init(&empty[1], kComputeWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Register budget: the launch splits the register file evenly over all warps; the producer warpgroup hands most of
    // its share back so that the compute warps can hold kR * kCC accumulators + kR * kWinW window values per thread.
    if (warp < kProducerWarps) {        // producers are the lowest warp ids: the issue arbiter favours high ids
        if constexpr (kHalf)
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegsHalf));
        else
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
        // ------------------------------------------------ producer warpgroup
        const int pt = threadIdx.x;
        Stage cur, nxt;
        int nfetch = 0;
        unsigned pre_ticket = kNoTicket;
        cur.chunk = 0;
        do {        // (device-planned launches skip the tiles of images that belong to another kernel)
            cur.tile = fetch_tile(p, s_first, tile_slots, nfetch, pt);
            if (cur.tile < 0) break;
            decode_tile(p, s_first, cur.tile, cur);
        } while (cur.nchunks == 0);
        if (cur.tile >= 0) cur.rec = load_chunk_rec(p, cur.img, 0);
        for (int n = 0;; ++n) {
            const int b = n & 1;
            DIB_TRACE_EVENT(15, n);
            next_stage(p, s_first, cur, nxt, tile_slots, nfetch, pt, pre_ticket);     // its chunk record is in flight during the issue below
            pre_ticket = kNoTicket;
            // half path: the producers are busy for the whole stage (they widen the rows), so the ticket the NEXT call of
            // next_stage will need is drawn now and its round trip to the counter hides behind this stage's work
            if (kHalf && pt == 0 && nxt.tile >= 0 && nxt.chunk + 1 >= nxt.nchunks) pre_ticket = atomicAdd(&p.sched->next_tile, 1u);
            DIB_TRACE_EVENT(13, n);
            const bool wait_empty = n >= 2;
            const uint32_t empty_parity = (uint32_t)(((n >> 1) - 1) & 1);
            // the float path waits inside issue_stage, after the row arithmetic
            if ((kHalf || cur.tile < 0) && wait_empty) producer_wait(&empty[b], empty_parity, pt);
            DIB_TRACE_EVENT(14, n);
            const StageSmem sm = stage_smem(smem, b);
            if (cur.tile < 0) {
                if (pt == 0) sm.hdr->tile = -1;
                if constexpr (kHalf) {
                    mbar_arrive(&full[b]);
                } else {
                    mbar_arrive_expect_tx(&full[b], 0);
                    cp_async_mbar_arrive(&full[b]);
                }
                // this CTA has stopped fetching; the last CTA to get here rewinds the scheduler for the next launch
                if (pt == 0) {
                    __threadfence();
                    if (atomicAdd(&p.sched->done_ctas, 1u) == gridDim.x - 1) {
                        p.sched->next_tile = 0u;
                        p.sched->done_ctas = 0u;
                        __threadfence();
                    }
                }
                break;
            }
            if constexpr (kHalf)
                issue_stage_half(p, cur, sm, &landed[b], (uint32_t)((n >> 1) & 1), &full[b], pt);
            else
                issue_stage(p, cur, sm, &full[b], pt, &empty[b], empty_parity, wait_empty);
            cur = nxt;
        }
    } else {
        if constexpr (kHalf)
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kComputeRegsHalf));
        else
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kComputeRegs));
        // ------------------------------------------------ compute warps
        const int cw = warp - kProducerWarps;
        const int wrow = cw / kWarpCols, wcol = cw % kWarpCols;
        float2 acc[kR][kCC];
        const uint32_t obuf = smem_u32(obuf_all + cw * 2 * kOutPitch);
        for (int n = 0;; ++n) {
            const int b = n & 1;
            const StageSmem sm = stage_smem(smem, b);
            DIB_TRACE_EVENT(0, n);
            mbar_wait(&full[b], (n >> 1) & 1);
            DIB_TRACE_EVENT(1, n);
            StageHdr h;
            {   // explicit vector loads keep the header in registers
                const int4 a = reinterpret_cast<const int4*>(sm.hdr)[0], b4 = reinterpret_cast<const int4*>(sm.hdr)[1];
                const int4 c2 = reinterpret_cast<const int4*>(sm.hdr)[2];
                h.tile = a.x; h.img = a.y; h.ch = a.z; h.i0 = a.w;
                h.j0 = b4.x; h.first_chunk = b4.y; h.last_chunk = b4.z; h.dy_hi = b4.w;
                h.dx_hi = c2.x; h.nseg = c2.y; h.wsteps = c2.z;
            }
            if (h.tile < 0) break;
            if (h.first_chunk) {
#pragma unroll
                for (int r = 0; r < kR; ++r)
#pragma unroll
                    for (int c = 0; c < kCC; ++c) acc[r][c] = make_float2(0.0f, 0.0f);
            }
            const TiledImage& im = p.img[h.img];
            const int row0 = h.i0 + wrow * kRows, col0 = h.j0 + wcol * kWarpW;
            const bool active = row0 < im.H && col0 < im.W;                       // warp-uniform
            if (active) compute_chunk(acc, smem_u32(sm.hdr), h.nseg, h.dy_hi, h.dx_hi, wrow, wcol);
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[b]);     // this warp is done reading the stage
            DIB_TRACE_EVENT(4, n);
#ifdef DIB_NOSTORE      // experiment: upper bound of what hiding the store phase could gain (results are not written)
            if (h.last_chunk && active && acc[0][0].x == 12345.678f) store_rows<kEpi, kHalf>(p, im, h.ch, row0, col0, acc, obuf);
#else
            if (h.last_chunk && active) store_rows<kEpi, kHalf>(p, im, h.ch, row0, col0, acc, obuf);
#endif
            DIB_TRACE_EVENT(5, n);
        }
    }
}

// ---------------------------------------------------------------- host launcher
int tiled_tile_counts(int H, int W, int* tiles_y, int* tiles_x) {
    *tiles_y = (H + kTH - 1) / kTH;
    *tiles_x = (W + kTW - 1) / kTW;
    return *tiles_y * *tiles_x;
}

int launch_tiled(const dib_image* images, const int* order, int n_sel, const dib_psf_meta* meta_host, const uint8_t* prog,
                 SchedWords* sched, uint64_t seed, uint64_t offset, int io_dtype, bool overlap_prev, const dib_psf_meta* meta_dev,
                 cudaStream_t st) {
    static thread_local int sm_count = 0;
    static thread_local int attr_set_dev = -1;
    int dev = 0;
    DIB_CUDA(cudaGetDevice(&dev));
    if (attr_set_dev != dev) {
        DIB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
        DIB_CUDA(cudaFuncSetAttribute(blur_masked_kernel<kEpiNone, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        DIB_CUDA(cudaFuncSetAttribute(blur_masked_kernel<kEpiAffine, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        DIB_CUDA(cudaFuncSetAttribute(blur_masked_kernel<kEpiGeneral, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        DIB_CUDA(cudaFuncSetAttribute(blur_masked_kernel<kEpiNone, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        DIB_CUDA(cudaFuncSetAttribute(blur_masked_kernel<kEpiAffine, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set_dev = dev;
    }
    TiledParams p;
    int total = 0;
    bool any_epi = false, any_general = false;
    for (int k = 0; k < n_sel; ++k) {   // `order` lists the images heaviest PSF first: tiles are handed out in this order
        const dib_image& im = images[order[k]];
        // planned on the device (meta_dev): no host copy of the summaries -- the kernel reads chunk counts itself, tiles are
        // counted for the widest shear, and every image is listed (each kernel skips the images of the other)
        dib_psf_meta planned = {};
        planned.prog_chunks = -1;
        planned.prog_shear = (int16_t)kDevicePlanShear;
        const dib_psf_meta& m = meta_dev != nullptr ? planned : meta_host[im.psf_index];
        TiledImage& t = p.img[k];
        t.src = im.src;
        t.dst = im.dst;
        t.noise = static_cast<const float*>(im.noise);
        t.src_rp = im.src_row_pitch;
        t.src_cp = im.src_chan_pitch;
        t.dst_rp = im.dst_row_pitch;
        t.dst_cp = im.dst_chan_pitch;
        t.C = im.C;
        t.H = im.H;
        t.W = im.W;
        const int per_ch = tiled_tile_counts(im.H, im.W, &t.tiles_y, &t.tiles_x);
        t.first_tile = total;
        t.psf_index = im.psf_index;
        t.nchunks = m.prog_chunks;
        t.epilogue = im.epilogue;
        t.zero_pad = (im.pad_mode == DIB_PAD_ZERO128);
        // A PSF whose program is one chunk: taps.cu builds that chunk from the support's bounding box alone (first group at
        // xmin, rows ymin .. ymax, data right after the chunk table), so the record is reproduced here from the host summary.
        t.rec0_valid = 0;
        if (meta_dev == nullptr && m.prog_chunks == 1) {
            const int centre = 63;
            const int g_last = (m.xmax - m.xmin) / kGroupW;
            t.rec0.dy_lo = (int16_t)(m.ymin - centre);
            t.rec0.dy_hi = (int16_t)(m.ymax - centre);
            t.rec0.dx_lo = (int16_t)(m.xmin - centre);
            t.rec0.dx_hi = (int16_t)(m.xmin - centre + g_last * kGroupW + kGroupW - 1);
            t.rec0.nseg = (int16_t)m.prog_segs;
            t.rec0.wsteps = (int16_t)m.prog_steps;
            t.rec0.data_off = (int32_t)kProgHeaderBytes;
            t.rec0_valid = 1;
        }
        t.aligned_out = ((reinterpret_cast<uintptr_t>(im.dst) & 15u) == 0 && (im.dst_row_pitch & 3) == 0 && (im.dst_chan_pitch & 3) == 0) ? 1 : 0;
        if ((t.epilogue & DIB_EPI_NOISE) && !(t.epilogue & DIB_EPI_PHILOX) && t.noise == nullptr) t.epilogue &= ~DIB_EPI_NOISE;
        any_epi |= (t.epilogue != 0);
        any_general |= (t.epilogue & ~DIB_EPI_NORMALIZE) != 0;
        t.philox_slot = order[k];
        t.noise_sd = im.noise_sd;
        t.gamma = im.gamma;
        for (int c = 0; c < 4; ++c) {
            t.mean[c] = im.mean[c];
            t.std[c] = im.std[c];
        }
        total += per_ch * im.C;
    }
    p.prog = prog;
    p.n_images = n_sel;
    p.total_tiles = total;
    p.philox_seed = seed;
    p.philox_offset = offset;
    p.sched = sched;
    p.meta_dev = meta_dev;
    // a grid smaller than the machine could be co-resident with many successors (more than the scheduler slots cover):
    // it always orders itself after its predecessor
    p.overlap_prev = (overlap_prev && total >= sm_count) ? 1 : 0;
    const int grid = total < sm_count ? total : sm_count;     // persistent: one CTA per SM
    // launched with the programmatic-stream-serialization attribute: the kernel itself decides (griddepcontrol.wait)
    // whether it orders itself after the previous launch
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (io_dtype == DIB_F16) {
        if (any_general) {
            set_error("dib_blur_batch: half images with a noise / clamp / gamma epilogue do not take the tiled kernel");
            return DIB_ERR_UNSUPPORTED;
        }
        if (any_epi)
            DIB_CUDA(cudaLaunchKernelEx(&cfg, blur_masked_kernel<kEpiAffine, true>, p));
        else
            DIB_CUDA(cudaLaunchKernelEx(&cfg, blur_masked_kernel<kEpiNone, true>, p));
    } else if (any_general) {
        DIB_CUDA(cudaLaunchKernelEx(&cfg, blur_masked_kernel<kEpiGeneral, false>, p));
    } else if (any_epi) {
        DIB_CUDA(cudaLaunchKernelEx(&cfg, blur_masked_kernel<kEpiAffine, false>, p));
    } else {
        DIB_CUDA(cudaLaunchKernelEx(&cfg, blur_masked_kernel<kEpiNone, false>, p));
    }
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

}  // namespace mk
}  // namespace dib

#ifdef DIB_TRACE
extern "C" __attribute__((visibility("default"))) int dib_debug_trace_masked(unsigned long long* out, int max_events) {
    unsigned int n[16];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(n, dib::mk::g_trace_n, sizeof(n));
    int total = 0;
    for (int w = 0; w < 16; ++w) {
        int c = (int)n[w];
        if (c > dib::mk::kTracePerWarp) c = dib::mk::kTracePerWarp;
        if (total + c > max_events) c = max_events - total;
        if (c > 0) cudaMemcpyFromSymbol(out + total, dib::mk::g_trace, sizeof(unsigned long long) * c, sizeof(unsigned long long) * w * dib::mk::kTracePerWarp);
        total += c > 0 ? c : 0;
    }
    return total;
}
#endif

// Tiled sparse-PSF blur for sm_100a (DIB_ALGO_TILED): the fast path of dib_blur_batch.
//
// Replaces the per-tap `output += torch.roll(pad(image), shift) * w` loop of manual_blur
// (models/blur_functions.py:59-69) -- O(taps) launches and ~7 passes over the padded tensor per tap -- with one
// persistent, warp-specialised launch per batch:
//   * work unit  = one kTH x kTW (32 x 448) output tile of one channel of one image; the CTAs (one per SM) take tiles
//                  from a global ticket counter, images with the heaviest PSFs first;
//   * producers  = one warpgroup (4 warps, 56 registers after setmaxnreg.dec).  Every image carries a ONE-DIMENSIONAL TMA
//                  tensor map over its elements (kernel parameter), so a staged row is two `cp.async.bulk.tensor.1d` boxes
//                  of 256 floats (SASS UTMALDG) whatever the row pitch: the reference's unpitched CHW layout (5332 B per
//                  row for W = 1333) needs no repacking and no per-row head / tail handling.  A box must start 16-byte
//                  aligned in global memory (measured: tools/exp/tma1d_probe.cu), so a row's box starts at the aligned
//                  element at or below its first wanted pixel and the row sits `skew` = 0..3 floats into its
//                  shared-memory row; for consecutive image rows the skew is an arithmetic progression mod 4 (step =
//                  row pitch mod 4), which the consumers fold into four precomputed addresses.  What a linear box cannot
//                  give -- rows above / below the image (reflect-101 or zeros) and the columns left and right of it --
//                  is patched by the producer threads once the boxes have landed (4-byte cp.async from the mirrored
//                  pixel, placed where the arithmetic progression expects it), then the stage is handed to the
//                  consumers through the "full" mbarrier.  Two stages are in flight;
//   * consumers  = kComputeWarps warps (4 x 2 over the tile, two per SM sub-partition, 224 registers after
//                  setmaxnreg.inc).  Each thread owns an 8-row x 7-column output block (lanes sit 7 floats apart in a
//                  row: odd stride -> conflict-free scalar LDS); rows r and r + 4 share 64-bit accumulators and every
//                  multiply-add is a packed FFMA2.  Taps are consumed as the program built by taps.cu: the PSF support,
//                  sheared so that a slanted streak becomes near-vertical, is cut into groups of 2 (or 4) columns; a group
//                  is swept row by row with a rotating 8-row register window of the input (one new row per step, loaded
//                  into the slot the previous step freed, consumed last) and EVERY step executes all taps of the group
//                  -- zero weights where the PSF has none -- so the sweep has no data-dependent branch at all.  The
//                  output block of a thread is sheared like the program (row r is shifted by shear * r columns), which
//                  makes the window's column offset a constant per step;
//   * epilogue   = noise / clamp / gamma / (x - mean) / std (blur_functions.py:72-74, net_transforms.py:135-139) fused
//                  on the way out, per warp and without block-level barriers: accumulators -> the warp's private
//                  two-row buffer -> 16-byte vector stores.
// No tensor cores: the contraction is sparse and data dependent.  Results differ from the exact-order kernel only
// by FMA contraction and tap order (measured <= 4e-7 on [0,1] images; bound 1e-5).
#include <cuda.h>

#include "dib_common.cuh"

namespace dib {

#ifndef DIB_PRODUCER_REGS
#define DIB_PRODUCER_REGS 88
#endif
constexpr int kDevicePlanShear = kShearMax;  // device-planned launches count tiles for the widest shear
constexpr int kProducerWarps = 4;           // one warpgroup: a thread issues at most one TMA box per stage
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kThreads = (kComputeWarps + kProducerWarps) * 32;
constexpr int kProducerRegs = DIB_PRODUCER_REGS;   // setmaxnreg budgets; together they must fit the 64K-register file
// setmaxnreg moves registers inside the CTA's launch-time allocation (threads x the per-thread count the launch bound
// allows, a multiple of 8); asking for more than the producers hand back blocks forever.
constexpr int kLaunchRegs = (65536 / (kThreads * kCtasPerSm)) / 8 * 8 > 255 ? 248 : (65536 / (kThreads * kCtasPerSm)) / 8 * 8;
constexpr int kComputeRegsRaw = (kThreads * kLaunchRegs - kProducerThreads * kProducerRegs) / (kComputeWarps * 32) / 8 * 8;
constexpr int kComputeRegs = kComputeRegsRaw > 232 ? 232 : kComputeRegsRaw;
static_assert(kComputeWarps % 4 == 0, "warpgroup-aligned compute warps");
constexpr int kBoxElems = 256;              // elements per TMA box (the hardware maximum per dimension)
constexpr int kBoxesPerRow = kPitch / kBoxElems;
constexpr int kRowBytes = kPitch * 4;
constexpr int kTileBytes = kRowsMax * kRowBytes;
constexpr int kStageBytes = kStageHdrBytes + kChunkAuxBytes + kTileBytes;
constexpr int kSmemBytes = kStages * kStageBytes + kOutBufBytes + 128;    // + 3 * kStages mbarriers + 4 ticket / flag slots
static_assert(kStages == 1 || kStages == 2, "one or two shared-memory stages");
static_assert((kStageHdrBytes + kChunkAuxBytes) % 128 == 0 && kStageBytes % 128 == 0, "TMA destinations are 128-byte aligned");
static_assert(kSmemBytes <= 232448 && kCtasPerSm * (kSmemBytes + 1024) <= 233472, "exceeds the shared memory of an sm_100 SM");

struct alignas(64) TiledImage {
    CUtensorMap tmap;     // 1-D map over the image's elements, from the 16-byte-aligned address at or below src
    const void* src;      // float, pitches in elements
    void* dst;
    const float* noise;
    int64_t src_rp, src_cp, dst_rp, dst_cp;
    int src_off;          // elements between the tensor map's base and src
    int C, H, W;
    int tiles_x, tiles_y;
    int first_tile;       // tiles of the images before this one
    int psf_index, nchunks;
    int shear;            // the program's shear: output row r of a warp's block starts shear * r columns further right
    int epilogue;
    int zero_pad;         // DIB_PAD_ZERO128: pixels outside the image read as 0 instead of being mirrored
    int aligned_out;      // every destination row starts 16-byte aligned (base, row pitch and channel pitch): lean row store
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
    // per-warp slices, the write position kept in shared memory (no global atomics: their round trip would dominate)
    __shared__ unsigned int pos[16];
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        unsigned int i = pos[w];
        if (i >= (unsigned)kTracePerWarp) i = 0;         // shared memory starts uninitialised; slices are reset by the host
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep in hardware, do not spin
}
// Producer-group wait: ONE thread polls the mbarrier (with a back-off: nobody needs the answer within nanoseconds), the
// other 127 block in a named barrier, which costs no issue slots -- four warps spinning on try_wait took ~14 % of the SM's
// issue slots away from the compute warps (profiles/round2_notes.md).
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
// non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// TMA: one box of a 1-D tensor map, global -> shared, completion counted in bytes on the mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_box_1d(uint32_t dst, const CUtensorMap* map, int coord, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(coord), "r"(smem_u32(bar))
                 : "memory");
}
// TMA bulk copy global -> shared of a 16-byte-aligned span (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
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
__device__ __forceinline__ void lds_entry(uint32_t addr, int& a, int& b) {
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_v2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_zero16(uint32_t addr) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "f"(0.0f) : "memory");
}

__device__ __forceinline__ int reflect101(int v, int n) {
    v = v < 0 ? -v : v;
    v = v >= n ? 2 * (n - 1) - v : v;
    return min(max(v, 0), n - 1);     // clamp only guards rows/cols that feed zero weights or masked outputs
}

// ---------------------------------------------------------------- stage bookkeeping
// What the producer tells the consumers about a stage (first 64 bytes of the stage's shared memory).
struct __align__(16) StageHdr {
    int tile;           // global tile index; -1 = no more work
    int img, ch, i0, j0;
    int first_chunk, last_chunk;
    int dy_hi, dx_hi, nseg, group_w, shear;
    int skew0, dskew;   // staged row r holds image column cl at float offset (skew0 + r * dskew) & 3
};
static_assert(sizeof(StageHdr) <= kStageHdrBytes, "stage header too large");

struct Stage {
    int tile;       // global tile index, -1: none
    int chunk;
    int nchunks;    // chunks of the image's program; 0: the image belongs to another kernel (device-planned launches)
    int img, ch, i0, j0;
    ChunkRec rec;
};

__device__ __forceinline__ uint32_t stage_base(uint32_t smem_base, int b) { return smem_base + (uint32_t)b * kStageBytes; }

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
        st.nchunks = psf_program_kind(mm) == 2 ? mm.prog_chunks : 0;
    }
}

__device__ __forceinline__ ChunkRec load_chunk_rec(const TiledParams& p, int img, int chunk) {
    const uint8_t* prog = p.prog + (size_t)p.img[img].psf_index * kProgBytes;
    const int4 v = __ldg(reinterpret_cast<const int4*>(prog) + chunk);
    ChunkRec r;
    r.dy_lo = (int16_t)(v.x & 0xffff);
    r.dy_hi = (int16_t)(v.x >> 16);
    r.dx_lo = (int16_t)(v.y & 0xffff);
    r.dx_hi = (int16_t)(v.y >> 16);
    r.wsteps = (int16_t)(v.z & 0xffff);
    r.nseg = (uint8_t)((v.z >> 16) & 0xff);
    r.shear = (int8_t)((v.z >> 24) & 0xff);
    r.group_w = (uint8_t)(v.w & 0xff);
    r.pad = 0;
    r.data_off16 = (uint16_t)((unsigned)v.w >> 16);
    return r;
}

// Hand the producer group its next tile.  Thread 0 of the group takes a ticket from the global counter and shares
// it through shared memory; the two slots alternate so one named barrier per fetch is enough.
__device__ __forceinline__ int fetch_tile(const TiledParams& p, const int* first, int* slots, int& nfetch, int pt) {
    int* slot = slots + (nfetch & 1);
    if (pt == 0) {
        // the first tile of a CTA is its block index (no round trip to the counter before the first load can go out);
        // tickets from the counter follow after the gridDim.x tiles handed out that way
        const unsigned t = nfetch == 0 ? blockIdx.x : gridDim.x + atomicAdd(&p.sched->next_tile, 1u);
        *slot = t < (unsigned)(first != nullptr ? first[DIB_MAX_BATCH] : p.total_tiles) ? (int)t : -1;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");
    ++nfetch;
    return *slot;
}

// successor of a stage: next chunk of the same tile, else chunk 0 of the next tile the scheduler hands out
__device__ __forceinline__ void next_stage(const TiledParams& p, const int* first, const Stage& cur, Stage& nx, int* slots, int& nfetch,
                                           int pt) {
    if (cur.tile < 0) {
        nx.tile = -1;
        return;
    }
    if (cur.chunk + 1 < cur.nchunks) {
        nx = cur;
        nx.chunk = cur.chunk + 1;
    } else {
        do {        // device-planned launches: tiles of images that belong to another kernel are skipped
            nx.tile = fetch_tile(p, first, slots, nfetch, pt);
            if (nx.tile < 0) return;
            nx.chunk = 0;
            decode_tile(p, first, nx.tile, nx);
        } while (nx.nchunks == 0);
    }
    nx.rec = load_chunk_rec(p, nx.img, nx.chunk);
}

// geometry of a stage: image row of staged row 0, staged rows, image column of staged column 0, last staged image column
struct StageGeom {
    int rt, nrows, cl, cr;
    int skew0, dskew;       // image column cl sits (skew0 + r * dskew) & 3 floats into staged row r
};
__device__ __forceinline__ StageGeom stage_geom(const TiledImage& im, const Stage& st) {
    StageGeom g;
    const int shear = st.rec.shear;
    const int pad = (kRows - 1) * (shear < 0 ? -shear : shear);   // a warp's sheared block reaches this far left of its nominal origin
    g.rt = st.i0 - st.rec.dy_hi;
    g.nrows = kTH + st.rec.dy_hi - st.rec.dy_lo;
    g.cl = st.j0 - pad - st.rec.dx_hi;
    g.cr = min(st.j0 + kTW, im.W) - 1 - st.rec.dx_lo;
    // element index (from the tensor map's aligned base) of (row rt, column cl), as if the image continued above row 0
    const int64_t e0 = (int64_t)im.src_off + (int64_t)st.ch * im.src_cp + (int64_t)g.rt * im.src_rp + g.cl;
    g.skew0 = (int)(e0 & 3);
    g.dskew = (int)(im.src_rp & 3);
    return g;
}

// Producer group, first half of a stage: start every bulk load.  Thread t issues box (t mod 2) of staged row t / 2: 256
// consecutive elements of the image row from the 16-byte-aligned element at or below image column cl (+ 256) -- elements
// before the row's first or past its last pixel are neighbouring rows' pixels (or zeros outside the tensor) and are
// patched by finish_stage, like the rows above and below the image, which no box is issued for.
__device__ __forceinline__ void issue_stage(const TiledParams& p, const Stage& st, uint32_t sbase, uint64_t* landed, int pt) {
    const TiledImage& im = p.img[st.img];
    const StageGeom g = stage_geom(im, st);
    fence_proxy_async();   // order the consumers' generic-proxy reads of this buffer before the async-proxy writes
    if (pt == 0) {
        StageHdr h;
        h.tile = st.tile; h.img = st.img; h.ch = st.ch; h.i0 = st.i0; h.j0 = st.j0;
        h.first_chunk = (st.chunk == 0);
        h.last_chunk = (st.chunk + 1 == st.nchunks);
        h.dy_hi = st.rec.dy_hi; h.dx_hi = st.rec.dx_hi; h.nseg = st.rec.nseg; h.group_w = st.rec.group_w; h.shear = st.rec.shear;
        h.skew0 = g.skew0; h.dskew = g.dskew;
        StageHdr* hp = reinterpret_cast<StageHdr*>(__cvta_shared_to_generic(sbase));
        *hp = h;
    }
    const uint32_t aux = sbase + kStageHdrBytes, tile = aux + kChunkAuxBytes;
    const uint32_t aux_bytes = (uint32_t)(kChunkSegBytes + 4 * st.rec.group_w * (st.rec.wsteps + 1) + 15) & ~15u;
    for (int t = pt; t < g.nrows * kBoxesPerRow; t += kProducerThreads) {
        const int row = t / kBoxesPerRow, box = t - row * kBoxesPerRow;
        const int irow = g.rt + row;
        if (irow >= 0 && irow < im.H) {
            const int64_t e = (int64_t)im.src_off + (int64_t)st.ch * im.src_cp + (int64_t)irow * im.src_rp + g.cl;
            tma_box_1d(tile + (uint32_t)row * kRowBytes + (uint32_t)box * (kBoxElems * 4), &im.tmap, (int)(e & ~int64_t(3)) + box * kBoxElems, landed);
        }
    }
    if (pt == kProducerThreads - 1)
        tma_bulk_g2s(aux, p.prog + (size_t)im.psf_index * kProgBytes + (size_t)st.rec.data_off16 * 16, aux_bytes, landed);
    if (pt == 0) {
        const int rows_in = max(0, min(g.rt + g.nrows, im.H) - max(g.rt, 0));
        mbar_arrive_expect_tx(landed, (uint32_t)rows_in * kRowBytes + aux_bytes);
    }
}

// Producer group, second half of a stage: once the boxes have landed, write what they could not deliver and hand the
// stage to the consumers.  Two passes, a warp per staged row, a lane per column:
//   1. rows inside the image: the columns left of the image's first and right of its last pixel are reflect-101 pixels of
//      the SAME row, which is already in shared memory -- one LDS + STS each (a mirrored column that the stage does not
//      hold -- only a last column tile a few pixels wide -- comes from global memory with a 4-byte cp.async);
//   2. rows above / below the image: reflect-101 copies of staged rows inside it, border columns included, shifted by the
//      difference of the two rows' skews.
// Zero-padding mode stores zeros instead.  Every producer thread then arrives on "full" (once its cp.asyncs, if any,
// have landed).
__device__ __forceinline__ void finish_stage(const TiledParams& p, const Stage& st, uint32_t sbase, uint64_t* landed, uint32_t parity,
                                             uint64_t* full, int pt) {
    const TiledImage& im = p.img[st.img];
    const StageGeom g = stage_geom(im, st);
    DIB_TRACE_EVENT(10, st.chunk);
    producer_wait(landed, parity, pt);
    DIB_TRACE_EVENT(11, st.chunk);
    const int top = min(g.nrows, max(0, -g.rt));                              // staged rows above the image
    const int bot = min(g.nrows - top, max(0, g.rt + g.nrows - im.H));         // staged rows below it
    const bool cols_out = g.cl < 0 || g.cr >= im.W;
    if (top > 0 || bot > 0 || cols_out) {
        const uint32_t tile = sbase + kStageHdrBytes + kChunkAuxBytes;
        const float* plane = static_cast<const float*>(im.src) + (int64_t)st.ch * im.src_cp;
        const int lane = pt & 31, pw = pt >> 5;
        const int ncols = g.cr - g.cl + 1;
        if (cols_out) {
            const int xa = max(g.cl, 0), xb1 = min(g.cr, im.W - 1) + 1;          // in-image columns [xa, xb1)
            // (a chunk whose taps all point past the right edge of a last column tile stages no image column: cl >= W; the
            //  mirrored columns then start at cl, not at the first column beyond the image)
            const int rb = max(xb1, g.cl);
            const int nleft = min(xa, g.cr + 1) - g.cl, nborder = nleft + g.cr + 1 - rb;
            for (int r2 = top + pw; r2 < g.nrows - bot; r2 += kProducerWarps) {
                const int skew = (g.skew0 + r2 * g.dskew) & 3;
                const uint32_t row = tile + (uint32_t)r2 * kRowBytes + 4u * (uint32_t)(skew - g.cl);     // + 4 * col
                for (int k = lane; k < nborder; k += 32) {
                    const int col = k < nleft ? g.cl + k : rb + (k - nleft);
                    if (im.zero_pad) {
                        sts_f32(row + 4u * (uint32_t)col, 0.0f);
                    } else {
                        const int sc = reflect101(col, im.W);
                        if (sc >= g.cl && sc - g.cl + skew < kPitch)
                            sts_f32(row + 4u * (uint32_t)col, lds_f32(row + 4u * (uint32_t)sc));
                        else
                            cp_async_4(row + 4u * (uint32_t)col, plane + (int64_t)(g.rt + r2) * im.src_rp + sc);
                    }
                }
            }
        }
        if (top > 0 || bot > 0) {
            // pass 2 reads what pass 1 wrote (and, through the fallback, what its cp.asyncs deliver)
            asm volatile("cp.async.wait_all;" ::: "memory");
            asm volatile("bar.sync 2, %0;" ::"n"(kProducerThreads) : "memory");
            for (int q = pw; q < top + bot; q += kProducerWarps) {
                const int r2 = q < top ? q : g.nrows - bot + (q - top);
                const int irow = g.rt + r2;
                const int skew = (g.skew0 + r2 * g.dskew) & 3;
                const uint32_t drow = tile + (uint32_t)r2 * kRowBytes + 4u * (uint32_t)skew;              // + 4 * (col - cl)
                if (im.zero_pad) {
                    for (int k = lane; k < ncols; k += 32) sts_f32(drow + 4u * (uint32_t)k, 0.0f);
                    continue;
                }
                const int srow_img = reflect101(irow, im.H);
                const int rs = srow_img - g.rt;                                  // staged row that holds the mirrored image row
                if (rs >= 0 && rs < g.nrows) {
                    const uint32_t srow = tile + (uint32_t)rs * kRowBytes + 4u * (uint32_t)((g.skew0 + rs * g.dskew) & 3);
                    for (int k = lane; k < ncols; k += 32) sts_f32(drow + 4u * (uint32_t)k, lds_f32(srow + 4u * (uint32_t)k));
                } else {
                    for (int k = lane; k < ncols; k += 32)
                        cp_async_4(drow + 4u * (uint32_t)k, plane + (int64_t)srow_img * im.src_rp + reflect101(g.cl + k, im.W));
                }
            }
        }
    }
    __threadfence_block();      // the patches are plain stores: order them before the arrive that publishes the stage
    cp_async_mbar_arrive(full);
    DIB_TRACE_EVENT(12, st.chunk);
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
template <int W, int U, int Q>
__device__ __forceinline__ float2 window_pair(const float2 (&win)[kR][W], int k) {
    constexpr int slot = ((Q - U) % kRows + kRows) % kRows;
    if constexpr (slot < kR)
        return win[slot][k];
    else
        return make_float2(win[slot - kR][k].y, win[slot - kR][k].x);
}

// All taps of one step on row pair P: weight w[e] multiplies the window shifted by e columns.
template <int G, int U, int P>
__device__ __forceinline__ void fma_pair(float2 (&acc)[kR][kCC], const float2 (&win)[kR][kCC + G - 1], const float (&w)[G]) {
#pragma unroll
    for (int e = 0; e < G; ++e) {
#pragma unroll
        for (int c = 0; c < kCC; ++c) acc[P][c] = ffma2(w[e], window_pair<kCC + G - 1, U, P>(win, c - e + G - 1), acc[P][c]);
    }
}
template <int G, int U, int P>
struct PairLoop {      // pairs kR-1 .. 1, then pair 0: it contains the row loaded at the start of the step and goes last
    __device__ __forceinline__ static void run(float2 (&acc)[kR][kCC], const float2 (&win)[kR][kCC + G - 1], const float (&w)[G]) {
        fma_pair<G, U, P>(acc, win, w);
        if constexpr (P > 0) PairLoop<G, U, (P == 1 ? 0 : P - 1)>::run(acc, win, w);
    }
};
template <int G, int U>
struct PairLoop<G, U, 0> {
    __device__ __forceinline__ static void run(float2 (&acc)[kR][kCC], const float2 (&win)[kR][kCC + G - 1], const float (&w)[G]) {
        fma_pair<G, U, 0>(acc, win, w);
    }
};

// load one input row into window slot SLOT
template <int W, int SLOT>
__device__ __forceinline__ void load_row(float2 (&win)[kR][W], uint32_t addr) {
#pragma unroll
    for (int k = 0; k < W; ++k) {
        if constexpr (SLOT < kR)
            win[SLOT][k].x = lds_f32(addr + 4 * k);
        else
            win[SLOT - kR][k].y = lds_f32(addr + 4 * k);
    }
}

// the G weights of one step
template <int G>
struct WeightVec {
    float v[G];
};
// Volatile loads: the vector fetched during step s is the one step s + 1 consumes, and ptxas must not sink the fetch below
// the loop's exit branch into the step that needs it (it did: ~30 cycles of exposed shared-memory latency per step).
template <int G>
__device__ __forceinline__ WeightVec<G> lds_weights(uint32_t addr) {
    WeightVec<G> w;
    if constexpr (G == 2) {
        asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(w.v[0]), "=f"(w.v[1]) : "r"(addr));
    } else {
        asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.v[0]), "=f"(w.v[1]), "=f"(w.v[2]), "=f"(w.v[3]) : "r"(addr));
    }
    return w;
}

// Step s of a segment sweep, s mod kRows == U: fetch the new top row into the slot the previous step freed and the NEXT
// step's weight vector (kRows is even, so the two weight registers simply alternate), then accumulate all G taps of this
// step's vector.  No data-dependent control flow: absent taps carry zero weights.  The row's address comes from one of
// four address registers (s mod 4): the skew of consecutive staged rows has period 4.
template <int G, int U>
__device__ __forceinline__ bool sweep_step(float2 (&acc)[kR][kCC], float2 (&win)[kR][kCC + G - 1], uint32_t (&addr)[4], uint32_t stride4,
                                           int& s, int nsteps, uint32_t& wp, WeightVec<G> (&wv)[2]) {
    load_row<kCC + G - 1, (kRows - U % kRows) % kRows>(win, addr[U & 3]);
    addr[U & 3] -= stride4;
    wp += 4u * G;
    wv[(U + 1) & 1] = lds_weights<G>(wp);                                 // weights of step s + 1 (zero vector past the end)
    PairLoop<G, U % kRows, kR - 1>::run(acc, win, wv[U & 1].v);
    ++s;
    return s < nsteps;
}

// The unrolled body: kPeriod consecutive steps = whole rotations of the window registers (period kRows) AND of the row
// skew (period 4), so both the window slot and the address register of a step are compile-time choices.
constexpr int kPeriod = (kRows % 4 == 0) ? kRows : (kRows % 2 == 0 ? 2 * kRows : 4 * kRows);
template <int G, int U>
struct SweepRound {
    __device__ __forceinline__ static bool run(float2 (&acc)[kR][kCC], float2 (&win)[kR][kCC + G - 1], uint32_t (&addr)[4], uint32_t stride4,
                                               int& s, int nsteps, uint32_t& wp, WeightVec<G> (&wv)[2]) {
        if (!sweep_step<G, U>(acc, win, addr, stride4, s, nsteps, wp, wv)) return false;
        if constexpr (U + 1 < kPeriod)
            return SweepRound<G, U + 1>::run(acc, win, addr, stride4, s, nsteps, wp, wv);
        else
            return true;
    }
};
static_assert(kPeriod % 4 == 0 && kPeriod % kRows == 0 && kPeriod % 2 == 0, "period covers skew, window and weight-register rotation");

// rows 1 .. kRows-1 of step 0's window (row 0 is step 0's own new row): logical row q -> slot q, staged row sr0 + q
template <int W, int Q>
__device__ __forceinline__ void fill_window(float2 (&win)[kR][W], uint32_t a0, uint32_t stride, int skew_sr0, int dskew) {
    load_row<W, Q>(win, a0 + (uint32_t)Q * stride + 4u * (uint32_t)((skew_sr0 + Q * dskew) & 3));
    if constexpr (Q + 1 < kRows) fill_window<W, Q + 1>(win, a0, stride, skew_sr0, dskew);
}

template <int G>
__device__ __forceinline__ void compute_chunk(float2 (&acc)[kR][kCC], uint32_t stage_addr, int nseg, int dy_hi, int dx_hi, int shear,
                                              int skew0, int dskew, int wrow, int wcol) {
    constexpr int kWinW = kCC + G - 1;
    const int lane = threadIdx.x & 31;
    const uint32_t aux = stage_addr + kStageHdrBytes;
    const uint32_t tile = aux + kChunkAuxBytes;
    // a staged row further down is one output row further down in the thread's block, whose columns start `shear` further
    // right: the window's column origin moves with the row, by a constant number of bytes per step
    const uint32_t stride = 4u * (uint32_t)(kPitch + shear);
    const int o1 = shear < 0 ? -(kRows - 1) * shear : 0;
    // The two compute warps that share an SM sub-partition (warp ids w and w + 4) walk the segments in opposite
    // orders: otherwise they run the same instruction sequence in lockstep and their load phases coincide.
    const bool reverse = (((threadIdx.x >> 5) - kProducerWarps) & 4) != 0 || (kComputeWarps == 4 && 2 * blockIdx.x >= gridDim.x);
#pragma unroll 1
    for (int sgi = 0; sgi < nseg; ++sgi) {
        const int sg = reverse ? nseg - 1 - sgi : sgi;
        int raw0, raw1;         // SegRec {dx0, dy0 | nsteps, woff} as two words
        lds_entry(aux + 8u * (uint32_t)sg, raw0, raw1);
        const int seg_dx0 = (int)(short)(raw0 & 0xffff), seg_dy0 = raw0 >> 16;
        const int nsteps = (int)(short)(raw1 & 0xffff), seg_woff = raw1 >> 16;
        const int colbase = wcol * kWarpW + kCC * lane - seg_dx0 - (G - 1) + o1 + dx_hi;
        const int sr0 = wrow * kRows - seg_dy0 + dy_hi;    // staged row of output row 0 at step 0
        DIB_TRACE_EVENT(2, nsteps);
        const uint32_t a0 = tile + 4u * (uint32_t)(sr0 * kPitch + colbase);
        const int skew_sr0 = skew0 + sr0 * dskew;           // (.. & 3) = skew of staged row sr0
        uint32_t wp = aux + kChunkSegBytes + 4u * G * (uint32_t)seg_woff;
        WeightVec<G> wv[2];
        wv[0] = lds_weights<G>(wp);
        float2 win[kR][kWinW];
        fill_window<kWinW, 1>(win, a0, stride, skew_sr0, dskew);
        // step s loads staged row sr0 - s: address a0 - s * stride + 4 * ((skew_sr0 - s * dskew) & 3); the last term has
        // period 4 in s, so step s uses addr[s & 3] and then moves it four steps on
        uint32_t addr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) addr[j] = a0 - (uint32_t)j * stride + 4u * (uint32_t)((skew_sr0 - j * dskew) & 3);
        int s = 0;
        DIB_TRACE_EVENT(6, sg);
#pragma unroll 1
        while (SweepRound<G, 0>::run(acc, win, addr, 4u * stride, s, nsteps, wp, wv)) {
        }
        DIB_TRACE_EVENT(3, sg);
    }
}

// ---------------------------------------------------------------- epilogue + store (per warp, no block barrier)
// One staged row -> global with the fused epilogue (noise / clamp / gamma / normalize).  Kept out of line: the
// common no-epilogue path below stays small and the register-tile code is not replicated around powf / Philox.
// `srow` is the 16-byte-aligned start of the staged row; element x of the row sits at srow[skew + x]; elements
// [x_lo, x_hi) are stored.
__device__ __noinline__ void store_row_epilogue(const float* srow, float* g, const float* nz_row, int x_lo, int x_hi, int skew, Epilogue ep,
                                                uint64_t seed, uint64_t stream, uint64_t pbase) {
    const int lane = threadIdx.x & 31;
    for (int k = lane; 4 * k - skew < x_hi; k += 32) {
        const int x0 = 4 * k - skew;
        const float4 v = *reinterpret_cast<const float4*>(srow + 4 * k);
        float o[4] = {v.x, v.y, v.z, v.w};
        const bool whole = (x0 >= x_lo) && (x0 + 3 < x_hi);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + j;
            if (x >= x_lo && x < x_hi) {
                float nz = 0.f;
                if (ep.flags & DIB_EPI_NOISE) nz = nz_row ? nz_row[x] : philox_normal(seed, stream, pbase + x);
                o[j] = apply_epilogue_f32(o[j], ep, nz);
                if (!whole) g[x] = o[j];
            }
        }
        if (whole) *reinterpret_cast<float4*>(g + x0) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// One staged row -> global without epilogue.  Element x of the row sits at srow + 4 * (skew + x); elements [x_lo, x_hi)
// are stored.  Aligned quads k0 .. k1-1 lie wholly inside the range and go out as 16-byte stores; the <= 3 elements before
// the first and after the last whole quad are stored one per lane.
template <bool kAffine>
__device__ __forceinline__ void store_row_plain(float* g, uint32_t srow, int skew, int x_lo, int x_hi, int lane, float scale, float shift) {
    const int k0 = (x_lo + skew + 3) >> 2, k1 = max((x_hi + skew) >> 2, k0);
    const int head1 = min(4 * k0 - skew, x_hi), tail0 = max(4 * k1 - skew, head1);
    const int nhead = head1 - x_lo;
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
    const int x = lane < nhead ? x_lo + lane : tail0 + (lane - nhead);
    if (x < x_hi && (lane < nhead || x >= tail0)) {
        float v = lds_f32(srow + 4u * (uint32_t)(skew + x));
        if (kAffine) v = fmaf(v, scale, shift);
        g[x] = v;
    }
}

// Destination rows that all start 16-byte aligned (a pitched output, e.g. the padded batch or the wrapper's own
// allocations) and an unsheared block need no per-row skew, no head elements and -- except in the last column tile -- no
// tail: a third of the general store's instructions.  Row q of the pass sits unskewed at obuf + q * kOutPitch.
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

// Epilogue variants of the kernel: none; normalize only, applied as one FMA per pixel, x * (1/std) - mean/std (the
// tiled kernel is not the bit-exact path, and an IEEE division per pixel would double its store cost); everything else.
constexpr int kEpiNone = 0, kEpiAffine = 1, kEpiGeneral = 2;

// `col0` is the nominal first column of the warp's block; with a sheared program output row r of the block starts at
// col0 + shear * r - (shear > 0 ? shear * (kRows - 1) : 0) and may begin left of the image or end right of it.
template <int kEpi>
__device__ __forceinline__ void store_rows(const TiledParams& p, const TiledImage& im, int ch, int row0, int col0, int shear,
                                           float2 (&acc)[kR][kCC], uint32_t obuf) {
    const int lane = threadIdx.x & 31;
    Epilogue ep;
    ep.flags = im.epilogue;
    ep.noise_sd = im.noise_sd;
    ep.gamma = im.gamma;
    ep.mean = im.mean[ch & 3];
    ep.std = im.std[ch & 3];
    const bool norm = (im.epilogue & DIB_EPI_NORMALIZE) != 0;
    const float aff_scale = norm ? 1.0f / ep.std : 1.0f, aff_shift = norm ? -ep.mean / ep.std : 0.0f;
    if (kEpi != kEpiGeneral && im.aligned_out && shear == 0) {
        store_rows_aligned<kEpi == kEpiAffine>(im, ch, row0, col0, acc, obuf, aff_scale, aff_shift);
        return;
    }
    const int nrows = min(kRows, im.H - row0);
    int gcol = col0 - (shear > 0 ? shear * (kRows - 1) : 0);            // first column of row 0 of the block
    float* g = static_cast<float*>(im.dst) + (int64_t)ch * im.dst_cp + (int64_t)row0 * im.dst_rp;   // column 0 of the row
    const float* nz = im.noise ? im.noise + (int64_t)ch * im.dst_cp + (int64_t)row0 * im.dst_rp : nullptr;
    const uint32_t base_phase = (uint32_t)(reinterpret_cast<uintptr_t>(im.dst) >> 2);
    int64_t row_elem = (int64_t)ch * im.dst_cp + (int64_t)row0 * im.dst_rp;
    // two rows per pass: accumulators -> the warp's two row buffers (each skewed so that shared and global addresses
    // agree mod 16 bytes: element x of a row sits at buffer[skew + x]), then 16-byte stores of both rows.
    // Output row q is the .x half of pair q for q < kR and the .y half of pair q - kR otherwise.
#pragma unroll
    for (int r = 0; r < kRows; r += 2) {
        const int gc0 = gcol, gc1 = gcol + shear;
        const int skew0 = (int)((base_phase + (uint32_t)(row_elem + gc0)) & 3u);
        const int skew1 = (int)((base_phase + (uint32_t)(row_elem + im.dst_rp + gc1)) & 3u);
        const int lo0 = max(0, -gc0), hi0 = min(kWarpW, im.W - gc0);
        const int lo1 = max(0, -gc1), hi1 = min(kWarpW, im.W - gc1);
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
        float* g0 = g + gc0;
        float* g1 = g + im.dst_rp + gc1;
        if (kEpi != kEpiGeneral) {
            if (r < nrows && hi0 > lo0) store_row_plain<kEpi == kEpiAffine>(g0, b0, skew0, lo0, hi0, lane, aff_scale, aff_shift);
            if (r + 1 < nrows && hi1 > lo1) store_row_plain<kEpi == kEpiAffine>(g1, b1, skew1, lo1, hi1, lane, aff_scale, aff_shift);
        } else {
            const uint64_t stream = p.philox_offset + (uint64_t)im.philox_slot;
            const float* sm0 = reinterpret_cast<const float*>(__cvta_shared_to_generic(b0));
            if (r < nrows && hi0 > lo0)
                store_row_epilogue(sm0, g0, nz ? nz + gc0 : nullptr, lo0, hi0, skew0, ep, p.philox_seed, stream,
                                   ((uint64_t)ch * im.H + (row0 + r)) * im.W + gc0);
            if (r + 1 < nrows && hi1 > lo1)
                store_row_epilogue(sm0 + kOutPitch, g1, nz ? nz + im.dst_rp + gc1 : nullptr, lo1, hi1, skew1, ep, p.philox_seed, stream,
                                   ((uint64_t)ch * im.H + (row0 + r + 1)) * im.W + gc1);
        }
        __syncwarp();
        g += 2 * im.dst_rp;
        if (nz) nz += 2 * im.dst_rp;
        row_elem += 2 * im.dst_rp;
        gcol += 2 * shear;
    }
}

// ---------------------------------------------------------------- kernel
// kEpi selects the epilogue variant (kEpiNone / kEpiAffine / kEpiGeneral); batches without one run the leanest.
template <int kEpi>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) blur_tiled_kernel(const __grid_constant__ TiledParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    float* obuf_all = reinterpret_cast<float*>(smem + kStages * kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kOutBufBytes);
    uint64_t* full = bars;                  // [kStages] producers -> consumers: stage loaded and patched
    uint64_t* empty = bars + kStages;       // [kStages] consumers -> producers: stage may be refilled
    uint64_t* landed = bars + 2 * kStages;  // [kStages] TMA -> producers: the stage's boxes have arrived, borders may be patched
    int* tile_slots = reinterpret_cast<int*>(bars + 3 * kStages);
    const int warp = threadIdx.x >> 5;

    // Programmatic dependent launch: let the next launch on the stream start filling SMs as this grid's CTAs retire, and
    // -- unless the caller declared this batch independent of the previous launch -- wait for that launch to complete
    // (and flush) before touching global memory.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!p.overlap_prev) asm volatile("griddepcontrol.wait;" ::: "memory");
    DIB_TRACE_EVENT(63, 0);

    __shared__ int s_first_tab[DIB_MAX_BATCH + 1];
    const int* s_first = p.meta_dev != nullptr ? s_first_tab : nullptr;
    tile_table(p, 2, s_first_tab);
    if (threadIdx.x == 0) {
        for (int b = 0; b < kStages; ++b) {
            mbar_init(&full[b], kProducerThreads);      // one cp.async-arrive per producer thread
            mbar_init(&landed[b], 1);                   // one arrive.expect_tx; the boxes complete the transaction bytes
            mbar_init(&empty[b], kComputeWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Register budget: the launch splits the register file evenly over all warps; the producer warpgroup hands most of
    // its share back so that the compute warps can hold kR * kCC accumulators + kR * (kCC + 3) window values per thread.
    // The producers are the LOWEST warp ids: the issue arbiter favours high warp ids, and producer warps only ever wait,
    // issue a copy or patch a border.
    // Stage n lives in buffer n % kStages; the k-th use of a buffer completes phase k of its barriers (parity k & 1).
    if (warp < kProducerWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kProducerRegs));
        // ------------------------------------------------ producer warpgroup
        const int pt = threadIdx.x;
        Stage cur, nxt;
        int nfetch = 0;
        cur.chunk = 0;
        do {        // (device-planned launches skip the tiles of images that belong to another kernel)
            cur.tile = fetch_tile(p, s_first, tile_slots, nfetch, pt);
            if (cur.tile < 0) break;
            decode_tile(p, s_first, cur.tile, cur);
        } while (cur.nchunks == 0);
        if (cur.tile >= 0) {
            cur.rec = load_chunk_rec(p, cur.img, 0);
            issue_stage(p, cur, stage_base(smem_base, 0), &landed[0], pt);
        }
        // Per iteration: look up stage n + 1; with two buffers, if the other one is already free start its loads right away
        // (memory-bound regime: two stages in flight); patch and publish stage n; else start the loads of n + 1 once the
        // consumers have released its buffer.
        for (int n = 0;; ++n) {
            const int b = n % kStages, b1 = (n + 1) % kStages;
            if (cur.tile < 0) {
                // no more work: publish a stop marker in the buffer stage n would have used
                if (n >= kStages) producer_wait(&empty[b], (uint32_t)((n / kStages - 1) & 1), pt);
                if (pt == 0) reinterpret_cast<StageHdr*>(smem + (size_t)b * kStageBytes)->tile = -1;
                __threadfence_block();
                cp_async_mbar_arrive(&full[b]);
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
            next_stage(p, s_first, cur, nxt, tile_slots, nfetch, pt);          // its chunk record is in flight during the work below
            const uint32_t empty_parity = (uint32_t)(((n + 1) / kStages - 1) & 1);
            bool issued = false;
            if (kStages > 1 && nxt.tile >= 0) {
                // thread 0 probes the buffer of stage n + 1 and shares the answer (the flag slot alternates with n)
                int* flag = tile_slots + 2 + (n & 1);
                if (pt == 0) *flag = (n + 1 < kStages || mbar_test(&empty[b1], empty_parity)) ? 1 : 0;
                asm volatile("bar.sync 2, %0;" ::"n"(kProducerThreads) : "memory");
                if (*flag) {
                    issue_stage(p, nxt, stage_base(smem_base, b1), &landed[b1], pt);
                    issued = true;
                }
            }
            finish_stage(p, cur, stage_base(smem_base, b), &landed[b], (uint32_t)((n / kStages) & 1), &full[b], pt);
            if (nxt.tile >= 0 && !issued) {
                DIB_TRACE_EVENT(13, n);
                if (n + 1 >= kStages) producer_wait(&empty[b1], empty_parity, pt);
                DIB_TRACE_EVENT(14, n);
                issue_stage(p, nxt, stage_base(smem_base, b1), &landed[b1], pt);
                DIB_TRACE_EVENT(15, n);
            }
            cur = nxt;
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kComputeRegs));
        // ------------------------------------------------ compute warps
        const int cw = warp - kProducerWarps;
        const int wrow = cw / kWarpCols, wcol = cw % kWarpCols;
        float2 acc[kR][kCC];
        const uint32_t obuf = smem_u32(obuf_all + cw * 2 * kOutPitch);
        for (int n = 0;; ++n) {
            const int b = n % kStages;
            const uint32_t sbase = stage_base(smem_base, b);
            DIB_TRACE_EVENT(0, n);
            mbar_wait(&full[b], (n / kStages) & 1);
            DIB_TRACE_EVENT(1, n);
            StageHdr h;
            {   // explicit vector loads keep the header in registers
                const int4* hp = reinterpret_cast<const int4*>(smem + (size_t)b * kStageBytes);
                const int4 a = hp[0], b4 = hp[1], c2 = hp[2];
                h.tile = a.x; h.img = a.y; h.ch = a.z; h.i0 = a.w;
                h.j0 = b4.x; h.first_chunk = b4.y; h.last_chunk = b4.z; h.dy_hi = b4.w;
                h.dx_hi = c2.x; h.nseg = c2.y; h.group_w = c2.z; h.shear = c2.w;
                const int2 d2 = *reinterpret_cast<const int2*>(hp + 3);
                h.skew0 = d2.x; h.dskew = d2.y;
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
            const int pad = (kRows - 1) * (h.shear < 0 ? -h.shear : h.shear);
            const bool active = row0 < im.H && col0 - pad < im.W;                 // warp-uniform
            if (active) {
                if (h.group_w == 2)
                    compute_chunk<2>(acc, sbase, h.nseg, h.dy_hi, h.dx_hi, h.shear, h.skew0, h.dskew, wrow, wcol);
                else
                    compute_chunk<4>(acc, sbase, h.nseg, h.dy_hi, h.dx_hi, h.shear, h.skew0, h.dskew, wrow, wcol);
            }
            __syncwarp();
            DIB_TRACE_EVENT(4, n);
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[b]);     // this warp is done reading the stage
            if (h.last_chunk && active) store_rows<kEpi>(p, im, h.ch, row0, col0, h.shear, acc, obuf);
            DIB_TRACE_EVENT(5, n);
        }
    }
}

// ---------------------------------------------------------------- host launcher
int tiled_tile_counts(int H, int W, int shear, int* tiles_y, int* tiles_x) {
    const int pad = (kRows - 1) * (shear < 0 ? -shear : shear);
    *tiles_y = (H + kTH - 1) / kTH;
    *tiles_x = (W + pad + kTW - 1) / kTW;
    return *tiles_y * *tiles_x;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;      // resolved once per process through the runtime (no link against libcuda)
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
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
        DIB_CUDA(cudaFuncSetAttribute(blur_tiled_kernel<kEpiNone>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        DIB_CUDA(cudaFuncSetAttribute(blur_tiled_kernel<kEpiAffine>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        DIB_CUDA(cudaFuncSetAttribute(blur_tiled_kernel<kEpiGeneral>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set_dev = dev;
    }
    if (io_dtype != DIB_F32) {
        set_error("dib_blur_batch: the tiled kernel takes float32 images");
        return DIB_ERR_UNSUPPORTED;
    }
    const EncodeTiledFn encode = encode_tiled_fn();
    if (encode == nullptr) {
        set_error("dib_blur_batch: cuTensorMapEncodeTiled is not available from this driver");
        return DIB_ERR_CUDA;
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
        t.shear = m.prog_shear;
        const int per_ch = tiled_tile_counts(im.H, im.W, t.shear, &t.tiles_y, &t.tiles_x);
        t.first_tile = total;
        t.psf_index = im.psf_index;
        t.nchunks = m.prog_chunks;
        t.epilogue = im.epilogue;
        t.zero_pad = (im.pad_mode == DIB_PAD_ZERO128);
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
        // 1-D tensor map over the image's elements: base = the 16-byte-aligned address at or below src, extent = up to the
        // image's last element; a box is 256 consecutive elements starting at any coordinate (zero fill outside)
        const uintptr_t src_addr = reinterpret_cast<uintptr_t>(im.src);
        const uintptr_t base = src_addr & ~uintptr_t(15);
        t.src_off = (int)((src_addr - base) >> 2);
        const uint64_t span = (uint64_t)t.src_off + (uint64_t)(im.C - 1) * (uint64_t)im.src_chan_pitch +
                              (uint64_t)(im.H - 1) * (uint64_t)im.src_row_pitch + (uint64_t)im.W;
        if (span >= (1ull << 31)) {
            set_error("dib_blur_batch: image %d spans %llu elements, more than a 1-D tensor map's signed coordinates address", order[k],
                      (unsigned long long)span);
            return DIB_ERR_UNSUPPORTED;
        }
        const cuuint64_t gdim[1] = {(cuuint64_t)span};
        const cuuint64_t gstride[1] = {0};
        const cuuint32_t box[1] = {kBoxElems};
        const cuuint32_t estride[1] = {1};
        const CUresult cr = encode(&t.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, reinterpret_cast<void*>(base), gdim, gstride, box, estride,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            set_error("dib_blur_batch: cuTensorMapEncodeTiled failed with %d for image %d", (int)cr, order[k]);
            return DIB_ERR_CUDA;
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
    // Overlapped launches share SMs only while the earlier grid drains: with one CTA per SM and a full grid at most two
    // launches are ever co-resident, which the four scheduler slots cover.  A grid smaller than the machine could be
    // co-resident with many successors, so it always orders itself after its predecessor.
    const int max_ctas = sm_count * kCtasPerSm;
    p.overlap_prev = (overlap_prev && total >= max_ctas) ? 1 : 0;
    const int grid = total < max_ctas ? total : max_ctas;     // persistent: kCtasPerSm CTAs per SM
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
    if (any_general) {
        DIB_CUDA(cudaLaunchKernelEx(&cfg, blur_tiled_kernel<kEpiGeneral>, p));
    } else if (any_epi) {
        DIB_CUDA(cudaLaunchKernelEx(&cfg, blur_tiled_kernel<kEpiAffine>, p));
    } else {
        DIB_CUDA(cudaLaunchKernelEx(&cfg, blur_tiled_kernel<kEpiNone>, p));
    }
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

}  // namespace dib

#ifdef DIB_TRACE
// kernel experiments: copy out (and reset) the timeline recorded by CTA 0 of the launches since the last call
extern "C" __attribute__((visibility("default"))) int dib_debug_trace(unsigned long long* out, int max_events) {
    unsigned int n[16];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(n, dib::g_trace_n, sizeof(n));
    int total = 0;
    for (int w = 0; w < 16; ++w) {
        int c = (int)n[w];
        if (c > dib::kTracePerWarp) c = dib::kTracePerWarp;
        if (total + c > max_events) c = max_events - total;
        if (c > 0) cudaMemcpyFromSymbol(out + total, dib::g_trace, sizeof(unsigned long long) * c, sizeof(unsigned long long) * w * dib::kTracePerWarp);
        total += c > 0 ? c : 0;
    }
    return total;
}
#endif

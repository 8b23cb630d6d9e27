// Tap compaction: dense PSF batch -> per-PSF normalised (y, x, w) lists in the reference's accumulation order,
// extents / PCA moments, and the column-group program the tiled blur kernel executes.
//
// Replaces (per PSF, on the host thread, with device->host syncs) in the reference:
//   psf_GPU = psf_GPU / psf_GPU.sum()                       models/blur_functions.py:98
//   non_zero_points = psf_GPU.nonzero(as_tuple=False)       models/blur_functions.py:63
//   non_zero_points[i, 0] - 63, psf_GPU[y, x] per tap       models/blur_functions.py:67   (2 syncs per tap)
//   min/max of the nonzero coordinates                      utils.py:372-380 (expand_targets)
//   first/second moments of the support                     transforms.py:366-376 (PSF PCA)
// One CTA per PSF, one launch per batch.  PSFs up to 129 x 129 are first staged in shared memory with independent
// (unrolled) loads, so the three passes over the cells and the program builder pay one round of global latency
// instead of one per loop iteration; larger canvases (the 256 branch) read global memory throughout.
#include "masked_common.cuh"

namespace dib {

constexpr int kCompactThreads = 1024;

template <typename T>
struct PsfNum;
template <>
struct PsfNum<float> {
    // fp32 PSF: torch sums in fp32 and divides with IEEE division.  The sum is accumulated in fp64 here and
    // rounded once; for PSFs on the fp16 grid summing below 1 (every stored / generated PSF) every fp32
    // partial sum is exact, so this equals torch's result whatever its reduction order.
    __device__ static float load(const float* p, int64_t i) { return p[i]; }
    __device__ static float round_sum(double s) { return (float)s; }
    __device__ static float normalized(float v, float s) { return __fdiv_rn(v, s); }
};
template <>
struct PsfNum<__half> {
    // half PSF: torch's CUDA sum accumulates in fp32 and rounds to half; half / half divides in fp32 and rounds.
    __device__ static float load(const __half* p, int64_t i) { return __half2float(p[i]); }
    __device__ static float round_sum(double s) { return __half2float(__float2half_rn((float)s)); }
    __device__ static float normalized(float v, float s) { return __half2float(__float2half_rn(__fdiv_rn(v, s))); }
};

#ifdef DIB_COMPACT_TIMING       // experiment build: phase timestamps of block 0 (tools/exp/time_compact.py)
__device__ unsigned long long g_compact_t[16];
#define DIB_CT(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_compact_t[k] = clock64(); } while (0)
#else
#define DIB_CT(k)
#endif

static_assert(kCompactThreads == 1024, "the two-level reductions below assume 32 warps");

// Block-wide fp64 sum: shuffle tree inside each warp, then the same tree over the 32 warp totals (a fixed order, so the
// result is deterministic), one barrier pair instead of a 32-term serial loop in every thread.
__device__ __forceinline__ double block_sum_double(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = sh[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

constexpr int kMaxGroups = 192;       // sheared groups of one PSF (support width + shear drift, in groups of 2)
constexpr int kMaxBands = 64;
#ifndef DIB_MASKED_MAX_CHUNKS
#define DIB_MASKED_MAX_CHUNKS 4
#endif
constexpr int kMaskedMaxChunks = DIB_MASKED_MAX_CHUNKS;      // PSFs whose masked program needs more chunks take the dense sheared kernel
constexpr int kNumShears = 2 * kShearMax + 1;
constexpr int kNumCand = 2 * kNumShears;          // group width 2 or 4 x shear -kShearMax .. kShearMax

// Modelled cost (cycles of one compute warp) of one dense step / one window fill of a segment, for ranking the
// candidates: a step is bound by its 2 * G * kR * kCC FMA-pipe cycles or by the shared-memory traffic of its window row
// (4 warps share the load pipe), a fill is shared-memory traffic only.
__host__ __device__ constexpr int step_cost(int G) {
    const int fma = 2 * G * kR * kCC, lds = 4 * (kCC + G);
    return (fma > lds ? fma : lds) + 4;
}
__host__ __device__ constexpr int fill_cost(int G) { return 4 * (kRows - 1) * (kCC + G - 1) + 60; }
constexpr int kChunkCost = 250;

// 128-bit row masks (bit y of word y >> 5) travel as uint4 VALUES: with pointers to thread-local arrays the dynamic word
// index put them in local memory, and the single threads that walk the bands spent most of their time there.
__device__ __forceinline__ unsigned mask_word(const uint4& m, int k) { return k == 0 ? m.x : k == 1 ? m.y : k == 2 ? m.z : m.w; }
__device__ __forceinline__ uint4 mask_load(const unsigned* p) { return make_uint4(p[0], p[1], p[2], p[3]); }

// first set bit at position >= from, or -1
__device__ __forceinline__ int mask_first_from(const uint4& m, int from) {
    from = max(from, 0);
    int r = -1;
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        unsigned w = mask_word(m, k);
        if (k == (from >> 5)) w &= 0xffffffffu << (from & 31);
        if (k < (from >> 5)) w = 0u;
        if (w) r = k * 32 + __ffs(w) - 1;
    }
    return r;
}

// first and last set bit within [y0, y1] (f = -1 when none)
__device__ __forceinline__ void mask_range_first_last(const uint4& m, int y0, int y1, int& f, int& l) {
    f = -1;
    l = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned w = mask_word(m, k);
        if (k < (y0 >> 5) || k > (y1 >> 5)) w = 0u;
        if (k == (y0 >> 5)) w &= 0xffffffffu << (y0 & 31);
        if (k == (y1 >> 5)) w &= 0xffffffffu >> (31 - (y1 & 31));
        if (w) {
            if (f < 0) f = k * 32 + __ffs(w) - 1;
            l = k * 32 + 31 - __clz(w);
        }
    }
}

constexpr int kStageMaxCells = 129 * 129;      // staged PSFs: 66 564 B of dynamic shared memory

// cell i of the PSF as float: from the shared-memory copy when staged, else from global memory
template <typename T, bool kStaged>
__device__ __forceinline__ float psf_cell(const T* psf, const float* staged, int64_t i) {
    if constexpr (kStaged)
        return staged[i];
    else
        return PsfNum<T>::load(psf, i);
}

namespace mk {
constexpr int kMaxBands = (32 + kChunkGroups - 1) / kChunkGroups;   // <= 32 groups of kGroupW columns in a 128-wide PSF
constexpr int kBandMaxChunks = 8;                                   // 128 rows / (kChunkHaloRows + 1) rounded up

// Program of the masked tiled kernel (layout in masked_common.cuh, consumer blur_masked.cu): unsheared groups of 4 columns,
// a 4-wide weight vector per row, zero where the PSF has no tap (the kernel skips those).  Block-wide; returns through
// shared memory: chunks (<= 0: none built), weight vectors, segments.
template <typename T, bool kStaged>
__device__ void build_program(const T* psf, const float* sh_psf, int side, int normalize, float s, bool s_finite, int centre, int ymin, int ymax,
                              int xmin, int xmax, uint8_t* my_prog, int max_chunks, int& out_chunks_n, int& out_steps, int& out_segs) {
    __shared__ unsigned sh_occ[32 * 4];
    __shared__ ChunkRec sh_chunks[kProgMaxChunks];
    __shared__ SegRec sh_segs[kProgMaxChunks * kChunkGroups];
    __shared__ ChunkRec sh_band_chunks[kMaxBands * kBandMaxChunks];
    __shared__ SegRec sh_band_segs[kMaxBands * kBandMaxChunks * kChunkGroups];
    __shared__ int sh_band_count[kMaxBands];
    __shared__ int sh_nchunks, sh_nsegs, sh_total_steps;
    const int tid = threadIdx.x;
    ChunkRec* out_chunks = reinterpret_cast<ChunkRec*>(my_prog);
    if (tid == 0) {
        sh_nchunks = 0;
        sh_nsegs = 0;
        sh_total_steps = 0;
    }
    __syncthreads();
    const int ngroups = (xmax - xmin + kGroupW) / kGroupW;   // <= 32 for side <= 129
    const int nrows_box = ymax - ymin + 1;
    // 3a. per group: bitmask of the PSF rows holding a tap in the group's columns (rows 0..127 -> 4 words)
    for (int k = tid; k < ngroups * 4; k += kCompactThreads) sh_occ[k] = 0u;
    __syncthreads();
    for (int k = tid; k < ngroups * nrows_box; k += kCompactThreads) {
        const int g = k / nrows_box, y = ymin + k % nrows_box;
        bool any = false;
        for (int e = 0; e < kGroupW; ++e) {
            const int x = xmin + g * kGroupW + e;
            if (x < side) {
                const float v = psf_cell<T, kStaged>(psf, sh_psf, (int64_t)y * side + x);
                const float w = (normalize && (v != 0.0f || !s_finite)) ? PsfNum<T>::normalized(v, s) : v;   // 0 / s is 0 for a finite nonzero sum
                any |= (w != 0.0f);
            }
        }
        if (any) atomicOr(&sh_occ[g * 4 + (y >> 5)], 1u << (y & 31));
    }
    __syncthreads();
    // 3b. cut the support into chunks of segments: one thread per band of kChunkGroups groups walks that band's rows
    //     with 128-bit masks (find-first-set), then thread 0 concatenates the bands' chunk lists in band order
    const int nbands = (ngroups + kChunkGroups - 1) / kChunkGroups;
    if (tid < nbands) {
        const int g0 = tid * kChunkGroups, g1 = min(g0 + kChunkGroups, ngroups);
        uint4 band = make_uint4(0u, 0u, 0u, 0u);
        for (int g = g0; g < g1; ++g) {
            const uint4 o = mask_load(&sh_occ[g * 4]);
            band.x |= o.x; band.y |= o.y; band.z |= o.z; band.w |= o.w;
        }
        int cursor = ymin, nb = 0;
        while (nb < kBandMaxChunks) {
            int y0 = mask_first_from(band, cursor);
            if (y0 < 0) break;
            const int y1 = min(y0 + kChunkHaloRows, ymax);
            ChunkRec c;
            int nseg = 0, lo = 1 << 20, hi = -(1 << 20), xlo = 1 << 20, xhi = -(1 << 20);
            for (int g = g0; g < g1; ++g) {
                int f, l;
                mask_range_first_last(mask_load(&sh_occ[g * 4]), y0, y1, f, l);
                if (f < 0) continue;
                SegRec sg;
                sg.dx0 = (int16_t)(xmin + g * kGroupW - centre);
                sg.dy0 = (int16_t)(f - centre);
                sg.nsteps = (int16_t)(l - f + 1);
                sg.woff = 0;
                lo = min(lo, f - centre);
                hi = max(hi, l - centre);
                xlo = min(xlo, (int)sg.dx0);
                xhi = max(xhi, (int)sg.dx0 + kGroupW - 1);
                sh_band_segs[(tid * kBandMaxChunks + nb) * kChunkGroups + nseg] = sg;
                ++nseg;
            }
            c.dy_lo = (int16_t)lo; c.dy_hi = (int16_t)hi;
            c.dx_lo = (int16_t)xlo; c.dx_hi = (int16_t)xhi;
            c.nseg = (int16_t)nseg; c.wsteps = 0;
            c.data_off = 0;
            sh_band_chunks[tid * kBandMaxChunks + nb] = c;
            ++nb;
            cursor = y1 + 1;
        }
        // more rows left than kBandMaxChunks chunks can cover: no program (the generic kernel takes the PSF)
        sh_band_count[tid] = (nb == kBandMaxChunks && mask_first_from(band, cursor) >= 0) ? -1 : nb;
    }
    __syncthreads();
    if (tid == 0) {
        int nchunks = 0, nsegs = 0;
        bool ok = true;
        for (int bnd = 0; bnd < nbands && ok; ++bnd) {
            const int nb = sh_band_count[bnd];
            if (nb < 0 || nchunks + nb > kProgMaxChunks) { ok = false; break; }
            for (int k = 0; k < nb; ++k) {
                sh_chunks[nchunks] = sh_band_chunks[bnd * kBandMaxChunks + k];
                for (int q = 0; q < kChunkGroups; ++q)
                    sh_segs[nchunks * kChunkGroups + q] = sh_band_segs[(bnd * kBandMaxChunks + k) * kChunkGroups + q];
                nsegs += sh_chunks[nchunks].nseg;
                ++nchunks;
            }
        }
        sh_nchunks = ok ? nchunks : -1;
        sh_nsegs = nsegs;
    }
    __syncthreads();
    int nchunks = sh_nchunks;
    if (nchunks > max_chunks) {     // the caller will not use a program this long: skip the offsets and the weight vectors
        out_chunks_n = nchunks;
        out_steps = 0;
        out_segs = 0;
        return;
    }
    // 3c. offsets: weight vectors of a chunk's segments are laid out back to back
    if (nchunks > 0) {
        if (tid == 0) {
            int data_off = (int)kProgHeaderBytes, total_steps = 0;
            bool ok = true;
            for (int ci = 0; ci < nchunks && ok; ++ci) {
                int nw = 0;
                for (int sgi = 0; sgi < sh_chunks[ci].nseg; ++sgi) {
                    sh_segs[ci * kChunkGroups + sgi].woff = (int16_t)nw;
                    nw += sh_segs[ci * kChunkGroups + sgi].nsteps;
                }
                total_steps += nw;
                const int bytes = kChunkSegBytes + kStepBytes * (nw + 1);
                if (nw > kChunkMaxSteps || data_off + bytes > (int)kProgBytes) { ok = false; break; }
                sh_chunks[ci].wsteps = (int16_t)nw;
                sh_chunks[ci].data_off = data_off;
                data_off += bytes;
            }
            if (!ok) sh_nchunks = -1;
            sh_total_steps = total_steps;
        }
        __syncthreads();
        nchunks = sh_nchunks;
    }
    // 3d. write chunk records, segment records and weight vectors
    if (nchunks > 0) {
        for (int k = tid; k < nchunks; k += kCompactThreads) out_chunks[k] = sh_chunks[k];
        for (int ci = 0; ci < nchunks; ++ci) {
            const ChunkRec c = sh_chunks[ci];
            SegRec* seg_out = reinterpret_cast<SegRec*>(my_prog + c.data_off);
            float* wout = reinterpret_cast<float*>(my_prog + c.data_off + kChunkSegBytes);
            if (tid < kChunkSegBytes / (int)sizeof(SegRec)) {
                SegRec sg;
                sg.dx0 = 0; sg.dy0 = 0; sg.nsteps = 0; sg.woff = 0;
                if (tid < c.nseg) sg = sh_segs[ci * kChunkGroups + tid];
                seg_out[tid] = sg;
            }
            if (tid < kGroupW) wout[c.wsteps * kGroupW + tid] = 0.0f;     // the vector the kernel prefetches past the end
            for (int sgi = 0; sgi < c.nseg; ++sgi) {
                const SegRec sg = sh_segs[ci * kChunkGroups + sgi];
                for (int k = tid; k < sg.nsteps * kGroupW; k += kCompactThreads) {
                    const int step = k / kGroupW, e = k % kGroupW;
                    const int x = sg.dx0 + centre + e, y = sg.dy0 + centre + step;
                    float w = 0.0f;
                    if (x < side) {
                        const float v = psf_cell<T, kStaged>(psf, sh_psf, (int64_t)y * side + x);
                        w = (normalize && (v != 0.0f || !s_finite)) ? PsfNum<T>::normalized(v, s) : v;   // 0 / s is 0 for a finite nonzero sum
                    }
                    wout[(sg.woff + step) * kGroupW + e] = w;
                }
            }
        }
    }
    __syncthreads();
    out_chunks_n = sh_nchunks;
    out_steps = sh_total_steps;
    out_segs = sh_nsegs;
}
}  // namespace mk

template <typename T, bool kStaged>
__global__ void __launch_bounds__(kCompactThreads)
compact_taps_kernel(const T* __restrict__ psfs, int side, int64_t psf_stride, int flags_in, dib_psf_meta* __restrict__ meta,
                    dib_tap* __restrict__ taps, int max_taps, uint8_t* __restrict__ prog, SchedWords* __restrict__ sched) {
    __shared__ double sh_d[kCompactThreads / 32];
    __shared__ int sh_min4[4];                       // ymin, -ymax, xmin, -xmax of the taps
    __shared__ unsigned long long sh_sum6[6];        // support and first / second moments of the positive cells
    __shared__ int sh_warp_count[kCompactThreads / 32];
    __shared__ int sh_running;
    __shared__ unsigned sh_occ[kMaxGroups * 4];
    __shared__ ChunkRec sh_chunks[kProgMaxChunks];
    __shared__ SegRec sh_segs[kProgMaxChunks * kChunkSegSlots];
    __shared__ int sh_band_count[kMaxBands], sh_band_base[kMaxBands];
    __shared__ int sh_first[kNumCand][kMaxGroups], sh_last[kNumCand][kMaxGroups];
    __shared__ int sh_xpmin[kNumShears], sh_xpmax[kNumShears];
    __shared__ unsigned int sh_cost[kNumCand];      // modelled cycles (< 2^22 per candidate; 2^30 marks "does not fit"): 32-bit
                                                    // shared atomics are native, 64-bit ones are CAS loops that crawl under contention
    __shared__ int sh_nchunks, sh_nsegs, sh_total_steps, sh_choice;

    DIB_CT(0);
    const int normalize = flags_in & DIB_COMPACT_NORMALIZE;
    const int n = blockIdx.x;
    if (n == 0 && threadIdx.x == 0) {
        for (int k = 0; k < 2 * kSchedSlots; ++k) {      // masked kernel: slots 0-3, dense kernel: slots 4-7
            sched[k].next_tile = 0u;
            sched[k].done_ctas = 0u;
        }
    }
    const T* psf = psfs + (int64_t)n * psf_stride;
    const int cells = side * side;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 4) sh_min4[tid] = 1 << 20;
    if (tid >= 32 && tid < 38) sh_sum6[tid - 32] = 0ull;        // (published by the barriers of the sum below)
    extern __shared__ float sh_psf[];
    if constexpr (kStaged) {
#pragma unroll 16
        for (int i = tid; i < cells; i += kCompactThreads) sh_psf[i] = PsfNum<T>::load(psf, i);
        __syncthreads();
    }

    DIB_CT(1);
    // 1. psf.sum() (blur_functions.py:98)
    double part = 0.0;
    for (int i = tid; i < cells; i += kCompactThreads) part += (double)psf_cell<T, kStaged>(psf, sh_psf, i);
    const double total = block_sum_double(part, sh_d);
    const float s = PsfNum<T>::round_sum(total);
    const bool s_finite = s != 0.0f && fabsf(s) <= 3.0e38f;     // false for 0, inf and NaN sums: then every cell is divided, as torch does

    DIB_CT(2);
    // 2. ordered compaction of the normalised PSF (row-major nonzero order, blur_functions.py:63).  Every warp owns a
    //    contiguous slab of the PSF: pass 1 counts its taps (ballot + popc, no block barrier), one scan over the 8 warp
    //    totals gives each slab its output offset, pass 2 writes the taps in order.
    int ymin = 1 << 20, ymax_neg = 1 << 20, xmin = 1 << 20, xmax_neg = 1 << 20;  // max kept as min of negatives
    // per-thread moment sums stay in 32 bits (a thread sees <= slab / 32 cells of coordinates < 512) and widen at the reduction
    int support_i = 0, sy_i = 0, sx_i = 0, syy_i = 0, sxx_i = 0, sxy_i = 0;
    const int slab = ((cells + kCompactThreads / 32 - 1) / (kCompactThreads / 32) + 31) & ~31;
    const int slab_lo = warp * slab, slab_hi = min(cells, slab_lo + slab);
    const int lg_side = (side & (side - 1)) == 0 ? 31 - __clz(side) : -1;       // row of a cell by shift when side is 2^k
    // (loops kept rolled: the kernel runs every instruction about once, so code size is fetch time)
    int warp_count = 0;
#pragma unroll 1
    for (int base = slab_lo; base < slab_hi; base += 32) {
        const int i = base + lane;
        float v = 0.0f, w = 0.0f;
        if (i < slab_hi) {
            v = psf_cell<T, kStaged>(psf, sh_psf, i);
            w = (normalize && (v != 0.0f || !s_finite)) ? PsfNum<T>::normalized(v, s) : v;   // 0 / s is 0 for a finite nonzero sum
        }
        const bool nz = (w != 0.0f);
        warp_count += __popc(__ballot_sync(0xffffffffu, nz));
        if (nz || v > 0.0f) {
            const int y = lg_side >= 0 ? i >> lg_side : i / side, x = i - y * side;
            if (nz) {
                ymin = min(ymin, y);
                ymax_neg = min(ymax_neg, -y);
                xmin = min(xmin, x);
                xmax_neg = min(xmax_neg, -x);
            }
            if (v > 0.0f) {  // support of the PCA: psf > 0 (transforms.py:366)
                support_i += 1;
                sy_i += y;
                sx_i += x;
                syy_i += y * y;
                sxx_i += x * x;
                sxy_i += y * x;
            }
        }
    }
    if (lane == 0) sh_warp_count[warp] = warp_count;
    __syncthreads();
    DIB_CT(3);
    int offset = 0, total_count = 0;
    {
        const int c = sh_warp_count[lane];          // 32 warps: exclusive prefix over the warp totals by shuffle
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        total_count = __shfl_sync(0xffffffffu, incl, 31);
        offset = __shfl_sync(0xffffffffu, incl - c, warp);
    }
    if (warp_count != 0) {      // warp-uniform: most slabs of a 128 x 128 container hold no tap at all
#pragma unroll 1
        for (int base = slab_lo; base < slab_hi; base += 32) {
            const int i = base + lane;
            float w = 0.0f;
            if (i < slab_hi) {
                const float v = psf_cell<T, kStaged>(psf, sh_psf, i);
                w = (normalize && (v != 0.0f || !s_finite)) ? PsfNum<T>::normalized(v, s) : v;
            }
            const bool nz = (w != 0.0f);
            const unsigned ballot = __ballot_sync(0xffffffffu, nz);
            const int pos = offset + __popc(ballot & ((1u << lane) - 1u));
            if (nz && pos < max_taps) {
                const int y = lg_side >= 0 ? i >> lg_side : i / side, x = i - y * side;
                dib_tap t;
                t.y = (int16_t)y;
                t.x = (int16_t)x;
                t.w = w;
                taps[(int64_t)n * max_taps + pos] = t;
            }
            offset += __popc(ballot);
        }
    }
    if (tid == 0) sh_running = total_count;
    __syncthreads();
    DIB_CT(4);
    const int count = sh_running;
    // Extents and moments (integer sums and minima: exact in any order): only the warps whose slab holds a tap or a positive
    // cell take part.
    if (warp_count != 0 || __any_sync(0xffffffffu, support_i != 0)) {       // warp-uniform
        // shuffle trees inside the warp (32-bit: a warp's slab of a container up to 512 x 512 sums below 2^31 only for the
        // first moments, so the second moments go 64-bit when the container is larger than 256 x 256), then lane 0 adds the
        // warp's totals to the block's
        int m4[4] = {ymin, ymax_neg, xmin, xmax_neg};
        long long s6[6] = {support_i, sy_i, sx_i, syy_i, sxx_i, sxy_i};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) m4[k] = min(m4[k], __shfl_xor_sync(0xffffffffu, m4[k], o));
#pragma unroll
            for (int k = 0; k < 3; ++k) s6[k] = (int)s6[k] + __shfl_xor_sync(0xffffffffu, (int)s6[k], o);
            if (side > 256) {
#pragma unroll
                for (int k = 3; k < 6; ++k) s6[k] += __shfl_xor_sync(0xffffffffu, s6[k], o);
            } else {
#pragma unroll
                for (int k = 3; k < 6; ++k) s6[k] = (int)s6[k] + __shfl_xor_sync(0xffffffffu, (int)s6[k], o);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) atomicMin(&sh_min4[k], m4[k]);
#pragma unroll
            for (int k = 0; k < 6; ++k) atomicAdd(&sh_sum6[k], (unsigned long long)s6[k]);
        }
    }
    __syncthreads();
    ymin = sh_min4[0];
    const int ymax = -sh_min4[1];
    xmin = sh_min4[2];
    const int xmax = -sh_min4[3];
    long long sums6[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) sums6[k] = (long long)sh_sum6[k];
    const long long support = sums6[0], sy = sums6[1], sx = sums6[2], syy = sums6[3], sxx = sums6[4], sxy = sums6[5];

    DIB_CT(5);
    // 3. program for the tiled kernel: dense sheared column groups (layout in dib_common.cuh, consumer blur_tiled.cu)
    const int centre = side > 129 ? 127 : 63;
    int flags = 0;
    if (count > max_taps) flags |= DIB_META_TRUNCATED;
    uint8_t* my_prog = prog + (size_t)n * kProgBytes;
    ChunkRec* out_chunks = reinterpret_cast<ChunkRec*>(my_prog);
    // taps at dy or dx >= 64 (PSF row/col >= 127) hit torch.roll's wrap-around at output row/col 0; those PSFs
    // (never produced by the centred generator) stay on the generic kernel, which restates the wrap exactly.
    const bool want_prog = side <= 129 && count > 0 && ymax <= 126 && xmax <= 126;
    if (tid == 0) {
        sh_nchunks = -1;
        sh_nsegs = 0;
        sh_total_steps = 0;
        sh_choice = -1;
    }
    int prog_G = 0, prog_k = 0;
    __syncthreads();
    // Which tiled kernel takes the PSF is decided here, by measurement (tools/exp/route_probe.py, 8 x 3x800x1333 per cell of
    // the eval sweep): the masked kernel is the faster one for every PSF its program holds in up to kMaskedMaxChunks chunks
    // (all exposures up to 1/2, full exposure at param 0.005 / 0.001); the long streaks of param 0.00005 at full exposure,
    // which it cuts into 5-6 chunks, run 5 % faster on the dense sheared kernel.
    // A masked chunk spans at most kChunkHaloRows + 1 tap rows and kChunkGroups * kGroupW tap columns: a support box that
    // cannot fit kMaskedMaxChunks of them goes to the dense builder directly (saves the masked build, ~40 us for one PSF).
    const int mk_bound = max((ymax - ymin + mk::kChunkHaloRows + 1) / (mk::kChunkHaloRows + 1),
                             (xmax - xmin + mk::kChunkGroups * mk::kGroupW) / (mk::kChunkGroups * mk::kGroupW));
    const bool small_psf = want_prog && !(flags_in & DIB_COMPACT_DENSE_ONLY) &&
                           (mk_bound <= kMaskedMaxChunks || (flags_in & DIB_COMPACT_MASKED_ONLY));
    int mk_chunks = 0, mk_steps = 0, mk_segs = 0;
    if (small_psf) mk::build_program<T, kStaged>(psf, sh_psf, side, normalize, s, s_finite, centre, ymin, ymax, xmin, xmax, my_prog,
                                                    (flags_in & DIB_COMPACT_MASKED_ONLY) ? (1 << 20) : kMaskedMaxChunks, mk_chunks, mk_steps, mk_segs);
    DIB_CT(6);
    const bool masked_ok = small_psf && mk_chunks > 0 && (mk_chunks <= kMaskedMaxChunks || (flags_in & DIB_COMPACT_MASKED_ONLY));
    if (want_prog && !masked_ok) {
        const int nrows_box = ymax - ymin + 1;
        // 3a. rank the candidates (group width 2 / 4) x (shear -kShearMax .. kShearMax) by a cost model: per sheared group
        //     the span of rows it occupies -> dense steps, window fills and chunks.
        if (tid < kNumShears) {
            sh_xpmin[tid] = 1 << 20;
            sh_xpmax[tid] = -(1 << 20);
        }
        if (tid < kNumCand) sh_cost[tid] = 0u;
        for (int k = tid; k < kNumCand * kMaxGroups; k += kCompactThreads) {
            (&sh_first[0][0])[k] = 1 << 20;
            (&sh_last[0][0])[k] = -1;
        }
        __syncthreads();
        DIB_CT(9);
        const int box_lo = ymin * side, box_hi = (ymax + 1) * side;       // the support's rows only
        for (int i = box_lo + tid; i < box_hi; i += kCompactThreads) {
            const float v = psf_cell<T, kStaged>(psf, sh_psf, i);
            const float w = (normalize && (v != 0.0f || !s_finite)) ? PsfNum<T>::normalized(v, s) : v;   // 0 / s is 0 for a finite nonzero sum
            if (w != 0.0f) {
                const int y = i / side, x = i - y * side;
#pragma unroll
                for (int q = 0; q < kNumShears; ++q) {
                    const int xp = x - (q - kShearMax) * (y - ymin);
                    atomicMin(&sh_xpmin[q], xp);
                    atomicMax(&sh_xpmax[q], xp);
                }
            }
        }
        __syncthreads();
        for (int i = box_lo + tid; i < box_hi; i += kCompactThreads) {
            const float v = psf_cell<T, kStaged>(psf, sh_psf, i);
            const float w = (normalize && (v != 0.0f || !s_finite)) ? PsfNum<T>::normalized(v, s) : v;   // 0 / s is 0 for a finite nonzero sum
            if (w != 0.0f) {
                const int y = i / side, x = i - y * side;
#pragma unroll
                for (int q = 0; q < kNumShears; ++q) {
                    const int xp = x - (q - kShearMax) * (y - ymin) - sh_xpmin[q];
#pragma unroll
                    for (int gi = 0; gi < 2; ++gi) {
                        const int g = xp >> (gi + 1);               // group width 2 << gi
                        if (g < kMaxGroups) {
                            atomicMin(&sh_first[gi * kNumShears + q][g], y);
                            atomicMax(&sh_last[gi * kNumShears + q][g], y);
                        }
                    }
                }
            }
        }
        __syncthreads();
        DIB_CT(10);
        for (int k = tid; k < kNumCand * kMaxGroups; k += kCompactThreads) {
            const int cand = k / kMaxGroups, g = k - cand * kMaxGroups;
            const int gi = cand / kNumShears, q = cand - gi * kNumShears, G = 2 << gi;
            const int ngroups = ((sh_xpmax[q] - sh_xpmin[q]) >> (gi + 1)) + 1;
            unsigned int c = 0u;
            if (ngroups > kMaxGroups) {
                c = g == 0 ? (1u << 30) : 0u;                        // candidate does not fit the builder's tables
            } else if (g < ngroups) {
                const int f = sh_first[cand][g], l = sh_last[cand][g];
                if (l >= f) {
                    const int span = l - f + 1;
                    c = (unsigned int)(span * step_cost(G) + ((span + kChunkTapRows - 1) / kChunkTapRows) * fill_cost(G));
                }
                if (g == 0) {       // staging cost: bands x row windows
                    const int nb = band_cols(q - kShearMax) / G;
                    c += (unsigned int)(((ngroups + nb - 1) / nb) * ((nrows_box + kChunkTapRows - 1) / kChunkTapRows) * kChunkCost);
                }
            }
            if (c) atomicAdd(&sh_cost[cand], c);
        }
        __syncthreads();
        if (tid == 0) {
            int bestc = 0;
            for (int c = 1; c < kNumCand; ++c)
                if (sh_cost[c] < sh_cost[bestc]) bestc = c;
            sh_choice = bestc;
        }
        __syncthreads();
        DIB_CT(11);
        // 3b-3d. build with the chosen candidate; if it overflows a table, once more with unsheared groups of 4
        for (int attempt = 0; attempt < 2; ++attempt) {
            const int cand = attempt == 0 ? sh_choice : (1 * kNumShears + kShearMax);
            const int gi = cand / kNumShears, q = cand - gi * kNumShears;
            const int G = 2 << gi, shear = q - kShearMax;
            const int xp0 = sh_xpmin[q];
            const int ngroups = ((sh_xpmax[q] - xp0) >> (gi + 1)) + 1;
            const int nb = band_cols(shear) / G;                     // groups per band
            const int nbands = (ngroups + nb - 1) / nb;
            bool fits = ngroups <= kMaxGroups && nbands <= kMaxBands;
            __syncthreads();
            if (fits) {
                // per group: bitmask of the PSF rows holding a tap in the group's (sheared) columns
                for (int k = tid; k < ngroups * 4; k += kCompactThreads) sh_occ[k] = 0u;
                __syncthreads();
                for (int k = tid; k < ngroups * nrows_box; k += kCompactThreads) {
                    const int g = k / nrows_box, y = ymin + k % nrows_box;
                    bool any = false;
                    for (int e = 0; e < G; ++e) {
                        const int x = xp0 + g * G + shear * (y - ymin) + e;
                        if (x >= 0 && x < side) {
                            const float v = psf_cell<T, kStaged>(psf, sh_psf, (int64_t)y * side + x);
                            const float w = (normalize && (v != 0.0f || !s_finite)) ? PsfNum<T>::normalized(v, s) : v;   // 0 / s is 0 for a finite nonzero sum
                            any |= (w != 0.0f);
                        }
                    }
                    if (any) atomicOr(&sh_occ[g * 4 + (y >> 5)], 1u << (y & 31));
                }
                __syncthreads();
                DIB_CT(12);
                // one thread per band walks the band's rows in windows of kChunkTapRows: pass 0 counts the chunks, pass 1
                // (after a prefix sum over the bands) writes chunk and segment records
                for (int pass = 0; pass < 2; ++pass) {
                    if (tid < nbands) {
                        const int g0 = tid * nb, g1 = min(g0 + nb, ngroups);
                        uint4 band = make_uint4(0u, 0u, 0u, 0u);
                        for (int g = g0; g < g1; ++g) {
                            const uint4 o = mask_load(&sh_occ[g * 4]);
                            band.x |= o.x; band.y |= o.y; band.z |= o.z; band.w |= o.w;
                        }
                        int cursor = ymin, nbc = 0;
                        const int base = pass ? sh_band_base[tid] : 0;
                        // rows of the band are cut into windows of equal height (not kChunkTapRows, kChunkTapRows, ...,
                        // remainder): chunks of similar length hide each other's staging
                        int bf, bl;
                        mask_range_first_last(band, ymin, ymax, bf, bl);
                        const int bspan = bf < 0 ? 1 : bl - bf + 1;
                        const int nwin = (bspan + kChunkTapRows - 1) / kChunkTapRows;
                        const int win_rows = (bspan + nwin - 1) / nwin;
                        while (true) {
                            const int y0 = mask_first_from(band, cursor);
                            if (y0 < 0) break;
                            const int y1 = min(y0 + win_rows - 1, ymax);
                            if (pass) {
                                ChunkRec c;
                                int nseg = 0, lo = 1 << 20, hi = -(1 << 20), xlo = 1 << 20, xhi = -(1 << 20);
                                for (int g = g0; g < g1; ++g) {
                                    int f, l;
                                    mask_range_first_last(mask_load(&sh_occ[g * 4]), y0, y1, f, l);
                                    if (f < 0) continue;
                                    SegRec sg;
                                    const int d0 = xp0 + g * G + shear * (f - ymin) - centre, d1 = d0 + shear * (l - f);
                                    sg.dx0 = (int16_t)d0;
                                    sg.dy0 = (int16_t)(f - centre);
                                    sg.nsteps = (int16_t)(l - f + 1);
                                    sg.woff = 0;
                                    lo = min(lo, f - centre);
                                    hi = max(hi, l - centre);
                                    xlo = min(xlo, min(d0, d1));
                                    xhi = max(xhi, max(d0, d1) + G - 1);
                                    sh_segs[(base + nbc) * kChunkSegSlots + nseg] = sg;
                                    ++nseg;
                                }
                                c.dy_lo = (int16_t)lo; c.dy_hi = (int16_t)hi;
                                c.dx_lo = (int16_t)xlo; c.dx_hi = (int16_t)xhi;
                                c.wsteps = 0;
                                c.nseg = (uint8_t)nseg;
                                c.shear = (int8_t)shear;
                                c.group_w = (uint8_t)G;
                                c.pad = 0;
                                c.data_off16 = 0;
                                sh_chunks[base + nbc] = c;
                            }
                            ++nbc;
                            cursor = y1 + 1;
                        }
                        if (!pass) sh_band_count[tid] = nbc;
                    }
                    __syncthreads();
                    if (!pass) {
                        if (tid == 0) {
                            int total = 0;
                            for (int bnd = 0; bnd < nbands; ++bnd) {
                                sh_band_base[bnd] = total;
                                total += sh_band_count[bnd];
                            }
                            sh_nchunks = total <= kProgMaxChunks ? total : -1;
                        }
                        __syncthreads();
                        if (sh_nchunks < 0) break;                      // uniform: every thread reads the same word
                    }
                }
            } else if (tid == 0) {
                sh_nchunks = -1;
            }
            __syncthreads();
            DIB_CT(13);
            int nchunks = sh_nchunks;
            // offsets: weight vectors of a chunk's segments are laid out back to back
            if (nchunks > 0) {
                if (tid == 0) {
                    int data_off = (int)kProgHeaderBytes, total_steps = 0, nsegs = 0;
                    bool ok = true;
                    for (int ci = 0; ci < nchunks && ok; ++ci) {
                        int nw = 0;
                        for (int sgi = 0; sgi < sh_chunks[ci].nseg; ++sgi) {
                            sh_segs[ci * kChunkSegSlots + sgi].woff = (int16_t)nw;
                            nw += sh_segs[ci * kChunkSegSlots + sgi].nsteps;
                        }
                        total_steps += nw;
                        nsegs += sh_chunks[ci].nseg;
                        const int bytes = (kChunkSegBytes + 4 * G * (nw + 1) + 15) & ~15;
                        const ChunkRec& c = sh_chunks[ci];
                        if (4 * G * (nw + 1) > kChunkMaxWeightBytes + 16 || data_off + bytes > (int)kProgBytes ||
                            c.dx_hi - c.dx_lo > chunk_col_span(shear) || c.dy_hi - c.dy_lo >= kChunkTapRows) {
                            ok = false;
                            break;
                        }
                        sh_chunks[ci].wsteps = (int16_t)nw;
                        sh_chunks[ci].data_off16 = (uint16_t)(data_off >> 4);
                        data_off += bytes;
                    }
                    if (!ok) sh_nchunks = -1;
                    sh_total_steps = total_steps;
                    sh_nsegs = nsegs;
                }
                __syncthreads();
                nchunks = sh_nchunks;
            }
            if (nchunks > 0) {
                DIB_CT(14);
                // write chunk records, segment records and weight vectors
                for (int k = tid; k < nchunks; k += kCompactThreads) out_chunks[k] = sh_chunks[k];
                for (int ci = 0; ci < nchunks; ++ci) {
                    const ChunkRec c = sh_chunks[ci];
                    const size_t off = (size_t)c.data_off16 * 16;
                    SegRec* seg_out = reinterpret_cast<SegRec*>(my_prog + off);
                    float* wout = reinterpret_cast<float*>(my_prog + off + kChunkSegBytes);
                    if (tid < kChunkSegSlots) {
                        SegRec sg;
                        sg.dx0 = 0; sg.dy0 = 0; sg.nsteps = 0; sg.woff = 0;
                        if (tid < c.nseg) sg = sh_segs[ci * kChunkSegSlots + tid];
                        seg_out[tid] = sg;
                    }
                    if (tid < G) wout[c.wsteps * G + tid] = 0.0f;     // the vector the kernel prefetches past the end
                    for (int sgi = 0; sgi < c.nseg; ++sgi) {
                        const SegRec sg = sh_segs[ci * kChunkSegSlots + sgi];
                        for (int k = tid; k < sg.nsteps * G; k += kCompactThreads) {
                            const int step = k / G, e = k - step * G;
                            const int x = sg.dx0 + centre + shear * step + e, y = sg.dy0 + centre + step;
                            float w = 0.0f;
                            if (x >= 0 && x < side) {
                                const float v = psf_cell<T, kStaged>(psf, sh_psf, (int64_t)y * side + x);
                                w = (normalize && (v != 0.0f || !s_finite)) ? PsfNum<T>::normalized(v, s) : v;   // 0 / s is 0 for a finite nonzero sum
                            }
                            wout[(sg.woff + step) * G + e] = w;
                        }
                    }
                }
                prog_G = G;
                prog_k = shear;
                break;
            }
            if (attempt == 0 && cand == (1 * kNumShears + kShearMax)) break;      // the fallback candidate itself failed
        }
    }
    DIB_CT(7);
    int nchunks_final = want_prog ? sh_nchunks : -1;
    if (masked_ok) {           // masked program: group width 0 in the summary
        nchunks_final = mk_chunks;
        prog_G = 0;
        prog_k = 0;
        if (tid == 0) {
            sh_total_steps = mk_steps;
            sh_nsegs = mk_segs;
        }
    }
    if (nchunks_final <= 0) {
        flags |= DIB_META_NO_PROGRAM;
        nchunks_final = 0;
    }

    if (tid == 0) {
        dib_psf_meta m;
        m.count = count;
        m.ymin = (int16_t)(count ? ymin : 0);
        m.ymax = (int16_t)(count ? ymax : 0);
        m.xmin = (int16_t)(count ? xmin : 0);
        m.xmax = (int16_t)(count ? xmax : 0);
        m.sum = s;
        m.support = (int32_t)support;
        m.prog_chunks = nchunks_final;
        m.prog_steps = nchunks_final ? sh_total_steps : 0;
        m.flags = flags;
        m.prog_segs = nchunks_final ? sh_nsegs : 0;
        m.prog_group_w = (int16_t)(nchunks_final ? prog_G : 0);
        m.prog_shear = (int16_t)(nchunks_final ? prog_k : 0);
        m.sy = (double)sy;
        m.sx = (double)sx;
        m.syy = (double)syy;
        m.sxx = (double)sxx;
        m.sxy = (double)sxy;
        meta[n] = m;
    }
    DIB_CT(8);
}

}  // namespace dib

#ifdef DIB_COMPACT_TIMING
extern "C" __attribute__((visibility("default"))) int dib_debug_compact_times(unsigned long long* out) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, dib::g_compact_t, sizeof(unsigned long long) * 16);
}
#endif

extern "C" int dib_compact_taps(const void* psfs, int psf_dtype, int n_psfs, int side, int64_t psf_stride, int normalize,
                                void* tapset, int max_taps, void* stream) {
    using namespace dib;
    DIB_CHECK_ARG(psfs != nullptr && tapset != nullptr, "dib_compact_taps: null buffer");
    DIB_CHECK_ARG(n_psfs > 0, "dib_compact_taps: n_psfs must be > 0 (got %d)", n_psfs);
    DIB_CHECK_ARG(side >= 1 && side <= 512, "dib_compact_taps: PSF side %d outside [1, 512]", side);
    DIB_CHECK_ARG(psf_stride >= (int64_t)side * side, "dib_compact_taps: psf_stride %lld smaller than one PSF",
                  (long long)psf_stride);
    DIB_CHECK_ARG(max_taps > 0, "dib_compact_taps: max_taps must be > 0");
    DIB_CHECK_ARG((normalize & ~(DIB_COMPACT_NORMALIZE | DIB_COMPACT_DENSE_ONLY | DIB_COMPACT_MASKED_ONLY)) == 0, "dib_compact_taps: unknown flag bits 0x%x", normalize);
    DIB_CHECK_ARG(psf_dtype == DIB_F32 || psf_dtype == DIB_F16, "dib_compact_taps: PSF dtype must be DIB_F32 or DIB_F16");
    const dib_tapset_layout L = tapset_layout(n_psfs, max_taps);
    uint8_t* base = static_cast<uint8_t*>(tapset);
    dib_psf_meta* meta = reinterpret_cast<dib_psf_meta*>(base + L.meta_offset);
    dib_tap* taps = reinterpret_cast<dib_tap*>(base + L.taps_offset);
    uint8_t* prog = base + L.prog_offset;
    SchedWords* sched = reinterpret_cast<SchedWords*>(base + L.sched_offset);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool staged = side * side <= kStageMaxCells;
    const size_t smem = staged ? (size_t)side * side * sizeof(float) : 0;
    if (staged) {
        static thread_local int attr_dev = -1;      // opt in to > 48 KB of dynamic shared memory once per device
        int dev = 0;
        DIB_CUDA(cudaGetDevice(&dev));
        if (attr_dev != dev) {
            DIB_CUDA(cudaFuncSetAttribute(compact_taps_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kStageMaxCells * (int)sizeof(float)));
            DIB_CUDA(cudaFuncSetAttribute(compact_taps_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kStageMaxCells * (int)sizeof(float)));
            attr_dev = dev;
        }
    }
    if (psf_dtype == DIB_F32) {
        const float* p = static_cast<const float*>(psfs);
        if (staged)
            compact_taps_kernel<float, true><<<n_psfs, kCompactThreads, smem, st>>>(p, side, psf_stride, normalize, meta, taps,
                                                                                    max_taps, prog, sched);
        else
            compact_taps_kernel<float, false><<<n_psfs, kCompactThreads, 0, st>>>(p, side, psf_stride, normalize, meta, taps,
                                                                                  max_taps, prog, sched);
    } else {
        const __half* p = static_cast<const __half*>(psfs);
        if (staged)
            compact_taps_kernel<__half, true><<<n_psfs, kCompactThreads, smem, st>>>(p, side, psf_stride, normalize, meta, taps,
                                                                                     max_taps, prog, sched);
        else
            compact_taps_kernel<__half, false><<<n_psfs, kCompactThreads, 0, st>>>(p, side, psf_stride, normalize, meta, taps,
                                                                                   max_taps, prog, sched);
    }
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

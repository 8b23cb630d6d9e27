// GPU PSF rasterisation: seeded trajectories -> dense PSFs, bit-identical to the reference's fp64 Python loops.
//
// Replaces, per PSF (about 45 ms of Python on the host):
//   PSF.fit           motion_blur/generate_PSF.py:31-77   exposure-fraction time slicing (:47-56), 4-corner bilinear
//                                                          splat (:59-75), division by iters (:77)
//   PSF.centerPSF     motion_blur/generate_PSF.py:106-123 np.sum, weighted centroid, int() truncation, np.roll
//   crop [64:192]     transforms.py:334-335
//   astype(float16)   dataset_utils/generate_PSFs.py:60
// The trajectory itself (motion_blur/generate_trajectory.py:38-98) is a sequential random walk on numpy's MT19937
// stream and stays on the host: its samples are this kernel's input.
//
// Bit-exactness: a PSF cell's value is a chain of fp64 additions in ascending sample order.  The kernel gives
// every cell of the trajectory's bounding box to one thread, which walks the samples in order and adds exactly
// the contributions the Python loop adds (same expression, no FMA contraction), so parallelism is across cells
// only.  np.sum's pairwise tree (blocks of 128 with 8 strided partial sums, then a balanced binary tree) and the
// row-major centroid accumulation are reproduced as well, because the centring offset is an int() truncation.
//
// Cost: a cell is touched by a handful of short runs of consecutive samples (the walk moves ~0.05 cells per sample),
// so the samples are cut into blocks of 32 with a bounding box each and a thread only walks the blocks whose box
// contains its cell -- the additions it performs, and their order, are unchanged.  The scratch canvas is written
// inside the trajectory's bounding box only; everything outside is known to be zero and is never read or cleared.
#include <stdlib.h>

#include "dib_common.cuh"

namespace dib {

constexpr int kRasterThreads = 512;

__device__ __forceinline__ double tri(double v) { return fmax(0.0, __dsub_rn(1.0, fabs(v))); }

// generate_PSF.py:47-56 with a single fraction (prevT = 0)
__device__ __forceinline__ double time_weight(int t, double fraction, int iters) {
    const double fi = __dmul_rn(fraction, (double)iters);
    const double prev = 0.0;
    if (fi >= (double)t && prev < (double)(t - 1)) return 1.0;
    if (fi >= (double)(t - 1) && prev < (double)(t - 1)) return __dsub_rn(fi, (double)(t - 1));
    if (fi >= (double)t && prev < (double)t) return __dsub_rn((double)t, prev);
    if (fi >= (double)(t - 1) && prev < (double)t) return __dmul_rn(__dsub_rn(fraction, 0.0), (double)iters);
    return 0.0;
}

template <typename T>
__device__ __forceinline__ T cast_out(double v);
template <>
__device__ __forceinline__ double cast_out<double>(double v) { return v; }
template <>
__device__ __forceinline__ float cast_out<float>(double v) { return (float)v; }
template <>
__device__ __forceinline__ __half cast_out<__half>(double v) { return __double2half(v); }

// cluster-wide barrier with release/acquire ordering (the CTAs of a PSF exchange canvas cells through global memory)
__device__ __forceinline__ void cluster_sync_all() {
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// sequential fp64 chain acc + v[0] + v[1] + ... in order (one thread; the loads do not depend on the chain)
__device__ __forceinline__ double chain_add(double acc, const double* __restrict__ v, int count) {
    int i = 0;
    for (; i + 8 <= count; i += 8) {
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = v[i + j];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = __dadd_rn(acc, x[j]);
    }
    for (; i < count; ++i) acc = __dadd_rn(acc, v[i]);
    return acc;
}

constexpr int kListCap = 1024;      // centroid terms buffered between two runs of the sequential chain

// `split` CTAs (one thread-block cluster) work on one PSF: every CTA repeats the cheap per-sample preparation, the cells
// of the bounding box and the output pixels are divided among them, and each CTA repeats the canvas sum and the centroid
// (identical results) so that one cluster barrier is the only exchange.
template <typename T>
__global__ void __launch_bounds__(kRasterThreads)
rasterize_psf_kernel(const double* __restrict__ traj, const double* __restrict__ fractions, int iters, int canvas, int lg_canvas,
                     int center, int out_side, int split, T* __restrict__ out, int32_t* __restrict__ offsets,
                     double* __restrict__ scratch) {
    extern __shared__ __align__(16) uint8_t raster_smem[];
    double* s_re = reinterpret_cast<double*>(raster_smem);
    double* s_im = s_re + iters;
    double* s_w = s_im + iters;
    short2* s_m = reinterpret_cast<short2*>(s_w + iters);   // (m1 = row, m2 = col) per sample
    __shared__ int s_red[4][kRasterThreads / 32];
    __shared__ int s_box[4];
    __shared__ int s_tlast;
    __shared__ double s_tree[512];
    __shared__ int s_off[2];
    __shared__ int s_wcnt[kRasterThreads / 32];
    __shared__ short4 s_bb[4096 / 32];       // per block of 32 samples: rows [x, y], cols [z, w] its weighted samples touch
    __shared__ double s_px[kListCap], s_py[kListCap];

    const int n = blockIdx.x / split, rank = blockIdx.x % split;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kRasterThreads / 32;
    const double fraction = fractions[n];
    const double* tr = traj + (size_t)n * iters * 2;
    double* canvas_buf = scratch + (size_t)n * canvas * canvas;
    const int cells = canvas * canvas;
    const int cmask = canvas - 1;

    // 1. samples, base cells and time weights
    int ymin = 1 << 20, ymaxn = 1 << 20, xmin = 1 << 20, xmaxn = 1 << 20, tlast = 0;
    for (int t = tid; t < iters; t += kRasterThreads) {
        const double2 s = reinterpret_cast<const double2*>(tr)[t];
        const double re = s.x, im = s.y;
        s_re[t] = re;
        s_im[t] = im;
        const double w = time_weight(t, fraction, iters);
        s_w[t] = w;
        // m = min(canvas - 1, max(1, floor(v)))   (generate_PSF.py:59-62)
        const int m2 = (int)fmin((double)(canvas - 1), fmax(1.0, floor(re)));
        const int m1 = (int)fmin((double)(canvas - 1), fmax(1.0, floor(im)));
        s_m[t] = make_short2((short)m1, (short)m2);
        if (w != 0.0) {
            ymin = min(ymin, m1);
            ymaxn = min(ymaxn, -(m1 + 1));
            xmin = min(xmin, m2);
            xmaxn = min(xmaxn, -(m2 + 1));
            tlast = max(tlast, t);
        }
    }
    int v4[4] = {ymin, ymaxn, xmin, xmaxn};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v4[k] = min(v4[k], __shfl_xor_sync(0xffffffffu, v4[k], o));
        if (lane == 0) s_red[k][warp] = v4[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tlast = max(tlast, __shfl_xor_sync(0xffffffffu, tlast, o));
    if (tid == 0) s_tlast = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&s_tlast, tlast);
    if (tid < 4) {
        int m = s_red[tid][0];
        for (int k = 1; k < kWarps; ++k) m = min(m, s_red[tid][k]);
        s_box[tid] = m;
    }
    __syncthreads();
    const int y0 = s_box[0], y1 = min(-s_box[1], canvas - 1), x0 = s_box[2], x1 = min(-s_box[3], canvas - 1);
    const int t_last = s_tlast;
    const bool empty = (y0 > y1) || (x0 > x1);
    // 1b. bounding box of every block of 32 samples (weighted samples only: the others add an exact 0.0)
    const int nblk = t_last / 32 + 1;
    for (int k = warp; k < nblk; k += kWarps) {
        const int t = 32 * k + lane;
        int rlo = 1 << 14, rhin = 1 << 14, clo = 1 << 14, chin = 1 << 14;
        if (t <= t_last && s_w[t] != 0.0) {
            const short2 m = s_m[t];
            rlo = m.x;
            rhin = -(m.x + 1);
            clo = m.y;
            chin = -(m.y + 1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rlo = min(rlo, __shfl_xor_sync(0xffffffffu, rlo, o));
            rhin = min(rhin, __shfl_xor_sync(0xffffffffu, rhin, o));
            clo = min(clo, __shfl_xor_sync(0xffffffffu, clo, o));
            chin = min(chin, __shfl_xor_sync(0xffffffffu, chin, o));
        }
        if (lane == 0) s_bb[k] = make_short4((short)rlo, (short)(-rhin), (short)clo, (short)(-chin));   // empty block: lo > hi
    }
    __syncthreads();

    // 2. one thread per cell of the bounding box; contributions added in ascending sample order.  A warp takes a compact
    //    patch of cells (a compact patch meets few sample blocks) and skips the blocks that miss the patch as a whole.
    //    The patch is 8 x 4 cells when that gives every warp of the cluster work, else smaller (idle lanes, more warps):
    //    a warp's time is set by the samples that touch ANY of its cells, and a jittering trajectory revisits the
    //    cells of a small box for hundreds of samples.
    const int bw = empty ? 0 : x1 - x0 + 1, bh = empty ? 0 : y1 - y0 + 1;
    {
        int lw = 3, lh = 2;                       // log2 of the patch width / height
        while (lw + lh > 2 && ((bw + (1 << lw) - 1) >> lw) * ((bh + (1 << lh) - 1) >> lh) < split * kWarps) {
            if (lw > lh) --lw; else --lh;
        }
        const int pw = (bw + (1 << lw) - 1) >> lw, ph = (bh + (1 << lh) - 1) >> lh;
        const int pxe = (1 << lw) - 1, pye = (1 << lh) - 1;
        const bool lane_on = lane < (1 << (lw + lh));
        for (int g = rank * kWarps + warp; g < pw * ph; g += split * kWarps) {
            const int py0 = y0 + ((g / pw) << lh), px0 = x0 + ((g % pw) << lw);
            const int cy = lane_on ? py0 + (lane >> lw) : -8, cx = px0 + (lane & pxe);
            double acc = 0.0;
            for (int k = 0; k < nblk; ++k) {
                const short4 bb = s_bb[k];
                if (py0 + pye < bb.x || py0 > bb.y || px0 + pxe < bb.z || px0 > bb.w) continue;      // warp-uniform
                const int t1 = min(32 * k + 31, t_last);
                for (int t = 32 * k; t <= t1; ++t) {
                    const short2 m = s_m[t];
                    const int dy = cy - m.x, dx = cx - m.y;
                    if ((unsigned)dy <= 1u && (unsigned)dx <= 1u) {
                        // t_proportion * triangle_fun_prod(re - col, im - row)   (generate_PSF.py:64-75)
                        const double v = __dmul_rn(s_w[t], __dmul_rn(tri(__dsub_rn(s_re[t], (double)cx)), tri(__dsub_rn(s_im[t], (double)cy))));
                        acc = __dadd_rn(acc, v);
                    }
                }
            }
            if (lane_on && cy <= y1 && cx <= x1) canvas_buf[cy * canvas + cx] = __ddiv_rn(acc, (double)iters);   // PSF / iters (:77)
        }
    }
    cluster_sync_all();

    int ox = 0, oy = 0;
    if (center && !empty) {
        // 3. totalSum = np.sum(psf): numpy's pairwise summation over the flattened canvas.  Blocks of 128 elements use
        //    8 strided partial sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); blocks combine in a balanced tree.
        const int nblocks = cells / 128;       // canvas is a power of two >= 16 (checked on the host)
        double total = 0.0;
        for (int base = 0; base < nblocks; base += 512) {
            const int nb = min(512, nblocks - base);
            for (int b = tid; b < nb; b += kRasterThreads) {
                const int e0 = (base + b) * 128;             // flattened index of the block's first element
                double sum = 0.0;                            // a block that misses the bounding box sums zeros
                if ((e0 >> lg_canvas) <= y1 && ((e0 + 127) >> lg_canvas) >= y0) {
                    double r[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = 0.0;
#pragma unroll 4
                    for (int i = 0; i < 128; i += 8) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int e = e0 + i + j, yy = e >> lg_canvas, xx = e & cmask;
                            const double a = (yy >= y0 && yy <= y1 && xx >= x0 && xx <= x1) ? __ldcg(canvas_buf + e) : 0.0;
                            r[j] = i == 0 ? a : __dadd_rn(r[j], a);
                        }
                    }
                    sum = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                    __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                }
                s_tree[b] = sum;
            }
            __syncthreads();
            for (int stride = 1; stride < nb; stride <<= 1) {
                for (int b = tid * 2 * stride; b + stride < nb; b += kRasterThreads * 2 * stride)
                    s_tree[b] = __dadd_rn(s_tree[b], s_tree[b + stride]);
                __syncthreads();
            }
            // canvases above 256 would need the partial results of successive 512-block groups combined pairwise too;
            // the host restricts canvas to <= 256 so there is exactly one group
            total = s_tree[0];
            __syncthreads();
        }
        // 4. weighted centroid over cells with psf > 0, row-major, sequential fp64 (generate_PSF.py:110-117).  All threads
        //    form the terms j * (psf / totalSum), i * (psf / totalSum) of the positive cells into an order-preserving list;
        //    thread 0 (x) and thread 32 (y) run the two addition chains over it.
        double axy = 0.0;
        int count = 0;
        const int nbox = bw * bh;
        for (int base = 0; base < nbox; base += kRasterThreads) {
            const int cidx = base + tid;
            double v = 0.0;
            int cy = 0, cx = 0;
            if (cidx < nbox) {
                cy = y0 + cidx / bw;
                cx = x0 + cidx % bw;
                v = __ldcg(canvas_buf + cy * canvas + cx);
            }
            const bool pos = v > 0.0;
            const unsigned ballot = __ballot_sync(0xffffffffu, pos);
            if (lane == 0) s_wcnt[warp] = __popc(ballot);
            __syncthreads();
            int before = 0, chunk = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const int c = s_wcnt[w];
                before += w < warp ? c : 0;
                chunk += c;
            }
            if (count + chunk > kListCap) {
                if (tid == 0) axy = chain_add(axy, s_px, count);
                if (tid == 32) axy = chain_add(axy, s_py, count);
                count = 0;
                __syncthreads();
            }
            if (pos) {
                const double weight = __ddiv_rn(v, total);
                const int at = count + before + __popc(ballot & ((1u << lane) - 1u));
                s_px[at] = __dmul_rn((double)cx, weight);
                s_py[at] = __dmul_rn((double)cy, weight);
            }
            count += chunk;
            __syncthreads();
        }
        if (tid == 0) s_off[0] = (int)__dsub_rn(chain_add(axy, s_px, count), (double)canvas / 2.0);   // int(): toward zero (:119-120)
        if (tid == 32) s_off[1] = (int)__dsub_rn(chain_add(axy, s_py, count), (double)canvas / 2.0);
        __syncthreads();
        ox = s_off[0];
        oy = s_off[1];
    }
    if (rank == 0 && tid == 0 && offsets != nullptr) {
        offsets[2 * n] = ox;
        offsets[2 * n + 1] = oy;
    }

    // 5. np.roll by (-offsetX, -offsetY), central crop, cast   (generate_PSF.py:122-123, transforms.py:334-335)
    const int crop0 = (canvas - out_side) / 2;
    T* o = out + (size_t)n * out_side * out_side;
    for (int i = rank * kRasterThreads + tid; i < out_side * out_side; i += split * kRasterThreads) {
        const int yy = crop0 + i / out_side, xx = crop0 + i % out_side;
        const int sy = (yy + oy) & cmask, sx = (xx + ox) & cmask;     // canvas is a power of two: & == floored modulo
        const bool inbox = !empty && sy >= y0 && sy <= y1 && sx >= x0 && sx <= x1;
        o[i] = cast_out<T>(inbox ? __ldcg(canvas_buf + sy * canvas + sx) : 0.0);
    }
}

}  // namespace dib

extern "C" int dib_rasterize_psf(const double* traj, const double* fractions, int n, int iters, int canvas, int center,
                                 int out_side, void* out, int out_dtype, int32_t* offsets, double* scratch, void* stream) {
    using namespace dib;
    DIB_CHECK_ARG(traj != nullptr && fractions != nullptr && out != nullptr && scratch != nullptr, "dib_rasterize_psf: null buffer");
    DIB_CHECK_ARG(n > 0, "dib_rasterize_psf: n must be > 0");
    DIB_CHECK_ARG(iters >= 2 && iters <= 4096, "dib_rasterize_psf: iters %d outside [2, 4096]", iters);
    DIB_CHECK_ARG(canvas >= 16 && canvas <= 256 && (canvas & (canvas - 1)) == 0,
                  "dib_rasterize_psf: canvas must be a power of two in [16, 256] (got %d)", canvas);
    DIB_CHECK_ARG(out_side > 0 && out_side <= canvas && ((canvas - out_side) % 2) == 0,
                  "dib_rasterize_psf: out_side %d must be <= canvas with an even margin", out_side);
    DIB_CHECK_ARG(out_dtype == DIB_F64 || out_dtype == DIB_F32 || out_dtype == DIB_F16, "dib_rasterize_psf: bad out_dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)iters * (3 * sizeof(double) + sizeof(short2));
    DIB_CHECK_ARG(smem <= 160 * 1024, "dib_rasterize_psf: %d samples do not fit in shared memory", iters);
    int lg = 0;
    while ((1 << lg) < canvas) ++lg;
    // CTAs per PSF (one cluster): few PSFs -> latency matters, spread each over 8 SMs; many -> the grid fills the GPU anyway
    int split = n <= 18 ? 8 : n <= 64 ? 4 : n <= 160 ? 2 : 1;
    if (const char* e = getenv("DIB_RASTER_SPLIT")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) split = v;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n * split));
    cfg.blockDim = dim3(kRasterThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)split;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (out_dtype == DIB_F64) {
        DIB_CUDA(cudaFuncSetAttribute(rasterize_psf_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DIB_CUDA(cudaLaunchKernelEx(&cfg, rasterize_psf_kernel<double>, traj, fractions, iters, canvas, lg, center, out_side, split,
                                    static_cast<double*>(out), offsets, scratch));
    } else if (out_dtype == DIB_F32) {
        DIB_CUDA(cudaFuncSetAttribute(rasterize_psf_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DIB_CUDA(cudaLaunchKernelEx(&cfg, rasterize_psf_kernel<float>, traj, fractions, iters, canvas, lg, center, out_side, split,
                                    static_cast<float*>(out), offsets, scratch));
    } else {
        DIB_CUDA(cudaFuncSetAttribute(rasterize_psf_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DIB_CUDA(cudaLaunchKernelEx(&cfg, rasterize_psf_kernel<__half>, traj, fractions, iters, canvas, lg, center, out_side, split,
                                    static_cast<__half*>(out), offsets, scratch));
    }
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

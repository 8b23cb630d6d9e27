// Exact-order sparse-PSF blur (DIB_ALGO_GENERIC): every boundary mode of the reference, bit-identical arithmetic.
//
// Restates manual_blur (models/blur_functions.py:17-69): for each nonzero tap in row-major order,
//   output += roll(pad(image), (y - c, x - c)) * psf[y, x]
// i.e. one correctly-rounded multiply and one correctly-rounded add per tap in the image dtype (fp32, or fp16
// emulated through fp32 exactly as torch's half kernels do), with the pad/roll/crop index map of
// dib::src_index (reflect / zeros / replicate, including torch.roll's wrap-around).  This kernel serves the
// cases the tiled kernel does not take (tiny images, the 256-px branch, taps on the PSF border, fp16 I/O,
// arbitrary channel counts) and is the on-device bit-exact cross-check of the tiled kernel.
#include "dib_common.cuh"

namespace dib {

struct GenericParams {
    dib_image img[DIB_MAX_BATCH];
    const dib_tap* taps;
    const dib_psf_meta* meta;
    int max_taps;
    int n_images;
    uint64_t philox_seed, philox_offset;
    uint32_t skip_mask;      // bit i set: image i is handled by another kernel
    uint32_t planned_mask;   // DIB_ALGO_DEVICE_PLAN: bit i set: a tiled kernel takes image i if its PSF has a program (decided here
                             // from the device-side summary, as the tiled kernels do)
};

template <typename T>
struct IoNum;
template <>
struct IoNum<float> {
    __device__ static float load(const void* p, int64_t i) { return static_cast<const float*>(p)[i]; }
    __device__ static void store(void* p, int64_t i, float v) { static_cast<float*>(p)[i] = v; }
    __device__ static float mul(float a, float w) { return __fmul_rn(a, w); }
    __device__ static float add(float a, float b) { return __fadd_rn(a, b); }
};
template <>
struct IoNum<__half> {
    __device__ static float load(const void* p, int64_t i) { return __half2float(static_cast<const __half*>(p)[i]); }
    __device__ static void store(void* p, int64_t i, float v) { static_cast<__half*>(p)[i] = __float2half_rn(v); }
    // torch half kernels: widen, operate in fp32, round to half after every operation
    __device__ static float mul(float a, float w) { return __half2float(__float2half_rn(__fmul_rn(a, w))); }
    __device__ static float add(float a, float b) { return __half2float(__float2half_rn(__fadd_rn(a, b))); }
};

constexpr int kGenericThreads = 256;
constexpr int kChanChunk = 4;

template <typename T>
__global__ void __launch_bounds__(kGenericThreads)
blur_generic_kernel(const __grid_constant__ GenericParams p) {
    const int n = blockIdx.y;
    if ((p.skip_mask >> n) & 1u) return;
    const dib_image& im = p.img[n];
    if (((p.planned_mask >> n) & 1u) && im.psf_index >= 0) {
        const int kind = psf_program_kind(p.meta[im.psf_index]);
        if (sizeof(T) == 2 ? kind == 1 : kind != 0) return;      // half images: only the masked kernel has a half path
    }
    int count = 0;
    const dib_tap* taps = nullptr;
    if (im.psf_index >= 0) {
        count = min(p.meta[im.psf_index].count, p.max_taps);
        taps = p.taps + (int64_t)im.psf_index * p.max_taps;
    }
    // grid-stride over the image's pixels: the grid is bounded (a few CTAs per SM and image), so a device-planned launch
    // whose images all belong to the tiled kernels retires ~2 k CTAs instead of one per 256 pixels
    const int64_t npix = (int64_t)im.H * im.W;
    for (int64_t pix = (int64_t)blockIdx.x * kGenericThreads + threadIdx.x; pix < npix; pix += (int64_t)gridDim.x * kGenericThreads) {
    const int i = (int)(pix / im.W), j = (int)(pix - (int64_t)i * im.W);
    for (int c0 = 0; c0 < im.C; c0 += kChanChunk) {
        const int nc = min(kChanChunk, im.C - c0);
        float acc[kChanChunk];
#pragma unroll
        for (int c = 0; c < kChanChunk; ++c) acc[c] = 0.0f;
        if (im.psf_index >= 0) {
            for (int t = 0; t < count; ++t) {
                const dib_tap tp = taps[t];
                const int sr = src_index(i, im.H, tp.y, im.pad_mode);
                const int sc = src_index(j, im.W, tp.x, im.pad_mode);
                const bool inside = (sr >= 0) && (sc >= 0);
                const int64_t off = (int64_t)(inside ? sr : 0) * im.src_row_pitch + (inside ? sc : 0);
#pragma unroll
                for (int c = 0; c < kChanChunk; ++c) {
                    if (c < nc) {
                        const float v = inside ? IoNum<T>::load(im.src, (int64_t)(c0 + c) * im.src_chan_pitch + off) : 0.0f;
                        acc[c] = IoNum<T>::add(acc[c], IoNum<T>::mul(v, tp.w));
                    }
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < kChanChunk; ++c)
                if (c < nc) acc[c] = IoNum<T>::load(im.src, (int64_t)(c0 + c) * im.src_chan_pitch + (int64_t)i * im.src_row_pitch + j);
        }
#pragma unroll
        for (int c = 0; c < kChanChunk; ++c) {
            if (c < nc) {
                const int ch = c0 + c;
                const int64_t o = (int64_t)ch * im.dst_chan_pitch + (int64_t)i * im.dst_row_pitch + j;
                float v = acc[c];
                if (im.epilogue) {
                    Epilogue e;
                    e.flags = im.epilogue;
                    e.noise_sd = im.noise_sd;
                    e.gamma = im.gamma;
                    e.mean = im.mean[ch & 3];
                    e.std = im.std[ch & 3];
                    float nz = 0.0f;
                    if (im.epilogue & DIB_EPI_NOISE) {
                        if (im.noise != nullptr)
                            nz = IoNum<T>::load(im.noise, o);
                        else
                            nz = philox_normal(p.philox_seed, p.philox_offset + (uint64_t)n,
                                               ((uint64_t)ch * im.H + i) * im.W + j);
                    }
                    if constexpr (sizeof(T) == 2) {
                        // half images: torch rounds to half after every elementwise operation (blur_functions.py:74:
                        // randn * sd, + output, clamp; net_transforms.py:135-139: - mean, / std with half mean / std tensors)
                        if (e.flags & DIB_EPI_NOISE) v = IoNum<T>::add(v, IoNum<T>::mul(nz, e.noise_sd));
                        if (e.flags & DIB_EPI_CLAMP) v = fminf(fmaxf(v, 0.0f), 1.0f);
                        if (e.flags & DIB_EPI_GAMMA) v = __half2float(__float2half_rn(powf(v, e.gamma)));
                        if (e.flags & DIB_EPI_NORMALIZE) {
                            const float mh = __half2float(__float2half_rn(e.mean)), sh = __half2float(__float2half_rn(e.std));
                            v = __half2float(__float2half_rn(__fsub_rn(v, mh)));
                            v = __half2float(__float2half_rn(__fdiv_rn(v, sh)));
                        }
                    } else {
                        v = apply_epilogue_f32(v, e, nz);
                    }
                }
                IoNum<T>::store(im.dst, o, v);
            }
        }
    }
    }
}

// Launch helper used by dib_blur_batch (blur_api.cu).
int launch_generic(const dib_image* images, int n_images, const dib_tap* taps, const dib_psf_meta* meta, int max_taps,
                   int io_dtype, uint32_t skip_mask, uint32_t planned_mask, uint64_t seed, uint64_t offset, cudaStream_t st) {
    GenericParams p;
    int64_t max_pix = 0;
    for (int k = 0; k < n_images; ++k) {
        p.img[k] = images[k];
        if (!((skip_mask >> k) & 1u)) {
            const int64_t px = (int64_t)images[k].H * images[k].W;
            if (px > max_pix) max_pix = px;
        }
    }
    if (max_pix == 0) return DIB_OK;
    p.taps = taps;
    p.meta = meta;
    p.max_taps = max_taps;
    p.n_images = n_images;
    p.philox_seed = seed;
    p.philox_offset = offset;
    p.skip_mask = skip_mask;
    p.planned_mask = planned_mask;
    static thread_local int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        DIB_CUDA(cudaGetDevice(&dev));
        DIB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    // two full waves of resident CTAs (8 of 256 threads per SM) shared by the images; never more than the pixels need
    const int64_t blocks_needed = (max_pix + kGenericThreads - 1) / kGenericThreads;
    const int64_t blocks_cap = ((int64_t)sm_count * 16 + n_images - 1) / n_images;
    dim3 grid((unsigned)(blocks_needed < blocks_cap ? blocks_needed : blocks_cap), (unsigned)n_images);
    if (io_dtype == DIB_F32)
        blur_generic_kernel<float><<<grid, kGenericThreads, 0, st>>>(p);
    else
        blur_generic_kernel<__half><<<grid, kGenericThreads, 0, st>>>(p);
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

}  // namespace dib

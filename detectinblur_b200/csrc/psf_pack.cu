// Device-side reader of the packed sparse PSF bank (detectinblur_b200/psf_bank.py): one 32-bit word per stored tap,
// y | x << 8 | fp16 bits << 16 on the bank's 256 x 256 canvas, expanded into the dense crop the reference's reader
// produces (transforms.py:301-309: np.load of float16[256,256], then [64:192, 64:192]).
#include "dib_common.cuh"

namespace dib {

template <typename T>
__global__ void __launch_bounds__(256) unpack_psfs_kernel(const uint32_t* __restrict__ taps, const int64_t* __restrict__ offsets,
                                                          int crop_lo, int side, T* __restrict__ out) {
    const int n = blockIdx.x;
    T* dst = out + (size_t)n * side * side;
    for (int k = threadIdx.x; k < side * side; k += blockDim.x) dst[k] = T(0);
    __syncthreads();
    const int64_t t0 = offsets[n], t1 = offsets[n + 1];
    for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const uint32_t v = taps[t];
        const int y = (int)(v & 0xffu) - crop_lo, x = (int)((v >> 8) & 0xffu) - crop_lo;
        if (y < 0 || y >= side || x < 0 || x >= side) continue;      // outside the crop, exactly as the dense reader drops it
        const __half h = __ushort_as_half((unsigned short)(v >> 16));
        if constexpr (sizeof(T) == 2)
            dst[y * side + x] = h;
        else
            dst[y * side + x] = (T)__half2float(h);                  // fp16 -> fp32 / fp64 is exact
    }
}

}  // namespace dib

extern "C" int dib_unpack_psfs(const uint32_t* taps, const int64_t* offsets, int n, int crop_lo, int out_side, void* out,
                               int out_dtype, void* stream) {
    using namespace dib;
    DIB_CHECK_ARG(n > 0, "dib_unpack_psfs: n must be positive (got %d)", n);
    DIB_CHECK_ARG(taps != nullptr && offsets != nullptr && out != nullptr, "dib_unpack_psfs: NULL buffer");
    DIB_CHECK_ARG(out_side > 0 && out_side <= 256 && crop_lo >= 0 && crop_lo + out_side <= 256,
                  "dib_unpack_psfs: crop [%d, %d) leaves the 256-cell canvas", crop_lo, crop_lo + out_side);
    DIB_CHECK_ARG(out_dtype == DIB_F32 || out_dtype == DIB_F16 || out_dtype == DIB_F64, "dib_unpack_psfs: bad out_dtype %d", out_dtype);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (out_dtype == DIB_F16)
        unpack_psfs_kernel<__half><<<n, 256, 0, st>>>(taps, offsets, crop_lo, out_side, static_cast<__half*>(out));
    else if (out_dtype == DIB_F32)
        unpack_psfs_kernel<float><<<n, 256, 0, st>>>(taps, offsets, crop_lo, out_side, static_cast<float*>(out));
    else
        unpack_psfs_kernel<double><<<n, 256, 0, st>>>(taps, offsets, crop_lo, out_side, static_cast<double*>(out));
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

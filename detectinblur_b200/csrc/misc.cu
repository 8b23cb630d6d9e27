// Checksum and FP32-pipe probe entry points of the C ABI (include/dib.h).
#include "dib_common.cuh"

namespace dib {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Sum over elements of mix64(index, bits) modulo 2^64: independent of the order in which threads / ranks add, so
// per-shard values can be compared or combined after an all-gather (utils.py:536-576 is the reference's pattern).
template <typename BitsT>
__global__ void checksum_kernel(const BitsT* __restrict__ data, int64_t n, unsigned long long* __restrict__ out) {
    uint64_t acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += mix64(((uint64_t)i << 32) ^ ((uint64_t)i >> 32) ^ ((uint64_t)data[i] * 0x9E3779B97F4A7C15ull) ^ (uint64_t)i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ unsigned long long sh[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        atomicAdd(out, t);
    }
}

__global__ void zero_u64_kernel(unsigned long long* p) { *p = 0ull; }

constexpr int kProbeChains = 16;
constexpr int kProbeInner = 64;
__global__ void __launch_bounds__(256) fp32_probe_kernel(int iters, float a, float b, float* __restrict__ sink) {
    float acc[kProbeChains];
#pragma unroll
    for (int k = 0; k < kProbeChains; ++k) acc[k] = (float)(threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < kProbeInner; ++j) {
#pragma unroll
            for (int k = 0; k < kProbeChains; ++k) acc[k] = fmaf(acc[k], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kProbeChains; ++k) s += acc[k];
    if (s == 12345.678f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true in practice; keeps the chain alive
}

// uint8 <-> float image planes, four pixels per thread.  Rows may be pitched on the float side (the blur's results are
// [:, :, :W] views of row-aligned buffers); the uint8 side is dense.
// The 256 possible results of byte / 255 (IEEE division in fp32, what torchvision's to_tensor computes) are tabulated once
// per CTA: a lookup per pixel instead of a division.
template <typename F>
__global__ void __launch_bounds__(256) u8_to_float_kernel(const uint8_t* __restrict__ src, F* __restrict__ dst, int64_t rows, int W,
                                                          int64_t dst_pitch) {
    __shared__ float lut[256];
    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
    __syncthreads();
    if (dst_pitch == W && (reinterpret_cast<uintptr_t>(src) & 3u) == 0 && (reinterpret_cast<uintptr_t>(dst) & (4 * sizeof(F) - 1)) == 0) {
        // dense on both sides: the planes are one flat array -- four bytes in, four values out per thread, all vector accesses
        const int64_t total = rows * W, nq = total >> 2;
        for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
            const uchar4 b = reinterpret_cast<const uchar4*>(src)[q];
            const float4 v = make_float4(lut[b.x], lut[b.y], lut[b.z], lut[b.w]);
            if constexpr (sizeof(F) == 4) {
                reinterpret_cast<float4*>(dst)[q] = v;
            } else {
                __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
                uint2 o;
                o.x = *reinterpret_cast<unsigned*>(&lo);
                o.y = *reinterpret_cast<unsigned*>(&hi);
                reinterpret_cast<uint2*>(dst)[q] = o;
            }
        }
        if (blockIdx.x == 0 && threadIdx.x < (int)(total & 3)) dst[(nq << 2) + threadIdx.x] = (F)lut[src[(nq << 2) + threadIdx.x]];
        return;
    }
    const int quads = (W + 3) >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * quads; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / quads;
        const int x0 = (int)(i - r * quads) * 4;
        const uint8_t* s = src + r * W + x0;
        F* d = dst + r * dst_pitch + x0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (x0 + j < W) d[j] = (F)lut[s[j]];
    }
}
template <typename F>
__global__ void float_to_u8_kernel(const F* __restrict__ src, uint8_t* __restrict__ dst, int64_t rows, int W, int64_t src_pitch) {
    const int quads = (W + 3) >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * quads; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / quads;
        const int x0 = (int)(i - r * quads) * 4;
        const F* s = src + r * src_pitch + x0;
        uint8_t* d = dst + r * W + x0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (x0 + j < W) {
                const float v = fminf(fmaxf(__fmul_rn((float)s[j], 255.0f), 0.0f), 255.0f);
                d[j] = (uint8_t)v;                                          // truncation, as numpy's astype(uint8)
            }
    }
}

}  // namespace dib

extern "C" int dib_u8_to_float(const uint8_t* src, void* dst, int dst_dtype, int64_t rows, int W, int64_t dst_row_pitch, void* stream) {
    using namespace dib;
    DIB_CHECK_ARG(src != nullptr && dst != nullptr && rows >= 0 && W > 0 && dst_row_pitch >= W, "dib_u8_to_float: bad arguments");
    DIB_CHECK_ARG(dst_dtype == DIB_F32 || dst_dtype == DIB_F16, "dib_u8_to_float: destination must be float32 or float16");
    if (rows == 0) return DIB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t work = rows * ((W + 3) / 4);
    const int blocks = (int)(work / 256 + 1 < 148 * 16 ? work / 256 + 1 : 148 * 16);
    if (dst_dtype == DIB_F32)
        u8_to_float_kernel<float><<<blocks, 256, 0, st>>>(src, static_cast<float*>(dst), rows, W, dst_row_pitch);
    else
        u8_to_float_kernel<__half><<<blocks, 256, 0, st>>>(src, static_cast<__half*>(dst), rows, W, dst_row_pitch);
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

extern "C" int dib_float_to_u8(const void* src, int src_dtype, uint8_t* dst, int64_t rows, int W, int64_t src_row_pitch, void* stream) {
    using namespace dib;
    DIB_CHECK_ARG(src != nullptr && dst != nullptr && rows >= 0 && W > 0 && src_row_pitch >= W, "dib_float_to_u8: bad arguments");
    DIB_CHECK_ARG(src_dtype == DIB_F32 || src_dtype == DIB_F16, "dib_float_to_u8: source must be float32 or float16");
    if (rows == 0) return DIB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t work = rows * ((W + 3) / 4);
    const int blocks = (int)(work / 256 + 1 < 148 * 16 ? work / 256 + 1 : 148 * 16);
    if (src_dtype == DIB_F32)
        float_to_u8_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(src), dst, rows, W, src_row_pitch);
    else
        float_to_u8_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(src), dst, rows, W, src_row_pitch);
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

extern "C" int dib_checksum(const void* data, int dtype, int64_t n_elements, uint64_t* out, int accumulate, void* stream) {
    using namespace dib;
    DIB_CHECK_ARG(out != nullptr, "dib_checksum: out is NULL");
    DIB_CHECK_ARG(n_elements >= 0 && (data != nullptr || n_elements == 0), "dib_checksum: bad buffer");
    DIB_CHECK_ARG(dtype == DIB_F32 || dtype == DIB_F16 || dtype == DIB_F64, "dib_checksum: bad dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* o = reinterpret_cast<unsigned long long*>(out);
    if (!accumulate) zero_u64_kernel<<<1, 1, 0, st>>>(o);
    if (n_elements > 0) {
        int sms = 148;
        dib_device_info(&sms, nullptr);
        const int threads = 256;
        int64_t want = (n_elements + threads - 1) / threads;
        const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
        if (dtype == DIB_F32)
            checksum_kernel<uint32_t><<<grid, threads, 0, st>>>(static_cast<const uint32_t*>(data), n_elements, o);
        else if (dtype == DIB_F16)
            checksum_kernel<uint16_t><<<grid, threads, 0, st>>>(static_cast<const uint16_t*>(data), n_elements, o);
        else
            checksum_kernel<uint64_t><<<grid, threads, 0, st>>>(static_cast<const uint64_t*>(data), n_elements, o);
    }
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

extern "C" int dib_fp32_probe(int iters, float* sink, uint64_t* fma_count, void* stream) {
    using namespace dib;
    DIB_CHECK_ARG(iters > 0 && sink != nullptr, "dib_fp32_probe: iters must be > 0 and sink non-NULL");
    int sms = 148;
    const int rc = dib_device_info(&sms, nullptr);
    if (rc != DIB_OK) return rc;
    const int blocks = sms * 8, threads = 256;     // sink must hold blocks * threads floats
    fp32_probe_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(iters, 0.999f, 0.001f, sink);
    DIB_CUDA(cudaGetLastError());
    if (fma_count) *fma_count = (uint64_t)blocks * threads * (uint64_t)iters * kProbeChains * kProbeInner;
    return DIB_OK;
}

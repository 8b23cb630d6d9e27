// Throughput probe: FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, with and without register-only operands.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) probe(int iters, float a, float b, float* sink, const float* src) {
    constexpr int N = 16;
    if (MODE == 0) {          // FFMA, 3 register operands (weight in a register, window value in a register)
        float acc[N], x[N];
        for (int k = 0; k < N; ++k) { acc[k] = threadIdx.x + k; x[k] = src[threadIdx.x + k]; }
        float w = src[blockIdx.x];
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
#pragma unroll
                for (int k = 0; k < N; ++k) acc[k] = fmaf(w, x[(k + j) % N], acc[k]);
        }
        float s = 0; for (int k = 0; k < N; ++k) s += acc[k];
        if (s == 12345.678f) sink[threadIdx.x] = s;
    } else {                  // FFMA2: pairs of accumulators and window values, scalar weight
        unsigned long long acc[N / 2], x[N / 2];
        for (int k = 0; k < N / 2; ++k) {
            float2 t = make_float2(threadIdx.x + k, threadIdx.x - k), u = make_float2(src[threadIdx.x + k], src[threadIdx.x + 2 * k + 1]);
            acc[k] = *reinterpret_cast<unsigned long long*>(&t); x[k] = *reinterpret_cast<unsigned long long*>(&u);
        }
        float wf = src[blockIdx.x]; float2 w2 = make_float2(wf, wf);
        unsigned long long w = *reinterpret_cast<unsigned long long*>(&w2);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
#pragma unroll
                for (int k = 0; k < N / 2; ++k) acc[k] = ffma2(w, x[(k + j) % (N / 2)], acc[k]);
        }
        float s = 0; for (int k = 0; k < N / 2; ++k) { float2 t = *reinterpret_cast<float2*>(&acc[k]); s += t.x + t.y; }
        if (s == 12345.678f) sink[threadIdx.x] = s;
    }
}
int main() {
    float *sink, *src; cudaMalloc(&sink, 1 << 20); cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int warps_per_sm : {8, 16, 32, 64}) {
            int blocks = sms * warps_per_sm / 8, iters = 4000;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) probe<0><<<blocks, 256>>>(iters, 0.999f, 0.001f, sink, src);
                else probe<1><<<blocks, 256>>>(iters, 0.999f, 0.001f, sink, src);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fmas = (double)blocks * 256 * iters * 32 * 16;     // scalar FMAs in both modes
            printf("mode %s warps/SM %2d: %.2f TFLOP/s (%.3f ms) err=%s\n", mode ? "FFMA2" : "FFMA ", warps_per_sm, 2 * fmas / ms / 1e9, ms, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}

// Does a predicated-off FFMA2 occupy the FMA pipe?  Mix of real and predicated-off FFMA2 in a dependency-free stream.
#include <cstdio>
#include <cuda_runtime.h>
template <int OFF_PER_ON>
__global__ void __launch_bounds__(256) probe(int iters, float* sink, const float* src, int flag) {
    constexpr int N = 8;
    unsigned long long acc[N], x[N];
    for (int k = 0; k < N; ++k) {
        float2 t = make_float2(threadIdx.x + k, threadIdx.x - k), u = make_float2(src[threadIdx.x + k], src[threadIdx.x + 2 * k + 1]);
        acc[k] = *reinterpret_cast<unsigned long long*>(&t); x[k] = *reinterpret_cast<unsigned long long*>(&u);
    }
    float wf = src[blockIdx.x]; float woff = src[blockIdx.x + 1] + (float)flag;   // woff == 0 at run time -> predicate false
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            asm volatile("{ .reg .pred p; .reg .b64 w; setp.neu.f32 p, %8, %8; mov.b64 w, {%9, %9};\n"
                         "fma.rn.f32x2 %0, w, %4, %0; fma.rn.f32x2 %1, w, %5, %1; fma.rn.f32x2 %2, w, %6, %2; fma.rn.f32x2 %3, w, %7, %3; }"
                         : "+l"(acc[0]), "+l"(acc[1]), "+l"(acc[2]), "+l"(acc[3]) : "l"(x[0]), "l"(x[1]), "l"(x[2]), "l"(x[3]), "f"(woff), "f"(wf));
            asm volatile("{ .reg .b64 w; mov.b64 w, {%8, %8};\n"
                         "fma.rn.f32x2 %0, w, %4, %0; fma.rn.f32x2 %1, w, %5, %1; fma.rn.f32x2 %2, w, %6, %2; fma.rn.f32x2 %3, w, %7, %3; }"
                         : "+l"(acc[4]), "+l"(acc[5]), "+l"(acc[6]), "+l"(acc[7]) : "l"(x[4]), "l"(x[5]), "l"(x[6]), "l"(x[7]), "f"(wf));
#pragma unroll
            for (int o = 0; o < OFF_PER_ON; ++o) {
                asm volatile("{ .reg .pred p; .reg .b64 w; setp.neu.f32 p, %8, 0f00000000; mov.b64 w, {%9, %9};\n"
                             "@p fma.rn.f32x2 %0, w, %4, %0; @p fma.rn.f32x2 %1, w, %5, %1; @p fma.rn.f32x2 %2, w, %6, %2; @p fma.rn.f32x2 %3, w, %7, %3;\n"
                             "@p fma.rn.f32x2 %0, w, %5, %0; @p fma.rn.f32x2 %1, w, %6, %1; @p fma.rn.f32x2 %2, w, %7, %2; @p fma.rn.f32x2 %3, w, %4, %3; }"
                             : "+l"(acc[0]), "+l"(acc[1]), "+l"(acc[2]), "+l"(acc[3]) : "l"(x[4]), "l"(x[5]), "l"(x[6]), "l"(x[7]), "f"(woff), "f"(wf));
            }
        }
    }
    float s = 0; for (int k = 0; k < N; ++k) { float2 t = *reinterpret_cast<float2*>(&acc[k]); s += t.x + t.y; }
    if (s == 12345.678f) sink[threadIdx.x] = s;
}
int main() {
    float *sink, *src; cudaMalloc(&sink, 1 << 20); cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        int blocks = sms * 12 / 8 * 1, iters = 4000;     // 12 warps per SM
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<blocks, 256>>>(iters, sink, src, 0);
            if (mode == 1) probe<1><<<blocks, 256>>>(iters, sink, src, 0);
            if (mode == 2) probe<2><<<blocks, 256>>>(iters, sink, src, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double real = (double)blocks * 256 * iters * 16 * 8;     // real FFMA2 per launch
        printf("real FFMA2 : predicated-off FFMA2 = 8 : %2d -> %.3f ms, %.2f TFLOP/s of real math (%s)\n", 8 * mode, ms, 4 * real / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

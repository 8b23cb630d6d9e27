// GPU camera-shake trajectories: the Boracchi-Foi random walk of motion_blur/generate_trajectory.py:38-98 with a
// counter-based generator, one thread per trajectory.
//
// The reference draws from numpy's global MT19937 stream, conditionally (an impulsive shake consumes one extra
// uniform), so its samples cannot be reproduced bit for bit in parallel; this kernel is the statistical equivalent
// SURVEY.md section 8f row 4 asks for: same update equations in fp64, same distributions, but every step t of
// trajectory k takes its randomness from Philox4x32-10(seed; counter = t + 1, stream = first_index + k): word 0 decides
// the impulsive shake, word 1 is its angle jitter, words 2-3 make the two Gaussians (Box-Muller); counter 0 holds the
// four shape parameters.  Trajectories are therefore reproducible from (seed, index) alone, in any batch split.
// oracle/psf_oracle.py:trajectory_philox restates exactly this; tests compare the two and check the walk's statistics
// against the host (MT19937) mirror.
//
//   4 uniform draws (centripetal .7U, big-shake .2U, gaussian 10U, angle 360U)     generate_trajectory.py:48-53
//   v = v0 * expl (expl > 0)                                                       :61-62
//   per step: U < p * expl -> impulsive term 2 v exp(i (pi + U - .5))              :69-73
//             dv = kick + expl * (gauss * (N + iN) - centripetal * x) * step       :75-77
//             v = (v + dv) / |v + dv| * max_len / (iters - 1);  x[t+1] = x[t] + v  :79-82
//   x += canvas / 2 * (1 + 1j)                                                     :92
#include "dib_common.cuh"

namespace dib {

__device__ __forceinline__ double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

__global__ void __launch_bounds__(64) trajectory_kernel(uint64_t seed, uint64_t first_index, const uint64_t* __restrict__ indices,
                                                        int n, int iters, double max_len, double canvas,
                                                        const double* __restrict__ expl, double* __restrict__ out,
                                                        int32_t* __restrict__ big_count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const Philox ph(seed);
    const uint64_t stream = indices != nullptr ? indices[k] : first_index + (uint64_t)k;
    const double e = expl[k];
    const uint4 r0 = ph(0, stream);
    const double centripetal = 0.7 * u01(r0.x);
    const double prob_big_shake = 0.2 * u01(r0.y);
    const double gaussian_shake = 10.0 * u01(r0.z);
    const double ang = 360.0 * u01(r0.w) * (3.14159265358979323846 / 180.0);
    const double step = max_len / (double)(iters - 1);
    double vx, vy;
    sincos(ang, &vy, &vx);
    const double v_scale = e > 0.0 ? e : step;
    vx *= v_scale;
    vy *= v_scale;
    double x = 0.0, y = 0.0;
    const double half = canvas / 2.0;
    double* o = out + (size_t)k * iters * 2;
    o[0] = half;
    o[1] = half;
    const double threshold = prob_big_shake * e;
    int big = 0;
    for (int t = 0; t < iters - 1; ++t) {
        const uint4 r = ph((uint64_t)t + 1, stream);
        double kx = 0.0, ky = 0.0;
        if (u01(r.x) < threshold) {
            double s, c;
            sincos(3.14159265358979323846 + (u01(r.y) - 0.5), &s, &c);
            kx = 2.0 * (vx * c - vy * s);
            ky = 2.0 * (vx * s + vy * c);
            ++big;
        }
        const double rad = sqrt(-2.0 * log(u01(r.z)));
        double gs, gc;
        sincos(6.28318530717958647692 * u01(r.w), &gs, &gc);
        const double g0 = rad * gc, g1 = rad * gs;
        vx += kx + e * (gaussian_shake * g0 - centripetal * x) * step;
        vy += ky + e * (gaussian_shake * g1 - centripetal * y) * step;
        const double inv = step / sqrt(vx * vx + vy * vy);
        vx *= inv;
        vy *= inv;
        x += vx;
        y += vy;
        o[2 * (t + 1)] = x + half;
        o[2 * (t + 1) + 1] = y + half;
    }
    if (big_count != nullptr) big_count[k] = big;
}

}  // namespace dib

extern "C" int dib_generate_trajectories(uint64_t seed, uint64_t first_index, const uint64_t* indices, int n, int iters,
                                         double max_len, double canvas, const double* expl, double* out, int32_t* big_count,
                                         void* stream) {
    using namespace dib;
    DIB_CHECK_ARG(n > 0, "dib_generate_trajectories: n must be > 0 (got %d)", n);
    DIB_CHECK_ARG(iters >= 2, "dib_generate_trajectories: iters must be >= 2 (got %d)", iters);
    DIB_CHECK_ARG(expl != nullptr && out != nullptr, "dib_generate_trajectories: NULL buffer");
    DIB_CHECK_ARG(max_len > 0.0 && canvas > 0.0, "dib_generate_trajectories: max_len and canvas must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    trajectory_kernel<<<(n + 63) / 64, 64, 0, st>>>(seed, first_index, indices, n, iters, max_len, canvas, expl, out, big_count);
    DIB_CUDA(cudaGetLastError());
    return DIB_OK;
}

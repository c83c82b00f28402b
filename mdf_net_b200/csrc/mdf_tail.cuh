// mdf_tail.cuh -- per-column pieces shared by the head kernels (mdf_head.cu: from logits; mdf_prob_head.cu: from the
// regulariser's last feature volume): the confidence window (net/unit/regress.py:9-25), the nearest-neighbour
// upsampled store (net/core.py:75-77) and the curve fits of HyposByFit (net/unit/depthhypos.py:78-125, 169-215).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mdf {

// confidence of one pixel from its probability column (regress.py:13-18):
//   S[k] = n * avg_pool(pad_D(prob))[k] = n * ((sum_{j<n} prob[k - pad_front + j]) / n)
template <class ProbAt>
__device__ __forceinline__ float window_confidence(ProbAt prob_at, float expect_idx, int D, int n, int pad_front, int pad_back)
{
    const int Dp = D + pad_front + pad_back - n + 1;
    int k = (int)expect_idx;                       // .long() truncates toward zero
    k = max(0, min(k, Dp - 1));                    // torch.gather would raise; cannot happen for a softmax output
    float s = 0.0f;
    for (int j = 0; j < n; ++j) {
        const int d = k - pad_front + j;
        s = __fadd_rn(s, (d >= 0 && d < D) ? prob_at(d) : 0.0f);
    }
    const float fn = (float)n;
    return __fmul_rn(fn, __fdiv_rn(s, fn));
}

__device__ __forceinline__ void store_upsampled(float* __restrict__ conf, float c, int b, int y, int x, int H, int W, int up)
{
    const size_t Wu = (size_t)W * up;
    float* base = conf + ((size_t)b * H * up + (size_t)y * up) * Wu + (size_t)x * up;
    if (up == 2 && (reinterpret_cast<uintptr_t>(base) & 7) == 0) {
        const float2 v = make_float2(c, c);
        *reinterpret_cast<float2*>(base) = v;           // x*2 floats: 8-byte aligned
        *reinterpret_cast<float2*>(base + Wu) = v;
    } else {
        for (int uy = 0; uy < up; ++uy)
            for (int ux = 0; ux < up; ++ux) base[(size_t)uy * Wu + ux] = c;
    }
}

// ------------------------------------------------------------------------------------------------
// HyposByFit's per-pixel curve fit (net/unit/depthhypos.py:78-125 "laplace", :169-215 "gauss1") on the column the
// head has in its hands anyway, so that the probability volume is not read a second time (and need not be
// written at all when nobody else wants it).  Same arithmetic as hypos_fit_kernel (mdf_hypos.cu); FIT: 0 none,
// 1 gauss1, 2 laplace.  The partial sums of the DS depth slices of a pixel are combined with shfl.xor.
// ------------------------------------------------------------------------------------------------
struct GaussMoments {
    double m1 = 0, m2 = 0, m3 = 0, m4 = 0, r0 = 0, r1 = 0, r2 = 0;
    __device__ __forceinline__ void add(double u, float prob)
    {
        const double z = log((double)fmaxf(prob, 1e-40f));
        const double u2 = u * u;
        m1 += u; m2 += u2; m3 += u2 * u; m4 += u2 * u2;
        r0 += z; r1 += u * z; r2 += u2 * z;
    }
    template <int PW>
    __device__ __forceinline__ void reduce_slices()
    {
#pragma unroll
        for (int o = PW; o < 32; o <<= 1) {
            m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o);
            m3 += __shfl_xor_sync(0xffffffffu, m3, o); m4 += __shfl_xor_sync(0xffffffffu, m4, o);
            r0 += __shfl_xor_sync(0xffffffffu, r0, o); r1 += __shfl_xor_sync(0xffffffffu, r1, o);
            r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        }
    }
    // [[m4 m3 m2][m3 m2 m1][m2 m1 m0]] [c2 c1 c0]^T = [r2 r1 r0]^T, s = |-1 / c2| by Cramer's rule
    __device__ __forceinline__ float scale(int D) const
    {
        const double m0 = (double)D;
        const double k1 = m2 * m0 - m1 * m1, k2 = m3 * m0 - m1 * m2, k3 = m3 * m1 - m2 * m2;
        const double det = m4 * k1 - m3 * k2 + m2 * k3;
        const double num = r2 * k1 - m3 * (r1 * m0 - m1 * r0) + m2 * (r1 * m1 - m2 * r0);
        return (float)fabs(-det / num);
    }
};

struct LaplaceSums {
    float sxy = 0.0f, sxx = 0.0f;
    __device__ __forceinline__ void add(float hypo, float depth, float prob)
    {
        const float x = fabsf(__fsub_rn(hypo, depth));
        const float y = logf(fmaxf(prob, 1e-40f));
        sxy = __fadd_rn(sxy, __fmul_rn(x, y));
        sxx = __fadd_rn(sxx, __fmul_rn(x, x));
    }
    template <int PW>
    __device__ __forceinline__ void reduce_slices()
    {
#pragma unroll
        for (int o = PW; o < 32; o <<= 1) {
            sxy = __fadd_rn(sxy, __shfl_xor_sync(0xffffffffu, sxy, o));
            sxx = __fadd_rn(sxx, __shfl_xor_sync(0xffffffffu, sxx, o));
        }
    }
    __device__ __forceinline__ float scale() const { return __fdiv_rn(1.0f, fabsf(__fdiv_rn(sxy, sxx))); }
};

template <int PW>
__device__ __forceinline__ double reduce_slices_f64(double v)
{
#pragma unroll
    for (int o = PW; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace mdf

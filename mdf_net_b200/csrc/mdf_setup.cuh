// mdf_setup.cuh -- per-call scalar work shared by every kernel of the cost-volume family:
//   proj = src_proj @ inverse(ref_proj) for every (view, batch) -- base.py:98-100 -- in float64 (no host
//   sync, unlike torch.inverse), and the eval-mode BatchNorm of depth_weight folded into two constants
//   (homoaggregate.py:16-20, base.py:50-68).
#pragma once

#include <cuda_runtime.h>

#include "mdf_common.cuh"

namespace mdf {

struct SrcPtrs { const float* p[kMaxSrcViews]; };

// ------------------------------------------------------------------------------------------------
// setup: projections + folded depth_weight parameters
// ------------------------------------------------------------------------------------------------
__device__ inline void compose_proj_f64(const float* __restrict__ src, const float* __restrict__ ref, float* __restrict__ out12)
{
    // Gauss-Jordan with partial pivoting in float64, then rows 0..2 of src @ inv(ref), rounded once.
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { a[r][c] = (double)ref[r * 4 + c]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int k = 0; k < 4; ++k) {
        int p = k; double best = fabs(a[k][k]);
        for (int r = k + 1; r < 4; ++r) { double v = fabs(a[r][k]); if (v > best) { best = v; p = r; } }
        if (p != k) for (int c = 0; c < 8; ++c) { double t = a[k][c]; a[k][c] = a[p][c]; a[p][c] = t; }
        const double inv = 1.0 / a[k][k];
        for (int c = 0; c < 8; ++c) a[k][c] *= inv;
        for (int r = 0; r < 4; ++r) {
            if (r == k) continue;
            const double f = a[r][k];
            for (int c = 0; c < 8; ++c) a[r][c] -= f * a[k][c];
        }
    }
    for (int r = 0; r < 3; ++r) {
        double row[4];
        for (int c = 0; c < 4; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += (double)src[r * 4 + k] * a[k][4 + c];
            row[c] = acc;
        }
        out12[r * 3 + 0] = (float)row[0]; out12[r * 3 + 1] = (float)row[1]; out12[r * 3 + 2] = (float)row[2];
        out12[9 + r] = (float)row[3];
    }
}

// dwp: [0]=alpha [1]=beta' [2]=fc_w [3]=fc_b [4]=beta(raw) [5]=weight of an out-of-image sample [16..16+G)=conv weight
__device__ inline void fold_depth_weight(const float* __restrict__ conv_w, const float* __restrict__ bn_w,
                                         const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                                         const float* __restrict__ bn_var, float bn_eps,
                                         const float* __restrict__ fc_w, const float* __restrict__ fc_b, int G,
                                         float* __restrict__ dwp)
{
    // eval-mode BatchNorm3d(1) folded as ATen applies it: alpha = weight/sqrt(var+eps), beta = bias - mean*alpha
    const float invstd = __frcp_rn(__fsqrt_rn(__fadd_rn(bn_var[0], bn_eps)));
    const float alpha = __fmul_rn(invstd, bn_w[0]);
    const float beta = __fsub_rn(bn_b[0], __fmul_rn(bn_mean[0], alpha));
    float cw_sum = 0.0f;
    for (int g = 0; g < G; ++g) cw_sum += conv_w[g];
    dwp[0] = alpha;
    dwp[1] = beta + alpha * 0.5f * cw_sum;   // the staged kernel accumulates sum_g cw_g*(vol_g - 0.5)
    dwp[2] = fc_w[0];
    dwp[3] = fc_b[0];
    dwp[4] = beta;
    // weight of a (sample, view) pair whose taps all fall outside the source image: every similarity is 0.5
    const float hv = fmaf(fmaxf(dwp[1], 0.0f), fc_w[0], fc_b[0]);
    dwp[5] = 1.0f / (1.0f + expf(-hv));
    for (int g = 0; g < G && g < 32; ++g) dwp[16 + g] = conv_w[g];
}

struct DepthWeightPtrs {
    const float *conv_w, *bn_w, *bn_b, *bn_mean, *bn_var, *fc_w, *fc_b;
    float bn_eps;
};

// threads [0, V*B) of the calling block compose the projections; thread 0 also folds depth_weight (if dwp)
__device__ inline void setup_work(int i, const SrcPtrs& src_projs, const float* __restrict__ ref_proj, int V, int B,
                                  float* __restrict__ rt_all, const DepthWeightPtrs& dw, int G, float* __restrict__ dwp)
{
    if (i < V * B) {
        const int v = i / B, b = i % B;
        compose_proj_f64(src_projs.p[v] + 16 * b, ref_proj + 16 * b, rt_all + (size_t)i * 12);
    }
    if (i == 0 && dwp != nullptr) fold_depth_weight(dw.conv_w, dw.bn_w, dw.bn_b, dw.bn_mean, dw.bn_var, dw.bn_eps, dw.fc_w, dw.fc_b, G, dwp);
}

static __global__ void setup_kernel(SrcPtrs src_projs, const float* __restrict__ ref_proj, int V, int B,
                             float* __restrict__ rt_all, DepthWeightPtrs dw, int G, float* __restrict__ dwp)
{
    // programmatic dependent launch: a kernel launched behind this one with the stream-serialization attribute
    // (the layout pass, which does not read anything computed here) may start right away
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    setup_work(blockIdx.x * blockDim.x + threadIdx.x, src_projs, ref_proj, V, B, rt_all, dw, G, dwp);
}

}  // namespace mdf

// mdf_fpn.cu -- the FPN hand-off: the feature pyramid's 1x1 output convolutions emit the hot kernel's
// input layout directly, and the cost volume has a second entry point that takes it.
//
// The reference's backbone ends every scale with a bias-free 1x1 convolution (net/unit/backbone.py:43-45, 59-63):
//     y = out_k(x),  x: (B, Cin, H, W) the FPN's merged map, y: (B, C = 2G, H, W) = one entry of `features`.
// The drop-in path (mdf_cost_volume_fwd) then reads y and writes the planar-float4 pair-difference maps the hot kernel
// gathers from (prep_kernel: 12 % of the step, 1.5x the feature bytes of HBM traffic).  Here the 1x1 convolution itself
// writes those maps:
//     source views:   S4[j][y][x] = (y[2g+1] - y[2g]) * log2(e),  g = 4j..4j+3
//     reference view: Q4 = 2*sigmoid(y[2g] - y[2g+1]) - 1,        CQ4 = depth_weight.0.conv.weight[g] * Q4
// (each channel sum is formed separately and the pair is subtracted afterwards, exactly as prep_kernel does from the NCHW
// features: only the summation order inside the 1x1 convolution differs from cuDNN's), and mdf_cost_volume_fwd_prepped
// runs setup + the hot kernel on them: no layout pass.  The NCHW entry point stays the drop-in.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdf_common.cuh"
#include "mdf_host.cuh"
#include "mdf_setup.cuh"
#include "mdf_staged.cuh"

namespace mdf {

// One thread = PX pixels (PX consecutive rows of 256 pixels apart: every load stays a 128-byte coalesced row), all 2G output
// channels of each in registers; the weights sit in shared memory as [o][Cin] and are fetched four input channels at a time
// (LDS.128 broadcast) -- one fetch serves PX pixels --, the inputs of eight channels are requested before the first FMA.
template <int G, int PX>
__global__ void __launch_bounds__(256)
fpn_out_prepped_kernel(const float* __restrict__ x, const float* __restrict__ w, int Cin, int HW, const float* __restrict__ conv_w,
                       float4* __restrict__ S4, float4* __restrict__ Q4, float4* __restrict__ CQ4)
{
    extern __shared__ float4 w_s[];                       // [2G][Cin / 4] float4
    constexpr int C = 2 * G, J = G / 4;
    const int q4 = Cin / 4;
    for (int k = threadIdx.x; k < C * q4; k += blockDim.x) w_s[k] = __ldg(reinterpret_cast<const float4*>(w) + k);
    __syncthreads();
    const int pix0 = blockIdx.x * (blockDim.x * PX) + threadIdx.x;
    const int b = blockIdx.y;
    if (pix0 >= HW) return;
    const float* __restrict__ xp = x + (size_t)b * Cin * HW;
    int pix[PX];
#pragma unroll
    for (int u = 0; u < PX; ++u) pix[u] = min(pix0 + u * (int)blockDim.x, HW - 1);       // a clamped duplicate is not stored
    float acc[PX][C];
#pragma unroll
    for (int u = 0; u < PX; ++u)
#pragma unroll
        for (int o = 0; o < C; ++o) acc[u][o] = 0.0f;
    for (int c8 = 0; c8 + 1 < q4 + 1; c8 += 2) {
        float xv[PX][8];
#pragma unroll
        for (int u = 0; u < PX; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) xv[u][k] = (4 * c8 + k) < Cin ? __ldg(xp + (size_t)(4 * c8 + k) * HW + pix[u]) : 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (c8 + h >= q4) break;
#pragma unroll
            for (int o = 0; o < C; ++o) {
                const float4 wv = w_s[o * q4 + c8 + h];
#pragma unroll
                for (int u = 0; u < PX; ++u)
                    acc[u][o] = fmaf(wv.w, xv[u][4 * h + 3], fmaf(wv.z, xv[u][4 * h + 2], fmaf(wv.y, xv[u][4 * h + 1], fmaf(wv.x, xv[u][4 * h], acc[u][o]))));
            }
        }
    }
#pragma unroll
    for (int u = 0; u < PX; ++u) {
        if (pix0 + u * (int)blockDim.x >= HW) break;
        if (S4 != nullptr) {                                   // a source view
            float4* __restrict__ dst = S4 + (size_t)b * J * HW + pix[u];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                float d[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) d[k] = (acc[u][2 * (4 * j + k) + 1] - acc[u][2 * (4 * j + k)]) * kLog2e;
                dst[(size_t)j * HW] = make_float4(d[0], d[1], d[2], d[3]);
            }
        } else {                                               // the reference view
            float4* __restrict__ qd = Q4 + (size_t)b * J * HW + pix[u];
            float4* __restrict__ cd = CQ4 + (size_t)b * J * HW + pix[u];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                float d[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) d[k] = 2.0f / (1.0f + expf(acc[u][2 * (4 * j + k) + 1] - acc[u][2 * (4 * j + k)])) - 1.0f;
                const float w0 = __ldg(conv_w + 4 * j), w1 = __ldg(conv_w + 4 * j + 1), w2 = __ldg(conv_w + 4 * j + 2), w3 = __ldg(conv_w + 4 * j + 3);
                qd[(size_t)j * HW] = make_float4(d[0], d[1], d[2], d[3]);
                cd[(size_t)j * HW] = make_float4(w0 * d[0], w1 * d[1], w2 * d[2], w3 * d[3]);
            }
        }
    }
}

template <int G>
static int launch_fpn(const float* x, const float* w, int Cin, int B, int HW, const float* conv_w, float4* S4, float4* Q4, float4* CQ4,
                      cudaStream_t stream)
{
    const size_t smem = (size_t)2 * G * Cin * sizeof(float);
    constexpr int PX = G == 32 ? 1 : 2;          // G16: 64 accumulators per thread; G8: 32 (4 pixels per thread leave too few blocks: 67 vs 58 us)
    auto kern = fpn_out_prepped_kernel<G, PX>;
    if (smem > 48 * 1024) MDF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)((HW + 256 * PX - 1) / (256 * PX)), (unsigned)B), 256, smem, stream>>>(x, w, Cin, HW, conv_w, S4, Q4, CQ4);
    return launch_status();
}

}  // namespace mdf

using namespace mdf;

extern "C" {

int mdf_fpn_out_prepped_fwd(const float* x, const float* out_weight, int B, int Cin, int G, int H, int W,
                            const float* depth_weight_conv, float* s4, float* q4, float* cq4, mdf_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || Cin <= 0 || G <= 0 || H < 0 || W < 0) return MDF_ERR_INVALID_SHAPE;
    if (!(G == 8 || G == 16 || G == 32) || Cin % 4 != 0 || Cin > 256 || B > 65535) return MDF_ERR_UNSUPPORTED;
    if ((size_t)B * H * W == 0) return MDF_OK;
    const bool is_ref = s4 == nullptr;
    if (!x || !out_weight || (is_ref && (!q4 || !cq4 || !depth_weight_conv)) || (!is_ref && (q4 || cq4))) return MDF_ERR_NULL_POINTER;
    if ((long long)H * W > INT_MAX - 256) return MDF_ERR_UNSUPPORTED;
    if (((uintptr_t)out_weight & 15) != 0) return MDF_ERR_UNSUPPORTED;
    const int dev = device_of(is_ref ? q4 : s4);
    if (dev < 0) return dev;
    {
        const void* ptrs[] = {x, out_weight, is_ref ? (const void*)cq4 : (const void*)s4, is_ref ? (const void*)depth_weight_conv : (const void*)x};
        const int st = check_on_device(dev, ptrs, 4);
        if (st != MDF_OK) return st;
    }
    DeviceGuard guard(dev);
    const int HW = H * W;
    float4 *S = reinterpret_cast<float4*>(s4), *Q = reinterpret_cast<float4*>(q4), *CQ = reinterpret_cast<float4*>(cq4);
    if (G == 32) return launch_fpn<32>(x, out_weight, Cin, B, HW, depth_weight_conv, S, Q, CQ, stream);
    if (G == 16) return launch_fpn<16>(x, out_weight, Cin, B, HW, depth_weight_conv, S, Q, CQ, stream);
    return launch_fpn<8>(x, out_weight, Cin, B, HW, depth_weight_conv, S, Q, CQ, stream);
}

size_t mdf_cost_volume_prepped_workspace_bytes(int B, int N)
{
    if (B <= 0 || N < 2) return 0;
    return align_up((size_t)(N - 1) * B * 12 * sizeof(float), 256) + align_up(64 * sizeof(float), 256);
}

int mdf_cost_volume_fwd_prepped(const float* s4, const float* q4, const float* cq4, int N, const float* ref_proj,
                                const float* const* src_projs, const float* depth_hypos, int hypos_per_pixel,
                                const float* conv_weight, const float* bn_weight, const float* bn_bias, const float* bn_mean,
                                const float* bn_var, float bn_eps, const float* fc_weight, const float* fc_bias,
                                int B, int G, int D, int H, int W, float* cost_volume, void* workspace, size_t workspace_bytes,
                                mdf_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || G <= 0 || D < 0 || H < 0 || W < 0 || N < 2) return MDF_ERR_INVALID_SHAPE;
    if (N > MDF_MAX_VIEWS || !(G == 8 || G == 16 || G == 32)) return MDF_ERR_UNSUPPORTED;
    if ((size_t)B * D * H * W == 0) return MDF_OK;
    if (!s4 || !q4 || !cq4 || !src_projs || !ref_proj || !depth_hypos || !conv_weight || !bn_weight || !bn_bias || !bn_mean ||
        !bn_var || !fc_weight || !fc_bias || !cost_volume)
        return MDF_ERR_NULL_POINTER;
    const int V = N - 1;
    if ((long long)H * W > INT_MAX - 256 || (long long)V * B > 256) return MDF_ERR_UNSUPPORTED;
    if (!workspace || ((uintptr_t)workspace & 255) != 0 || workspace_bytes < mdf_cost_volume_prepped_workspace_bytes(B, N))
        return MDF_ERR_WORKSPACE;
    if ((((uintptr_t)s4 | (uintptr_t)q4 | (uintptr_t)cq4) & 15) != 0) return MDF_ERR_UNSUPPORTED;      // TMA needs 16-byte aligned maps
    const int dev = device_of(cost_volume);
    if (dev < 0) return dev;
    {
        const void* ptrs[MDF_MAX_VIEWS + 16];
        int n = 0;
        for (int i = 0; i < V; ++i) ptrs[n++] = src_projs[i];
        const void* more[] = {s4, q4, cq4, ref_proj, depth_hypos, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, fc_weight, fc_bias, workspace};
        for (const void* p : more) ptrs[n++] = p;
        const int st = check_on_device(dev, ptrs, n);
        if (st != MDF_OK) return st;
    }
    DeviceGuard guard(dev);
    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    float* rt = reinterpret_cast<float*>(wsb);
    float* dwp = reinterpret_cast<float*>(wsb + align_up((size_t)V * B * 12 * sizeof(float), 256));
    SrcPtrs sp;
    for (int v = 0; v < kMaxSrcViews; ++v) sp.p[v] = v < V ? src_projs[v] : nullptr;
    const DepthWeightPtrs dw = {conv_weight, bn_weight, bn_bias, bn_mean, bn_var, fc_weight, fc_bias, bn_eps};
    setup_kernel<<<(V * B + 63) / 64, 64, 0, stream>>>(sp, ref_proj, V, B, rt, dw, G, dwp);
    int st = launch_status();
    if (st != MDF_OK) return st;
    StagedArgs a;
    a.rt = rt; a.dwp = dwp; a.hypos = depth_hypos; a.out = cost_volume; a.vparams = nullptr; a.stats = nullptr;
    a.per_pixel = hypos_per_pixel; a.V = V; a.B = B; a.D = D; a.H = H; a.W = W;
    a.gn = make_grid_norm(H, W);
    a.tiles_x = a.tiles_y = a.slabs = 0;
    StagedBuffers buf;
    buf.S4 = s4; buf.Q4 = q4; buf.CQ4 = cq4;
    return launch_staged_eval(G, a, buf, stream);
}

}  // extern "C"

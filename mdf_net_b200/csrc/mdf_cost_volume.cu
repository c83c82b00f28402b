// mdf_cost_volume.cu -- plane-sweep cost volume of MDF-Net for sm_100a.
//
// Replaces, in one pass per stage, what the reference does with ~40 ATen launches per source view:
//   homo_warping                 net/unit/base.py:85-126
//   VectorAggregate.forward      net/unit/homoaggregate.py:25-46  (eval-mode depth_weight :16-20)
//   homo_aggregate_by_variance   net/unit/homoaggregate.py:49-69
//
// Kernels in this file
//   setup_kernel            per call: proj = src_proj @ inverse(ref_proj) for every (view, batch) in
//                           float64 (no host sync, unlike torch.inverse at base.py:98) and the folded
//                           eval-mode BatchNorm of depth_weight.
//   prep_kernel, cost_volume_staged   the hot path for C/G == 2: see mdf_staged.cuh
//   cost_volume_direct      any C/G: taps straight from the NCHW features (no staging), two passes.
//   homo_warp / variance    the standalone warp and the (unused by config.py) variance aggregate.
#include <cuda.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "../../include/mdf_b200_debug.h"
#include "mdf_common.cuh"
#include "mdf_host.cuh"
#include "mdf_setup.cuh"
#include "mdf_staged.cuh"
#ifdef MDF_TUNING
#include "experimental/mdf_pipe.cuh"
#endif

namespace mdf {

thread_local int g_last_cuda_error = 0;     // errno-like: the code behind the last MDF_ERR_CUDA of this thread

// ------------------------------------------------------------------------------------------------
// workspace layout (all offsets 256-byte aligned)
// ------------------------------------------------------------------------------------------------
struct Workspace {
    size_t rt_off;    // [V][B][12] float   rot|trans per source view
    size_t dwp_off;   // 64 floats: [0]=alpha [1]=beta' [2]=fc_w [3]=fc_b [4]=beta(raw) [5]=weight of an out-of-image sample
    size_t q_off;     // [B][G/4][H][W] float4 reference q maps           (staged path)
    size_t s_off;     // [V][B][G/4][H][W] float4 source difference maps  (staged path)
    size_t cq_off;    // [B][G/4][H][W] float4 conv_w * q                 (staged path)
    size_t total;
};

static Workspace make_workspace(int B, int N, int G, int H, int W, bool staged)
{
    Workspace w;
    const size_t V = (size_t)(N - 1);
    size_t off = 0;
    w.rt_off = off;  off = align_up(off + V * B * 12 * sizeof(float), 256);
    w.dwp_off = off; off = align_up(off + 64 * sizeof(float), 256);
    w.q_off = off;
    if (staged) off = align_up(off + (size_t)B * G * H * W * sizeof(float), 256);
    w.s_off = off;
    if (staged) off = align_up(off + V * B * (size_t)H * W * G * sizeof(float), 256);
    w.cq_off = off;
    if (staged) off = align_up(off + (size_t)B * G * H * W * sizeof(float), 256);
    w.total = off;
    return w;
}

// ------------------------------------------------------------------------------------------------
// direct kernel: any C/G.  One thread per (b, d, y, x); pass 1 computes the view weights, pass 2
// recomputes the similarities and forms the weighted mean (keeps registers independent of G).
// ------------------------------------------------------------------------------------------------
struct DirectArgs {
    FeaPtrs fea;          // [0] = reference
    const float* rt;      // [V][B][12]
    const float* dwp;     // [0]=alpha [4]=beta(raw) [2]=fc_w [3]=fc_b
    const float* conv_w;
    const float* hypos;
    float* out;
    int per_pixel, V, B, C, G, D, H, W;
};

__device__ __forceinline__ float group_similarity(const float* __restrict__ ref, const float* __restrict__ src,
                                                  size_t HW, size_t ref_off, int cpg, int H, int W, const Taps& t)
{
    // softmax over the cpg channels of the group for both views, then the dot product
    float rmax = -INFINITY, smax = -INFINITY;
    for (int k = 0; k < cpg; ++k) {
        rmax = fmaxf(rmax, __ldg(ref + k * HW + ref_off));
        smax = fmaxf(smax, sample_plane(src + k * HW, H, W, t));
    }
    float rs = 0.0f, ss = 0.0f, dot = 0.0f;
    for (int k = 0; k < cpg; ++k) {
        const float er = expf(__ldg(ref + k * HW + ref_off) - rmax);
        const float es = expf(sample_plane(src + k * HW, H, W, t) - smax);
        rs += er; ss += es; dot = fmaf(er, es, dot);
    }
    return dot / (rs * ss);
}

__global__ void __launch_bounds__(256)
cost_volume_direct_kernel(const DirectArgs a)
{
    const size_t HW = (size_t)a.H * a.W;
    const size_t total = (size_t)a.B * a.D * HW;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % a.W);
    const int y = (int)((idx / a.W) % a.H);
    const int d = (int)((idx / HW) % a.D);
    const int b = (int)(idx / (HW * a.D));
    const int cpg = a.C / a.G;
    const GridNorm gn = make_grid_norm(a.H, a.W);
    const float depth = a.per_pixel ? __ldg(a.hypos + ((size_t)b * a.D + d) * HW + (size_t)y * a.W + x)
                                    : __ldg(a.hypos + (size_t)b * a.D + d);
    const float alpha = __ldg(a.dwp + 0), beta = __ldg(a.dwp + 4), fcw = __ldg(a.dwp + 2), fcb = __ldg(a.dwp + 3);
    const size_t ref_off = (size_t)y * a.W + x;
    const float* ref = a.fea.p[0] + (size_t)b * a.C * HW;

    float wv[kMaxSrcViews];
    float wsum = 0.0f;
    for (int v = 0; v < a.V; ++v) {
        const float* rt = a.rt + ((size_t)v * a.B + b) * 12;
        float ix, iy;
        sample_position(rot_xyz(rt, (float)x, (float)y), rt, depth, gn, ix, iy);
        const Taps t = make_taps(ix, iy, gn);
        const float* src = a.fea.p[v + 1] + (size_t)b * a.C * HW;
        float z = 0.0f;
        for (int g = 0; g < a.G; ++g)
            z = fmaf(__ldg(a.conv_w + g), group_similarity(ref + (size_t)g * cpg * HW, src + (size_t)g * cpg * HW, HW, ref_off, cpg, a.H, a.W, t), z);
        float h = fmaf(z, alpha, beta);
        h = fmaxf(h, 0.0f);
        h = fmaf(h, fcw, fcb);
        const float w = 1.0f / (1.0f + expf(-h));
        wv[v] = w;
        wsum += w;
    }
    for (int g = 0; g < a.G; ++g) {
        float vs = 0.0f;
        for (int v = 0; v < a.V; ++v) {
            const float* rt = a.rt + ((size_t)v * a.B + b) * 12;
            float ix, iy;
            sample_position(rot_xyz(rt, (float)x, (float)y), rt, depth, gn, ix, iy);
            const Taps t = make_taps(ix, iy, gn);
            const float* src = a.fea.p[v + 1] + (size_t)b * a.C * HW;
            vs = fmaf(wv[v], group_similarity(ref + (size_t)g * cpg * HW, src + (size_t)g * cpg * HW, HW, ref_off, cpg, a.H, a.W, t), vs);
        }
        a.out[(((size_t)b * a.G + g) * a.D + d) * HW + ref_off] = vs / wsum;
    }
}

// ------------------------------------------------------------------------------------------------
// homo_warping (base.py:85-126): one thread per (b, d, y, x) computes the sample position and the four tap
// weights once, then walks the C channel planes.  The op is bound by writing the (B,C,D,H,W) volume (354 MB per
// call at BASELINE configs[1]): the channel loop is unrolled by 8 so that 32 independent tap loads (L1 / L2 hits:
// neighbouring lanes read neighbouring texels) are in flight per thread and the stores -- 128-byte rows per warp,
// written once (plain stores: a streaming hint measured 7 % slower, 16 channels in flight 14 % slower) -- keep HBM busy.  Warps whose 32 footprints are all inside the source map
// (the common case) take a path without bounds predicates.
// ------------------------------------------------------------------------------------------------
template <bool INTERIOR>
__device__ __forceinline__ void warp_channels(const float* __restrict__ p, float* __restrict__ o, int C, size_t HW, size_t ostride,
                                              int W, const Taps& t, bool x0in, bool x1in, bool y0in, bool y1in, bool live)
{
    constexpr int U = 8;
    // running pointers (north row, south row, output), advanced by one plane per channel: the loads use immediate
    // offsets and the loop carries no 64-bit multiplications (the first version spent more issue slots on address
    // arithmetic than on the taps)
    const float* __restrict__ pn = p;
    const float* __restrict__ ps = p + W;
    float* __restrict__ po = o;
    int c0 = 0;
    if (INTERIOR) {
        // full chunks of U channels, no predicate anywhere: 4*U loads, then U blends and stores
        for (; c0 + U <= C; c0 += U) {
            float nw[U], ne[U], sw[U], se[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                nw[u] = __ldg(pn); ne[u] = __ldg(pn + 1); sw[u] = __ldg(ps); se[u] = __ldg(ps + 1);
                pn += HW;
                ps += HW;
            }
            if (live) {
#pragma unroll
                for (int u = 0; u < U; ++u) po[(size_t)u * ostride] = blend4(nw[u], ne[u], sw[u], se[u], t);
            }
            po += (size_t)U * ostride;
        }
    }
    for (; c0 < C; c0 += U) {
        float nw[U], ne[U], sw[U], se[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool on = c0 + u < C;
            nw[u] = (on && (INTERIOR || (x0in && y0in))) ? __ldg(pn) : 0.0f;
            sw[u] = (on && (INTERIOR || (x0in && y1in))) ? __ldg(ps) : 0.0f;
            ne[u] = (on && (INTERIOR || (x1in && y0in))) ? __ldg(pn + 1) : 0.0f;
            se[u] = (on && (INTERIOR || (x1in && y1in))) ? __ldg(ps + 1) : 0.0f;
            pn += HW;
            ps += HW;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (live && c0 + u < C) *po = blend4(nw[u], ne[u], sw[u], se[u], t);
            po += ostride;
        }
    }
}

__global__ void __launch_bounds__(256)
homo_warp_kernel(const float* __restrict__ src, const float* __restrict__ rt_all, const float* __restrict__ hypos,
                 int per_pixel, int B, int C, int D, int H, int W, float* __restrict__ out)
{
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)B * D * HW;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < total;
    const size_t i = live ? idx : total - 1;          // idle lanes of the last warp shadow a real sample (the votes below are warp wide)
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int d = (int)((i / HW) % D);
    const int b = (int)(i / (HW * D));
    const GridNorm gn = make_grid_norm(H, W);
    const float depth = per_pixel ? __ldg(hypos + ((size_t)b * D + d) * HW + (size_t)y * W + x) : __ldg(hypos + (size_t)b * D + d);
    const float* rt = rt_all + (size_t)b * 12;
    float ix, iy;
    sample_position(rot_xyz(rt, (float)x, (float)y), rt, depth, gn, ix, iy);
    const Taps t = make_taps(ix, iy, gn);
    const bool x0in = t.valid && (unsigned)t.x0 < (unsigned)W, x1in = t.valid && (unsigned)(t.x0 + 1) < (unsigned)W;
    const bool y0in = (unsigned)t.y0 < (unsigned)H, y1in = (unsigned)(t.y0 + 1) < (unsigned)H;
    const bool interior = x0in && x1in && y0in && y1in;
    const float* p = src + (size_t)b * C * HW + (ptrdiff_t)t.y0 * W + t.x0;
    float* o = out + (((size_t)b * C * D + d) * H + y) * W + x;
    const size_t ostride = (size_t)D * HW;
    if (__all_sync(0xffffffffu, interior)) warp_channels<true>(p, o, C, HW, ostride, W, t, true, true, true, true, live);
    else warp_channels<false>(p, o, C, HW, ostride, W, t, x0in, x1in, y0in, y1in, live);
}

// ------------------------------------------------------------------------------------------------
// homo_aggregate_by_variance (homoaggregate.py:49-69).  One thread per (b, d, y, x).  The softmax
// runs over all C channels of each warped view, so the thread first finds (max, sum) per view and
// then re-samples channel by channel -- registers stay independent of C.
// ------------------------------------------------------------------------------------------------
struct VarArgs {
    FeaPtrs fea;
    const float* rt;
    const float* hypos;
    float* out;
    int per_pixel, V, B, C, D, H, W;
};

// Footprint of one (sample, source view): base pointer of channel 0 (clamped into the map), weights, tap validity.
struct VarTap {
    const float* p;
    Taps t;
    bool nw, ne, sw, se;
};

__device__ __forceinline__ VarTap var_tap(const VarArgs& a, int v, int b, int x, int y, float depth, const GridNorm& gn)
{
    const size_t HW = (size_t)a.H * a.W;
    const float* rt = a.rt + ((size_t)v * a.B + b) * 12;
    float ix, iy;
    sample_position(rot_xyz(rt, (float)x, (float)y), rt, depth, gn, ix, iy);
    VarTap f;
    f.t = make_taps(ix, iy, gn);
    const bool x0in = f.t.valid && (unsigned)f.t.x0 < (unsigned)a.W, x1in = f.t.valid && (unsigned)(f.t.x0 + 1) < (unsigned)a.W;
    const bool y0in = (unsigned)f.t.y0 < (unsigned)a.H, y1in = (unsigned)(f.t.y0 + 1) < (unsigned)a.H;
    f.nw = x0in && y0in; f.ne = x1in && y0in; f.sw = x0in && y1in; f.se = x1in && y1in;
    f.p = a.fea.p[v + 1] + (size_t)b * a.C * HW + (ptrdiff_t)f.t.y0 * a.W + f.t.x0;
    return f;
}

// U consecutive channels of one footprint: 4*U independent loads in flight
template <int U>
__device__ __forceinline__ void var_sample(const VarTap& f, int c0, int C, size_t HW, int W, float (&out)[U])
{
    float nw[U], ne[U], sw[U], se[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const bool live = c0 + u < C;
        const float* q = f.p + (size_t)(c0 + u) * HW;
        nw[u] = (live && f.nw) ? __ldg(q) : 0.0f;
        ne[u] = (live && f.ne) ? __ldg(q + 1) : 0.0f;
        sw[u] = (live && f.sw) ? __ldg(q + W) : 0.0f;
        se[u] = (live && f.se) ? __ldg(q + W + 1) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) out[u] = blend4(nw[u], ne[u], sw[u], se[u], f.t);
}

// CMAX >= C: every warped view is sampled ONCE -- its C channel values stay in registers for the softmax (max, sum of
// exponentials, the probabilities themselves), and the per-channel sum / sum of squares over the views are register
// arrays too.  CMAX = 64 costs ~200 registers (one block per SM) but a third of the taps and half of the exponentials of
// the chunked version below, which remains for C > 64.
template <int CMAX>
__global__ void __launch_bounds__(256)
variance_volume_reg_kernel(const VarArgs a)
{
    constexpr int U = 8;
    static_assert(CMAX % U == 0, "chunks of 8 channels");
    const size_t HW = (size_t)a.H * a.W;
    const size_t total = (size_t)a.B * a.D * HW;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % a.W);
    const int y = (int)((idx / a.W) % a.H);
    const int d = (int)((idx / HW) % a.D);
    const int b = (int)(idx / (HW * a.D));
    const int C = a.C;
    const GridNorm gn = make_grid_norm(a.H, a.W);
    const float depth = a.per_pixel ? __ldg(a.hypos + ((size_t)b * a.D + d) * HW + (size_t)y * a.W + x)
                                    : __ldg(a.hypos + (size_t)b * a.D + d);
    float s1[CMAX], s2[CMAX];
    const float* ref = a.fea.p[0] + (size_t)b * C * HW + (size_t)y * a.W + x;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        const float rv = c < C ? __ldg(ref + (size_t)c * HW) : 0.0f;        // the raw reference feature (homoaggregate.py:56-57)
        s1[c] = rv;
        s2[c] = rv * rv;
    }
    for (int v = 0; v < a.V; ++v) {
        const VarTap f = var_tap(a, v, b, x, y, depth, gn);
        float val[CMAX];
#pragma unroll
        for (int c0 = 0; c0 < CMAX; c0 += U) {
            float chunk[U];
            var_sample<U>(f, c0, C, HW, a.W, chunk);
#pragma unroll
            for (int u = 0; u < U; ++u) val[c0 + u] = chunk[u];
        }
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) m = fmaxf(m, val[c]);
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) { val[c] = expf(val[c] - m); s += val[c]; }          // channel order, as the chunked kernel
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) {
                const float pr = val[c] / s;
                s1[c] += pr;
                s2[c] = fmaf(pr, pr, s2[c]);
            }
    }
    const float nviews = (float)(a.V + 1);
    float* o = a.out + (((size_t)b * C * a.D + d) * a.H + y) * a.W + x;
    const size_t ostride = (size_t)a.D * HW;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
        if (c < C) {
            const float mean = s1[c] / nviews;
            o[(size_t)c * ostride] = s2[c] / nviews - mean * mean;
        }
}

// ---- the fast variant (C = 16 / 32 / 64, the reference's channel counts): C/16 lanes share a sample ----------------
// A layout pass writes the source views as channel quads, planar: P4[v][b][C/4][H][W] float4, so a tap of four channels is
// one 16-byte load (the NCHW kernels above issue four scalar loads with four address computations for it).  A thread keeps
// 16 channels of the softmax and of the two running sums in registers (3 x 16 instead of 3 x 64 at C = 64: three blocks per
// SM instead of one); with C = 32 / 64 a warp is 16 / 8 consecutive samples x 2 / 4 channel parts and the maximum and the sum
// of exponentials of a view meet over the lanes of a sample by shuffles.  The bilinear blend keeps the reference's tap order;
// the softmax runs on the MUFU (2^x and ONE reciprocal of the sum per view instead of expf and C true divisions).
__global__ void __launch_bounds__(256)
planar4_kernel(FeaPtrs fea, int V, int B, int C, int HW, float4* __restrict__ out)
{
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const int C4 = C / 4;
    const int c4 = blockIdx.y % C4, b = (blockIdx.y / C4) % B, v = blockIdx.y / (C4 * B);
    const float* f = fea.p[v + 1] + ((size_t)b * C + 4 * c4) * HW + pix;
    out[(((size_t)v * B + b) * C4 + c4) * HW + pix] = make_float4(__ldg(f), __ldg(f + HW), __ldg(f + 2 * (size_t)HW), __ldg(f + 3 * (size_t)HW));
}

template <int OFF>
__device__ __forceinline__ float4 ldg4_pred(const float4* p, bool pred)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4+%6];\n\t}"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) : "l"(p), "r"((int)pred), "n"(OFF));
    return v;
}

template <int LANES>   // lanes per sample: C = 16 * LANES channels, 16 per thread
__global__ void __launch_bounds__(256, 3)
variance_volume_quad_kernel(const VarArgs a, const float4* __restrict__ P4)
{
    constexpr int CQ = 16, Q = CQ / 4, SPW = 32 / LANES;       // channels / float4 quads per thread, samples per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int part = lane / SPW, slot = lane % SPW;  // which 16 channels; which sample of the warp
    const uint32_t HW = (uint32_t)a.H * a.W;
    const uint32_t total = (uint32_t)a.B * a.D * HW; // < 2^31 (host checked): 32-bit index arithmetic
    const uint32_t sidx = blockIdx.x * (8u * SPW) + warp * SPW + slot;
    const bool live = sidx < total;
    const uint32_t idx = live ? sidx : total - 1;    // idle lanes shadow the last sample (the shuffles below are warp wide)
    const uint32_t pix = idx % HW, bd = idx / HW;
    const int x = (int)(pix % (uint32_t)a.W), y = (int)(pix / (uint32_t)a.W);
    const int d = (int)(bd % (uint32_t)a.D), b = (int)(bd / (uint32_t)a.D);
    constexpr int C = CQ * LANES, C4 = C / 4;
    const int c0 = part * CQ;
    const GridNormFast gf = make_grid_norm_fast(a.H, a.W);       // the bit-exact position chain on shared reciprocals (mdf_common.cuh)
    const float depth = a.per_pixel ? __ldg(a.hypos + (size_t)bd * HW + pix) : __ldg(a.hypos + bd);
    float s1[CQ], s2[CQ];
    const float* ref = a.fea.p[0] + ((size_t)b * C + c0) * HW + pix;
#pragma unroll
    for (int c = 0; c < CQ; ++c) {
        const float rv = __ldg(ref + (size_t)c * HW);                       // the raw reference feature (homoaggregate.py:56-57)
        s1[c] = rv;
        s2[c] = rv * rv;
    }
    for (int v0 = 0; v0 < a.V; v0 += LANES) {
        // the LANES lanes of a sample find the positions of LANES views, one each, and hand them round by shuffle
        float myx, myy;
        {
            const int mv = min(v0 + part, a.V - 1);
            float rt[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) rt[k] = __ldg(a.rt + ((size_t)mv * a.B + b) * 12 + k);
            sample_position_fast(rot_xyz(rt, (float)x, (float)y), rt, depth, gf, myx, myy);
        }
#pragma unroll
        for (int k = 0; k < LANES; ++k) {
            const int v = v0 + k;
            if (v >= a.V) break;                     // warp uniform
            const float ix = LANES == 1 ? myx : __shfl_sync(0xffffffffu, myx, slot + k * SPW);
            const float iy = LANES == 1 ? myy : __shfl_sync(0xffffffffu, myy, slot + k * SPW);
            const Taps t = make_taps(ix, iy, gf.g);
            const bool x0in = t.valid && (unsigned)t.x0 < (unsigned)a.W, x1in = t.valid && (unsigned)(t.x0 + 1) < (unsigned)a.W;
            const bool y0in = (unsigned)t.y0 < (unsigned)a.H, y1in = (unsigned)(t.y0 + 1) < (unsigned)a.H;
            const bool inw = x0in && y0in, ine = x1in && y0in, isw = x0in && y1in, ise = x1in && y1in;
            const float4* pn = P4 + (((size_t)v * a.B + b) * C4 + c0 / 4) * HW + (t.y0 * a.W + t.x0);
            const float4* ps = pn + a.W;
            float val[CQ];
#pragma unroll
            for (int q = 0; q < Q; ++q, pn += HW, ps += HW) {
                const float4 nw = ldg4_pred<0>(pn, inw), ne = ldg4_pred<16>(pn, ine);
                const float4 sw = ldg4_pred<0>(ps, isw), se = ldg4_pred<16>(ps, ise);
                val[4 * q + 0] = blend4(nw.x, ne.x, sw.x, se.x, t); val[4 * q + 1] = blend4(nw.y, ne.y, sw.y, se.y, t);
                val[4 * q + 2] = blend4(nw.z, ne.z, sw.z, se.z, t); val[4 * q + 3] = blend4(nw.w, ne.w, sw.w, se.w, t);
            }
            float m = val[0];
#pragma unroll
            for (int c = 1; c < CQ; ++c) m = fmaxf(m, val[c]);
#pragma unroll
            for (int o = SPW; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            // exp(v - m) = 2^((v - m) log2e) on the MUFU (|error| <= ~3 ulp of the largest term: the terms that lose relative
            // precision are the ones that are tiny next to it), one reciprocal of the sum instead of C divisions
            const float ml = m * kLog2e;
            float sum = 0.0f;
#pragma unroll
            for (int c = 0; c < CQ; ++c) { val[c] = ex2_approx(fmaf(val[c], kLog2e, -ml)); sum += val[c]; }
#pragma unroll
            for (int o = SPW; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);      // the same value in all lanes of a sample
            const float rs = __frcp_rn(sum);
#pragma unroll
            for (int c = 0; c < CQ; ++c) {
                const float pr = val[c] * rs;
                s1[c] += pr;
                s2[c] = fmaf(pr, pr, s2[c]);
            }
        }
    }
    if (!live) return;
    // volume_sq_sum / n - (volume_sum / n)^2 (homoaggregate.py:66): n is a small integer, the divisions are exact-rounded
    // through its refined reciprocal (div_by, mdf_common.cuh)
    const float nviews = (float)(a.V + 1), rn = refine_rcp(nviews);
    float* o = a.out + ((size_t)b * C + c0) * ((size_t)a.D * HW) + (size_t)d * HW + pix;
    const size_t ostride = (size_t)a.D * HW;
#pragma unroll
    for (int c = 0; c < CQ; ++c) {
        const float mean = div_by(s1[c], nviews, rn);
        o[(size_t)c * ostride] = __fsub_rn(div_by(s2[c], nviews, rn), __fmul_rn(mean, mean));
    }
}

__global__ void __launch_bounds__(256)
variance_volume_kernel(const VarArgs a)
{
    constexpr int U = 8;
    const size_t HW = (size_t)a.H * a.W;
    const size_t total = (size_t)a.B * a.D * HW;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % a.W);
    const int y = (int)((idx / a.W) % a.H);
    const int d = (int)((idx / HW) % a.D);
    const int b = (int)(idx / (HW * a.D));
    const GridNorm gn = make_grid_norm(a.H, a.W);
    const float depth = a.per_pixel ? __ldg(a.hypos + ((size_t)b * a.D + d) * HW + (size_t)y * a.W + x)
                                    : __ldg(a.hypos + (size_t)b * a.D + d);
    // softmax statistics of every warped view over its C channels (homoaggregate.py:60): max, then the sum of
    // exp(v - max) in channel order; the position and the footprint are computed once per view and pass
    float vmax[kMaxSrcViews], vinv[kMaxSrcViews];
    for (int v = 0; v < a.V; ++v) {
        const VarTap f = var_tap(a, v, b, x, y, depth, gn);
        float m = -INFINITY;
        for (int c0 = 0; c0 < a.C; c0 += U) {
            float val[U];
            var_sample<U>(f, c0, a.C, HW, a.W, val);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c0 + u < a.C) m = fmaxf(m, val[u]);
        }
        float s = 0.0f;
        for (int c0 = 0; c0 < a.C; c0 += U) {
            float val[U];
            var_sample<U>(f, c0, a.C, HW, a.W, val);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c0 + u < a.C) s += expf(val[u] - m);
        }
        vmax[v] = m;
        vinv[v] = s;
    }
    const float nviews = (float)(a.V + 1);
    const float* ref = a.fea.p[0] + (size_t)b * a.C * HW + (size_t)y * a.W + x;
    float* o = a.out + (((size_t)b * a.C * a.D + d) * a.H + y) * a.W + x;
    const size_t ostride = (size_t)a.D * HW;
    // U channels at a time: sum and sum of squares over the views, in view order (homoaggregate.py:56-67)
    for (int c0 = 0; c0 < a.C; c0 += U) {
        float s1[U], s2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float rv = c0 + u < a.C ? __ldg(ref + (size_t)(c0 + u) * HW) : 0.0f;
            s1[u] = rv;
            s2[u] = rv * rv;
        }
        for (int v = 0; v < a.V; ++v) {
            const VarTap f = var_tap(a, v, b, x, y, depth, gn);
            float val[U];
            var_sample<U>(f, c0, a.C, HW, a.W, val);
            const float m = vmax[v], inv = vinv[v];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float pr = expf(val[u] - m) / inv;
                s1[u] += pr;
                s2[u] = fmaf(pr, pr, s2[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (c0 + u >= a.C) break;
            const float mean = s1[u] / nviews;
            o[(size_t)(c0 + u) * ostride] = s2[u] / nviews - mean * mean;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// diagnostic: the hot kernel's coordinate chain (sample_position_fast), written out for the parity
// test that pins it bit for bit on the oracle's positions.  Not used by any product path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sample_positions_kernel(const float* __restrict__ rt, const float* __restrict__ hypos, int per_pixel, int D, int H, int W,
                        float* __restrict__ ix_out, float* __restrict__ iy_out)
{
    const size_t HW = (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)D * HW) return;
    const int x = (int)(idx % W), y = (int)((idx / W) % H), d = (int)(idx / HW);
    const GridNormFast gf = make_grid_norm_fast(H, W);
    float r12[12];
    for (int k = 0; k < 12; ++k) r12[k] = __ldg(rt + k);
    const float depth = per_pixel ? __ldg(hypos + idx) : __ldg(hypos + d);
    float ix, iy;
    sample_position_fast(rot_xyz(r12, (float)x, (float)y), r12, depth, gf, ix, iy);
    ix_out[idx] = ix;
    iy_out[idx] = iy;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int run_setup(const float* const* src_projs, const float* ref_proj, int V, int B, float* rt,
                     const float* conv_w, const float* bn_w, const float* bn_b, const float* bn_mean,
                     const float* bn_var, float bn_eps, const float* fc_w, const float* fc_b, int G, float* dwp,
                     cudaStream_t stream)
{
    SrcPtrs sp;
    for (int v = 0; v < kMaxSrcViews; ++v) sp.p[v] = v < V ? src_projs[v] : nullptr;
    const int n = V * B;
    const DepthWeightPtrs dw = {conv_w, bn_w, bn_b, bn_mean, bn_var, fc_w, fc_b, bn_eps};
    setup_kernel<<<(n + 63) / 64, 64, 0, stream>>>(sp, ref_proj, V, B, rt, dw, G, dwp);
    return launch_status();
}

int launch_staged_eval(int G, const StagedArgs& a, const StagedBuffers& S, cudaStream_t stream)
{
    return launch_staged_default<0>(G, a, S, stream);
}

}  // namespace mdf

using namespace mdf;

extern "C" {

int mdf_abi_version(void) { return MDF_ABI_VERSION; }

int mdf_last_cuda_error(void) { return g_last_cuda_error; }

const char* mdf_status_string(int status)
{
    switch (status) {
        case MDF_OK: return "ok";
        case MDF_ERR_INVALID_SHAPE: return "invalid shape";
        case MDF_ERR_UNSUPPORTED: return "unsupported configuration";
        case MDF_ERR_NULL_POINTER: return "null pointer";
        case MDF_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
        case MDF_ERR_NOT_DEVICE: return "pointer is not device memory on the output's device (no CPU fallback)";
        case MDF_ERR_CUDA: return "CUDA error (see mdf_last_cuda_error)";
        default: return "unknown status";
    }
}

static bool staged_supported(int C, int G) { return C == 2 * G && (G == 32 || G == 16 || G == 8); }

size_t mdf_cost_volume_workspace_bytes(int B, int N, int C, int G, int D, int H, int W)
{
    (void)D;
    if (B <= 0 || N < 2 || C <= 0 || G <= 0 || H <= 0 || W <= 0) return 0;
    return make_workspace(B, N, G, H, W, staged_supported(C, G)).total;
}

int mdf_cost_volume_fwd_ex(const float* const* features, int N, const float* ref_proj, const float* const* src_projs,
                           const float* depth_hypos, int hypos_per_pixel, const float* conv_weight,
                           const float* bn_weight, const float* bn_bias, const float* bn_mean, const float* bn_var,
                           float bn_eps, const float* fc_weight, const float* fc_bias, int B, int C, int G, int D,
                           int H, int W, float* cost_volume, void* workspace, size_t workspace_bytes, int algo,
                           void* hot_start_event, void* hot_stop_event, mdf_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || C <= 0 || G <= 0 || D < 0 || H < 0 || W < 0 || N < 2 || C % G != 0) return MDF_ERR_INVALID_SHAPE;
    if (N > MDF_MAX_VIEWS || C / G > 16) return MDF_ERR_UNSUPPORTED;
    if ((size_t)B * D * H * W == 0) return MDF_OK;   // empty volume
    if (!features || !src_projs || !ref_proj || !depth_hypos || !conv_weight || !bn_weight || !bn_bias || !bn_mean ||
        !bn_var || !fc_weight || !fc_bias || !cost_volume)
        return MDF_ERR_NULL_POINTER;
    const int V = N - 1;
    // the staged path needs the grid of its layout pass and its tensor maps to stay in range; beyond that the direct kernel
    // serves the call when the caller did not insist on the staged one
    const bool staged_fits = (long long)H * W <= INT_MAX - 256 && (long long)N * B <= 65535 && (long long)V * B <= 256;
    if (algo == 1 && !staged_fits) return MDF_ERR_UNSUPPORTED;
    const bool staged = staged_supported(C, G) && algo != 2 && staged_fits;
    if ((algo == 1 || algo >= 16) && !staged) return MDF_ERR_UNSUPPORTED;
    if (algo < 0 || (algo > 2 && algo < 16)) return MDF_ERR_UNSUPPORTED;
#ifndef MDF_TUNING
    if (algo >= 16) return MDF_ERR_UNSUPPORTED;      // tuning variants exist in tuning builds only
#endif
    if ((hot_start_event == nullptr) != (hot_stop_event == nullptr)) return MDF_ERR_NULL_POINTER;
    const Workspace ws = make_workspace(B, N, G, H, W, staged);
    if (!workspace || ((uintptr_t)workspace & 255) != 0 || workspace_bytes < ws.total) return MDF_ERR_WORKSPACE;

    const int dev = device_of(cost_volume);
    if (dev < 0) return dev;
    {
        const void* ptrs[MDF_MAX_VIEWS * 2 + 16];
        int n = 0;
        for (int i = 0; i < N; ++i) ptrs[n++] = features[i];
        for (int i = 0; i < V; ++i) ptrs[n++] = src_projs[i];
        const void* more[] = {ref_proj, depth_hypos, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, fc_weight, fc_bias, workspace};
        for (const void* p : more) ptrs[n++] = p;
        int st = check_on_device(dev, ptrs, n);
        if (st != MDF_OK) return st;
    }
    DeviceGuard guard(dev);

    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    float* rt = reinterpret_cast<float*>(wsb + ws.rt_off);
    float* dwp = reinterpret_cast<float*>(wsb + ws.dwp_off);
    int st = MDF_OK;
    if (!staged) {   // (the staged path does this work inside its prep kernel)
        st = run_setup(src_projs, ref_proj, V, B, rt, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, bn_eps,
                       fc_weight, fc_bias, G, dwp, stream);
        if (st != MDF_OK) return st;
    }

    if (!staged) {
        DirectArgs a;
        for (int i = 0; i < MDF_MAX_VIEWS; ++i) a.fea.p[i] = i < N ? features[i] : nullptr;
        a.rt = rt; a.dwp = dwp; a.conv_w = conv_weight; a.hypos = depth_hypos; a.out = cost_volume;
        a.per_pixel = hypos_per_pixel; a.V = V; a.B = B; a.C = C; a.G = G; a.D = D; a.H = H; a.W = W;
        const size_t total = (size_t)B * D * H * W;
        if ((total + 255) / 256 > 0x7fffffffull) return MDF_ERR_UNSUPPORTED;
        cost_volume_direct_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a);
        return launch_status();
    }

    float4* Q4 = reinterpret_cast<float4*>(wsb + ws.q_off);
    float4* S4 = reinterpret_cast<float4*>(wsb + ws.s_off);
    float4* CQ4 = reinterpret_cast<float4*>(wsb + ws.cq_off);
    {
        FeaPtrs fp;
        for (int i = 0; i < MDF_MAX_VIEWS; ++i) fp.p[i] = i < N ? features[i] : nullptr;
        const long long HW = (long long)H * W;
        PrepSetup su;
        for (int v = 0; v < kMaxSrcViews; ++v) su.src_projs.p[v] = v < V ? src_projs[v] : nullptr;
        su.ref_proj = ref_proj; su.V = V; su.rt = rt; su.dwp = dwp;
        su.dw = {conv_weight, bn_weight, bn_bias, bn_mean, bn_var, fc_weight, fc_bias, bn_eps};
        st = launch_setup_and_prep(su, fp, N, B, G, (int)HW, Q4, CQ4, S4, stream);
        if (st != MDF_OK) return st;
    }
    StagedArgs a;
    a.rt = rt; a.dwp = dwp; a.hypos = depth_hypos; a.out = cost_volume; a.vparams = nullptr; a.stats = nullptr;
    a.per_pixel = hypos_per_pixel; a.V = V; a.B = B; a.D = D; a.H = H; a.W = W;
    a.gn = make_grid_norm(H, W);
    a.tiles_x = a.tiles_y = a.slabs = 0;
    StagedBuffers buf;
    buf.S4 = reinterpret_cast<const float*>(S4); buf.Q4 = reinterpret_cast<const float*>(Q4); buf.CQ4 = reinterpret_cast<const float*>(CQ4);
    HotEvents ev;
    ev.start = (cudaEvent_t)hot_start_event; ev.stop = (cudaEvent_t)hot_stop_event;
#ifdef MDF_TUNING
    // 32 + k (+ 256 * rounds per item): variant k of the experimental pipelined kernel; 16 + k: staged variant k
    if ((algo & 255) >= 32) return launch_pipe_variant(G, (algo & 255) - 32, algo >> 8, a, buf, stream, ev);
    if (algo >= 16) return launch_staged_variant(G, algo - 16, a, buf, stream, ev);
#endif
    return launch_staged_default<0>(G, a, buf, stream, ev);
}

int mdf_cost_volume_fwd(const float* const* features, int N, const float* ref_proj, const float* const* src_projs,
                        const float* depth_hypos, int hypos_per_pixel, const float* conv_weight,
                        const float* bn_weight, const float* bn_bias, const float* bn_mean, const float* bn_var,
                        float bn_eps, const float* fc_weight, const float* fc_bias, int B, int C, int G, int D, int H,
                        int W, float* cost_volume, void* workspace, size_t workspace_bytes, mdf_stream_t stream)
{
    return mdf_cost_volume_fwd_ex(features, N, ref_proj, src_projs, depth_hypos, hypos_per_pixel, conv_weight, bn_weight,
                                  bn_bias, bn_mean, bn_var, bn_eps, fc_weight, fc_bias, B, C, G, D, H, W, cost_volume,
                                  workspace, workspace_bytes, 0, nullptr, nullptr, stream);
}

#ifdef MDF_TUNING
// diagnostic: copy the phase timestamps of the last traced launch (StagedCfg::TRACE variants) to the host and reset them
int mdf_debug_read_trace(long long* host_dst, int max_slots)
{
    unsigned n = 0;
    if (cudaMemcpyFromSymbol(&n, g_trace_count, sizeof(n)) != cudaSuccess) return -1;
    if ((int)n > max_slots) n = max_slots;
    if (n > (unsigned)kTraceSlots) n = kTraceSlots;
    if (n && cudaMemcpyFromSymbol(host_dst, g_trace, (size_t)n * kTraceWords * sizeof(long long)) != cudaSuccess) return -1;
    unsigned zero = 0;
    cudaMemcpyToSymbol(g_trace_count, &zero, sizeof(zero));
    return (int)n;
}
#endif

int mdf_debug_sample_positions(const float* rot_trans, const float* depth_hypos, int hypos_per_pixel, int D, int H, int W,
                               float* ix, float* iy, mdf_stream_t stream)
{
    if (D < 0 || H < 0 || W < 0) return MDF_ERR_INVALID_SHAPE;
    const size_t total = (size_t)D * H * W;
    if (total == 0) return MDF_OK;
    if (!rot_trans || !depth_hypos || !ix || !iy) return MDF_ERR_NULL_POINTER;
    const int dev = device_of(ix);
    if (dev < 0) return dev;
    const void* ptrs[] = {rot_trans, depth_hypos, iy};
    const int st = check_on_device(dev, ptrs, 3);
    if (st != MDF_OK) return st;
    DeviceGuard guard(dev);
    sample_positions_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rot_trans, depth_hypos,
                                                                                              hypos_per_pixel, D, H, W, ix, iy);
    return launch_status();
}

size_t mdf_homo_warp_workspace_bytes(int B)
{
    return B <= 0 ? 0 : align_up((size_t)B * 12 * sizeof(float), 256);
}

int mdf_homo_warp_fwd(const float* src_fea, const float* src_proj, const float* ref_proj, const float* depth_hypos,
                      int hypos_per_pixel, int B, int C, int D, int H, int W, float* warped,
                      void* workspace, size_t workspace_bytes, mdf_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || C < 0 || D < 0 || H < 0 || W < 0) return MDF_ERR_INVALID_SHAPE;
    if ((size_t)B * C * D * H * W == 0) return MDF_OK;
    if (!src_fea || !src_proj || !ref_proj || !depth_hypos || !warped) return MDF_ERR_NULL_POINTER;
    if (!workspace || ((uintptr_t)workspace & 255) != 0 || workspace_bytes < mdf_homo_warp_workspace_bytes(B))
        return MDF_ERR_WORKSPACE;
    const int dev = device_of(warped);
    if (dev < 0) return dev;
    const void* ptrs[] = {src_fea, src_proj, ref_proj, depth_hypos, workspace};
    int st = check_on_device(dev, ptrs, 5);
    if (st != MDF_OK) return st;
    DeviceGuard guard(dev);
    float* rt = static_cast<float*>(workspace);
    const float* sp[1] = {src_proj};
    st = run_setup(sp, ref_proj, 1, B, rt, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, nullptr, 0, nullptr, stream);
    if (st != MDF_OK) return st;
    const size_t total = (size_t)B * D * H * W;
    if ((total + 255) / 256 > 0x7fffffffull) return MDF_ERR_UNSUPPORTED;
    homo_warp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src_fea, rt, depth_hypos, hypos_per_pixel, B, C, D, H, W, warped);
    return launch_status();
}

static bool variance_quad_supported(int C) { return C == 16 || C == 32 || C == 64; }

size_t mdf_variance_volume_workspace_bytes(int B, int N, int C, int D, int H, int W)
{
    (void)D;
    if (B <= 0 || N < 2) return 0;
    size_t bytes = align_up((size_t)(N - 1) * B * 12 * sizeof(float), 256);
    // the fast variant's channel-quad copy of the source views
    if (variance_quad_supported(C) && H > 0 && W > 0) bytes += align_up((size_t)(N - 1) * B * C * H * W * sizeof(float), 256);
    return bytes;
}

int mdf_variance_volume_fwd(const float* const* features, int N, const float* ref_proj, const float* const* src_projs,
                            const float* depth_hypos, int hypos_per_pixel, int B, int C, int D, int H, int W,
                            float* cost_volume, void* workspace, size_t workspace_bytes, mdf_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || C <= 0 || D < 0 || H < 0 || W < 0 || N < 2) return MDF_ERR_INVALID_SHAPE;
    if (N > MDF_MAX_VIEWS) return MDF_ERR_UNSUPPORTED;
    if ((size_t)B * D * H * W == 0) return MDF_OK;
    if (!features || !src_projs || !ref_proj || !depth_hypos || !cost_volume) return MDF_ERR_NULL_POINTER;
    const size_t need = mdf_variance_volume_workspace_bytes(B, N, C, D, H, W);
    if (!workspace || ((uintptr_t)workspace & 255) != 0 || workspace_bytes < need) return MDF_ERR_WORKSPACE;
    const int dev = device_of(cost_volume);
    if (dev < 0) return dev;
    {
        const void* ptrs[MDF_MAX_VIEWS * 2 + 4];
        int n = 0;
        for (int i = 0; i < N; ++i) ptrs[n++] = features[i];
        for (int i = 0; i < N - 1; ++i) ptrs[n++] = src_projs[i];
        ptrs[n++] = ref_proj; ptrs[n++] = depth_hypos; ptrs[n++] = workspace;
        int st = check_on_device(dev, ptrs, n);
        if (st != MDF_OK) return st;
    }
    DeviceGuard guard(dev);
    float* rt = static_cast<float*>(workspace);
    int st = run_setup(src_projs, ref_proj, N - 1, B, rt, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, nullptr, 0, nullptr, stream);
    if (st != MDF_OK) return st;
    VarArgs a;
    for (int i = 0; i < MDF_MAX_VIEWS; ++i) a.fea.p[i] = i < N ? features[i] : nullptr;
    a.rt = rt; a.hypos = depth_hypos; a.out = cost_volume;
    a.per_pixel = hypos_per_pixel; a.V = N - 1; a.B = B; a.C = C; a.D = D; a.H = H; a.W = W;
    const size_t total = (size_t)B * D * H * W;
    if ((total + 255) / 256 > 0x7fffffffull) return MDF_ERR_UNSUPPORTED;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    const bool grid_ok = (long long)(N - 1) * B * (C / 4) <= 65535 && (long long)H * W <= INT_MAX - 256 && total < 0x7fffffffull;
    if (variance_quad_supported(C) && grid_ok) {
        float4* P4 = reinterpret_cast<float4*>(static_cast<uint8_t*>(workspace) + align_up((size_t)(N - 1) * B * 12 * sizeof(float), 256));
        const int HWi = H * W;
        planar4_kernel<<<dim3((unsigned)((HWi + 255) / 256), (unsigned)((N - 1) * B * (C / 4))), 256, 0, stream>>>(a.fea, N - 1, B, C, HWi, P4);
        st = launch_status();
        if (st != MDF_OK) return st;
        const int lanes = C / 16, spb = 8 * (32 / lanes);                       // samples per block
        const unsigned qblocks = (unsigned)((total + spb - 1) / spb);
        if (lanes == 1) variance_volume_quad_kernel<1><<<qblocks, 256, 0, stream>>>(a, P4);
        else if (lanes == 2) variance_volume_quad_kernel<2><<<qblocks, 256, 0, stream>>>(a, P4);
        else variance_volume_quad_kernel<4><<<qblocks, 256, 0, stream>>>(a, P4);
        return launch_status();
    }
    if (C <= 16) variance_volume_reg_kernel<16><<<blocks, 256, 0, stream>>>(a);
    else if (C <= 32) variance_volume_reg_kernel<32><<<blocks, 256, 0, stream>>>(a);
    else if (C <= 64) variance_volume_reg_kernel<64><<<blocks, 256, 0, stream>>>(a);
    else variance_volume_kernel<<<blocks, 256, 0, stream>>>(a);
    return launch_status();
}

}  // extern "C"

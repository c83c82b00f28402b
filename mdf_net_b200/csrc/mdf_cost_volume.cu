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
//   prep_kernel             (C/G == 2) source features NCHW -> channels-last "pair difference" maps
//                           S[v][b][y][x][g] = (f[2g+1]-f[2g])*log2(e), reference -> q = tanh((r0-r1)/2).
//                           softmax([a,b]) = [sigmoid(a-b), 1-sigmoid(a-b)] and bilinear sampling is
//                           linear, so gathering the difference map is the same computation with half
//                           the taps and one exp per group.
//   cost_volume_staged      the hot kernel: a CTA owns a tile of reference pixels x a slab of depth
//                           planes; for each source view it finds the bounding box of its samples,
//                           pulls that box of S into shared memory with ONE TMA tile load (hardware
//                           zero fill = grid_sample's zero padding, hardware 128/64/32B swizzle =
//                           conflict-free 128-bit tap reads), then every thread walks its planes:
//                           4 x LDS.128 per 4 groups -> blend -> sigmoid -> similarity -> view weight.
//                           Output stores are 128-byte coalesced rows of the (B,G,D,H,W) volume.
//   cost_volume_direct      any C/G: taps straight from the NCHW features (no staging), two passes.
//   homo_warp / variance    the standalone warp and the (unused by config.py) variance aggregate.
#include <cuda.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "mdf_common.cuh"
#include "mdf_host.cuh"

namespace mdf {

thread_local int g_last_cuda_error = 0;

// ------------------------------------------------------------------------------------------------
// workspace layout (all offsets 256-byte aligned)
// ------------------------------------------------------------------------------------------------
struct Workspace {
    size_t rt_off;    // [V][B][12] float   rot|trans per source view
    size_t dwp_off;   // 64 floats: [0]=alpha [1]=beta' [2]=fc_w [3]=fc_b [4]=beta(raw)   [16..16+G)=conv weight (only G<=32 cached)
    size_t q_off;     // [B][G][H][W] float reference q maps           (staged path)
    size_t s_off;     // [V][B][H][W][G] float source difference maps  (staged path)
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
    w.total = off;
    return w;
}

struct SrcPtrs { const float* p[kMaxSrcViews]; };
struct FeaPtrs { const float* p[MDF_MAX_VIEWS]; };

// ------------------------------------------------------------------------------------------------
// setup: projections + folded depth_weight parameters
// ------------------------------------------------------------------------------------------------
__device__ void compose_proj_f64(const float* __restrict__ src, const float* __restrict__ ref, float* __restrict__ out12)
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

__global__ void setup_kernel(SrcPtrs src_projs, const float* __restrict__ ref_proj, int V, int B,
                             float* __restrict__ rt_all,
                             const float* __restrict__ conv_w, const float* __restrict__ bn_w,
                             const float* __restrict__ bn_b, const float* __restrict__ bn_mean,
                             const float* __restrict__ bn_var, float bn_eps,
                             const float* __restrict__ fc_w, const float* __restrict__ fc_b, int G,
                             float* __restrict__ dwp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V * B) {
        const int v = i / B, b = i % B;
        compose_proj_f64(src_projs.p[v] + 16 * b, ref_proj + 16 * b, rt_all + (size_t)i * 12);
    }
    if (i == 0 && dwp != nullptr) {
        // eval-mode BatchNorm3d(1) folded as ATen applies it: alpha = weight/sqrt(var+eps), beta = bias - mean*alpha
        const float invstd = __frcp_rn(__fsqrt_rn(__fadd_rn(bn_var[0], bn_eps)));
        const float alpha = __fmul_rn(invstd, bn_w[0]);
        const float beta = __fsub_rn(bn_b[0], __fmul_rn(bn_mean[0], alpha));
        float cw_sum = 0.0f;
        for (int g = 0; g < G; ++g) cw_sum += conv_w[g];
        dwp[0] = alpha;
        dwp[1] = beta + alpha * 0.5f * cw_sum;   // staged kernel accumulates sum_g cw_g*(vol_g - 0.5)
        dwp[2] = fc_w[0];
        dwp[3] = fc_b[0];
        dwp[4] = beta;
    }
}

// ------------------------------------------------------------------------------------------------
// prep (C/G == 2): blockDim (32, 8); one block = one image row segment of 32 pixels, one view.
//   view 0 (reference):  Q[b][g][y][x]    = 2*sigmoid(r[2g]-r[2g+1]) - 1          (plane major)
//   view v>0:            S[v-1][b][y][x][g] = (f[2g+1]-f[2g]) * log2(e)             (channels last)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_kernel(FeaPtrs feas, int B, int G, int H, int W, int xtiles, float* __restrict__ Q, float* __restrict__ S)
{
    extern __shared__ float tile[];           // [32][G+1]
    const int lane = threadIdx.x, wy = threadIdx.y;
    int it = blockIdx.x;
    const int xt = it % xtiles; it /= xtiles;
    const int y = it % H; it /= H;
    const int b = it % B;
    const int v = it / B;
    const int x = xt * 32 + lane;
    const size_t HW = (size_t)H * W;
    const float* __restrict__ f = feas.p[v] + (size_t)b * 2 * G * HW + (size_t)y * W;
    if (v == 0) {
        if (x < W)
            for (int g = wy; g < G; g += 8) {
                const float a = __ldg(f + (size_t)(2 * g) * HW + x), c = __ldg(f + (size_t)(2 * g + 1) * HW + x);
                const float e = expf(c - a);                       // exp(-(a-c))
                Q[((size_t)(b * G + g) * H + y) * W + x] = 2.0f / (1.0f + e) - 1.0f;
            }
        return;
    }
    const int ld = G + 1;
    for (int g = wy; g < G; g += 8) {
        float d = 0.0f;
        if (x < W) {
            const float a = __ldg(f + (size_t)(2 * g) * HW + x), c = __ldg(f + (size_t)(2 * g + 1) * HW + x);
            d = (c - a) * kLog2e;
        }
        tile[lane * ld + g] = d;
    }
    __syncthreads();
    const int npx = min(32, W - xt * 32);
    float* __restrict__ dst = S + ((((size_t)(v - 1) * B + b) * H + y) * W + (size_t)xt * 32) * G;
    for (int k = wy * 32 + lane; k < npx * G; k += 256) dst[k] = tile[(k / G) * ld + (k % G)];
}

// ------------------------------------------------------------------------------------------------
// TMA / mbarrier primitives (inline PTX; SASS: UTMALDG, SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// staged kernel configuration
//   G   groups (= channels of the difference map), 32 / 16 / 8 at the three stages
//   PT  depth planes walked by one thread (accumulators PT*G registers)
//   TH  tile height in pixels (tile width is one warp = 32 pixels)
//   PG  plane groups per CTA  -> the CTA's slab is PT*PG planes, blockDim = (32, TH, PG)
//   BW,BH box (pixels) of the source difference map staged per view
// ------------------------------------------------------------------------------------------------
template <int G_, int PT_, int TH_, int PG_, int BW_, int BH_, int MINB_>
struct StagedCfg {
    static constexpr int G = G_, PT = PT_, TH = TH_, PG = PG_, BW = BW_, BH = BH_, MINB = MINB_;
    static constexpr int THREADS = 32 * TH * PG;
    static constexpr int PXB = G * 4;                       // bytes per staged pixel
    static constexpr int BOX_BYTES = BW * BH * PXB;
    static constexpr int SWZ = (G == 32) ? 7 : (G == 16) ? 3 : 1;   // 128B / 64B / 32B swizzle span
    static constexpr int SLAB = PT * PG;
    static constexpr size_t SMEM = BOX_BYTES + 1024 /*align slack*/ + 64;
};

struct StagedArgs {
    const float* S;       // [V][B][H][W][G]
    const float* Q;       // [B][G][H][W]
    const float* rt;      // [V][B][12]
    const float* dwp;     // folded depth_weight
    const float* conv_w;  // (G,)
    const float* hypos;
    float* out;           // (B,G,D,H,W)
    int per_pixel, V, B, D, H, W, tiles_x, tiles_y, slabs;
};

// sigmoid(a-b) for 4 groups of one sample: 4 x LDS.128 from the swizzled box.
template <class Cfg>
__device__ __forceinline__ void taps4(const uint8_t* __restrict__ box, const uint32_t (&A)[4], int j, const Taps& t, float (&p)[4])
{
    const float4 nw = *reinterpret_cast<const float4*>(box + (A[0] ^ (uint32_t)(j << 4)));
    const float4 ne = *reinterpret_cast<const float4*>(box + (A[1] ^ (uint32_t)(j << 4)));
    const float4 sw = *reinterpret_cast<const float4*>(box + (A[2] ^ (uint32_t)(j << 4)));
    const float4 se = *reinterpret_cast<const float4*>(box + (A[3] ^ (uint32_t)(j << 4)));
    p[0] = rcp_approx(1.0f + ex2_approx(blend4(nw.x, ne.x, sw.x, se.x, t)));
    p[1] = rcp_approx(1.0f + ex2_approx(blend4(nw.y, ne.y, sw.y, se.y, t)));
    p[2] = rcp_approx(1.0f + ex2_approx(blend4(nw.z, ne.z, sw.z, se.z, t)));
    p[3] = rcp_approx(1.0f + ex2_approx(blend4(nw.w, ne.w, sw.w, se.w, t)));
}

// Out-of-box sample: same arithmetic from the global difference map with explicit zero padding.
template <int G>
__device__ __noinline__ void taps_global(const float* __restrict__ Sv, int H, int W, const Taps& t, float* __restrict__ p)
{
    const bool x0in = (unsigned)t.x0 < (unsigned)W, x1in = (unsigned)(t.x0 + 1) < (unsigned)W;
    const bool y0in = (unsigned)t.y0 < (unsigned)H, y1in = (unsigned)(t.y0 + 1) < (unsigned)H;
    const float* base = Sv + ((ptrdiff_t)t.y0 * W + t.x0) * G;
    for (int g = 0; g < G; ++g) {
        const float nw = (x0in && y0in) ? __ldg(base + g) : 0.0f;
        const float ne = (x1in && y0in) ? __ldg(base + G + g) : 0.0f;
        const float sw = (x0in && y1in) ? __ldg(base + (ptrdiff_t)W * G + g) : 0.0f;
        const float se = (x1in && y1in) ? __ldg(base + (ptrdiff_t)(W + 1) * G + g) : 0.0f;
        p[g] = rcp_approx(1.0f + ex2_approx(blend4(nw, ne, sw, se, t)));
    }
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
cost_volume_staged_kernel(const __grid_constant__ CUtensorMap tmap, const StagedArgs a)
{
    constexpr int G = Cfg::G, PT = Cfg::PT, TH = Cfg::TH, BW = Cfg::BW, BH = Cfg::BH, PXB = Cfg::PXB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* box = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(box + Cfg::BOX_BYTES);
    int* org = reinterpret_cast<int*>(box + Cfg::BOX_BYTES + 16);   // [2][2]: (x,y) origin slots

    const int lane = threadIdx.x, ty = threadIdx.y, pg = threadIdx.z;
    const int tid = lane + 32 * (ty + TH * pg);

    int it = blockIdx.x;
    const int tile_x = it % a.tiles_x; it /= a.tiles_x;
    const int tile_y = it % a.tiles_y; it /= a.tiles_y;
    const int slab = it % a.slabs;
    const int b = it / a.slabs;

    const int H = a.H, W = a.W, D = a.D;
    const int px = tile_x * 32 + lane, py = tile_y * TH + ty;
    const bool pix_ok = (px < W) && (py < H);
    const int d0 = slab * Cfg::SLAB + pg * PT;
    const size_t HW = (size_t)H * W;
    const GridNorm gn = make_grid_norm(H, W);

    if (tid == 0) {
        mbar_init(bar, 1);
        org[0] = org[1] = org[2] = org[3] = INT_MAX;
        fence_barrier_init();
    }

    // per-thread constants: hypotheses of my planes, cq_g = conv_w[g] * q_g of my pixel
    float depth[PT];
    bool plane_ok[PT];
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        const int d = d0 + i;
        plane_ok[i] = pix_ok && (d < D);
        depth[i] = 0.0f;
        if (plane_ok[i])
            depth[i] = a.per_pixel ? __ldg(a.hypos + ((size_t)b * D + d) * HW + (size_t)py * W + px)
                                   : __ldg(a.hypos + (size_t)b * D + d);
    }
    float cq[G];
    float ksum = 0.0f;
    {
        const float* qp = a.Q + (size_t)b * G * HW + (size_t)py * W + px;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            cq[g] = pix_ok ? __ldg(a.conv_w + g) * __ldg(qp + (size_t)g * HW) : 0.0f;
            ksum += cq[g];
        }
        ksum *= 0.5f;
    }
    const float alpha = __ldg(a.dwp + 0), betap = __ldg(a.dwp + 1), fcw = __ldg(a.dwp + 2), fcb = __ldg(a.dwp + 3);

    float acc[PT][G];
    float wsum[PT];
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        wsum[i] = 0.0f;
#pragma unroll
        for (int g = 0; g < G; ++g) acc[i][g] = 0.0f;
    }
    __syncthreads();

    for (int v = 0; v < a.V; ++v) {
        // ---- 1. sample positions of my planes in view v, CTA-wide bounding-box origin ----
        float rt[12];
        {
            const float* rp = a.rt + ((size_t)v * a.B + b) * 12;
#pragma unroll
            for (int k = 0; k < 12; ++k) rt[k] = __ldg(rp + k);
        }
        const RotXYZ r = rot_xyz(rt, (float)px, (float)py);
        float ix[PT], iy[PT];
        int mnx = INT_MAX, mny = INT_MAX;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            sample_position(r, rt, depth[i], gn, ix[i], iy[i]);
            const bool ok = plane_ok[i] && (ix[i] > -1.0f) && (ix[i] < gn.fw) && (iy[i] > -1.0f) && (iy[i] < gn.fh);
            if (ok) {
                mnx = min(mnx, (int)floorf(ix[i]));
                mny = min(mny, (int)floorf(iy[i]));
            } else {
                ix[i] = -2.0f;   // marks "no tap in bounds" for make_taps below
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        }
        int* slot = org + 2 * (v & 1);
        if (lane == 0 && mnx != INT_MAX) { atomicMin(slot, mnx); atomicMin(slot + 1, mny); }
        if (tid == 0) { int* other = org + 2 * ((v + 1) & 1); other[0] = INT_MAX; other[1] = INT_MAX; }
        __syncthreads();
        int ox = slot[0], oy = slot[1];
        if (ox == INT_MAX) { ox = 0; oy = 0; }

        // ---- 2. one TMA tile load of the [BH][BW][G] box (zero filled outside the image) ----
        if (tid == 0) {
            mbar_expect_tx(bar, Cfg::BOX_BYTES);
            tma_load_4d(box, &tmap, bar, 0, ox, oy, v * a.B + b);
        }
        mbar_wait(bar, (uint32_t)(v & 1));

        // ---- 3. walk my planes ----
        const float* Sv = a.S + ((size_t)v * a.B + b) * HW * G;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            if (!plane_ok[i]) continue;
            const Taps t = make_taps(ix[i], iy[i], gn);
            float p[G];
            if (t.valid) {
                const int rx = t.x0 - ox, ry = t.y0 - oy;
                if ((unsigned)rx < (unsigned)(BW - 1) && (unsigned)ry < (unsigned)(BH - 1)) {
                    const uint32_t o00 = (uint32_t)(ry * BW + rx) * PXB;
                    uint32_t A[4] = {o00, o00 + PXB, o00 + BW * PXB, o00 + (BW + 1) * PXB};
#pragma unroll
                    for (int k = 0; k < 4; ++k) A[k] |= ((A[k] >> 7) & Cfg::SWZ) << 4;
#pragma unroll
                    for (int j = 0; j < G / 4; ++j) {
                        float p4[4];
                        taps4<Cfg>(box, A, j, t, p4);
                        p[4 * j + 0] = p4[0]; p[4 * j + 1] = p4[1]; p[4 * j + 2] = p4[2]; p[4 * j + 3] = p4[3];
                    }
                } else {
                    taps_global<G>(Sv, H, W, t, p);
                }
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) p[g] = 0.5f;   // warped feature = 0 -> softmax = (.5,.5)
            }
            float z = -ksum;
#pragma unroll
            for (int g = 0; g < G; ++g) z = fmaf(cq[g], p[g], z);
            float h = fmaf(z, alpha, betap);              // BatchNorm3d (eval)
            h = fmaxf(h, 0.0f);                           // ReLU
            h = fmaf(h, fcw, fcb);                        // Conv3d(1,1,1)
            const float w = rcp_approx(1.0f + ex2_approx(-kLog2e * h));   // Sigmoid
            wsum[i] += w;
#pragma unroll
            for (int g = 0; g < G; ++g) acc[i][g] = fmaf(w, p[g], acc[i][g]);
        }
        __syncthreads();   // box and origin slot are reused by the next view
    }

    // ---- 4. volume_sum / weight_sum (homoaggregate.py:46), coalesced 128B rows ----
    const float* qp = a.Q + (size_t)b * G * HW + (size_t)py * W + px;
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        if (!plane_ok[i]) continue;
        const float rw = __frcp_rn(wsum[i]);
        float* op = a.out + (((size_t)b * G) * D + (d0 + i)) * HW + (size_t)py * W + px;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float q = __ldg(qp + (size_t)g * HW);
            op[(size_t)g * D * HW] = fmaf(q, fmaf(acc[i][g], rw, -0.5f), 0.5f);
        }
    }
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
// homo_warping (base.py:85-126): one thread per (b, d, y, x), loop over channels.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
homo_warp_kernel(const float* __restrict__ src, const float* __restrict__ rt_all, const float* __restrict__ hypos,
                 int per_pixel, int B, int C, int D, int H, int W, float* __restrict__ out)
{
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)B * D * HW;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int d = (int)((idx / HW) % D);
    const int b = (int)(idx / (HW * D));
    const GridNorm gn = make_grid_norm(H, W);
    const float depth = per_pixel ? __ldg(hypos + ((size_t)b * D + d) * HW + (size_t)y * W + x) : __ldg(hypos + (size_t)b * D + d);
    const float* rt = rt_all + (size_t)b * 12;
    float ix, iy;
    sample_position(rot_xyz(rt, (float)x, (float)y), rt, depth, gn, ix, iy);
    const Taps t = make_taps(ix, iy, gn);
    for (int c = 0; c < C; ++c)
        out[((((size_t)b * C + c) * D + d) * H + y) * W + x] = sample_plane(src + ((size_t)b * C + c) * HW, H, W, t);
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

__global__ void __launch_bounds__(256)
variance_volume_kernel(const VarArgs a)
{
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
    float vmax[kMaxSrcViews], vinv[kMaxSrcViews];
    for (int v = 0; v < a.V; ++v) {
        const float* rt = a.rt + ((size_t)v * a.B + b) * 12;
        float ix, iy;
        sample_position(rot_xyz(rt, (float)x, (float)y), rt, depth, gn, ix, iy);
        const Taps t = make_taps(ix, iy, gn);
        const float* src = a.fea.p[v + 1] + (size_t)b * a.C * HW;
        float m = -INFINITY;
        for (int c = 0; c < a.C; ++c) m = fmaxf(m, sample_plane(src + c * HW, a.H, a.W, t));
        float s = 0.0f;
        for (int c = 0; c < a.C; ++c) s += expf(sample_plane(src + c * HW, a.H, a.W, t) - m);
        vmax[v] = m;
        vinv[v] = s;
    }
    const float nviews = (float)(a.V + 1);
    const float* ref = a.fea.p[0] + (size_t)b * a.C * HW + (size_t)y * a.W + x;
    for (int c = 0; c < a.C; ++c) {
        const float rv = __ldg(ref + c * HW);
        float s1 = rv, s2 = rv * rv;
        for (int v = 0; v < a.V; ++v) {
            const float* rt = a.rt + ((size_t)v * a.B + b) * 12;
            float ix, iy;
            sample_position(rot_xyz(rt, (float)x, (float)y), rt, depth, gn, ix, iy);
            const Taps t = make_taps(ix, iy, gn);
            const float* src = a.fea.p[v + 1] + (size_t)b * a.C * HW;
            const float p = expf(sample_plane(src + c * HW, a.H, a.W, t) - vmax[v]) / vinv[v];
            s1 += p;
            s2 = fmaf(p, p, s2);
        }
        const float mean = s1 / nviews;
        a.out[((((size_t)b * a.C + c) * a.D + d) * a.H + y) * a.W + x] = s2 / nviews - mean * mean;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    // resolved through the runtime: the library does not link libcuda
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

template <class Cfg>
static int launch_staged(const StagedArgs& args, cudaStream_t stream)
{
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) return MDF_ERR_UNSUPPORTED;
    CUtensorMap tmap;
    const cuuint64_t dims[4] = {(cuuint64_t)Cfg::G, (cuuint64_t)args.W, (cuuint64_t)args.H, (cuuint64_t)args.V * args.B};
    const cuuint64_t strides[3] = {(cuuint64_t)Cfg::PXB, (cuuint64_t)args.W * Cfg::PXB, (cuuint64_t)args.H * args.W * Cfg::PXB};
    const cuuint32_t box[4] = {(cuuint32_t)Cfg::G, (cuuint32_t)Cfg::BW, (cuuint32_t)Cfg::BH, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUtensorMapSwizzle swz = Cfg::G == 32 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : Cfg::G == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(args.S), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return MDF_ERR_CUDA; }
    auto kern = cost_volume_staged_kernel<Cfg>;
    MDF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    const long long items = (long long)args.tiles_x * args.tiles_y * args.slabs * args.B;
    if (items <= 0) return MDF_OK;
    if (items > INT_MAX) return MDF_ERR_UNSUPPORTED;
    kern<<<(unsigned)items, dim3(32, Cfg::TH, Cfg::PG), Cfg::SMEM, stream>>>(tmap, args);
    return launch_status();
}

//                    G  PT TH PG  BW  BH MINB
using CfgG32 = StagedCfg<32, 1, 4, 2, 48, 8, 2>;    // 256 thr, box 48 KiB
using CfgG16 = StagedCfg<16, 4, 4, 2, 48, 8, 2>;    // 256 thr, box 24 KiB, slab 8 planes
using CfgG8  = StagedCfg<8, 8, 8, 1, 48, 12, 2>;    // 256 thr, box 18 KiB, slab 8 planes

static int run_setup(const float* const* src_projs, const float* ref_proj, int V, int B, float* rt,
                     const float* conv_w, const float* bn_w, const float* bn_b, const float* bn_mean,
                     const float* bn_var, float bn_eps, const float* fc_w, const float* fc_b, int G, float* dwp,
                     cudaStream_t stream)
{
    SrcPtrs sp;
    for (int v = 0; v < kMaxSrcViews; ++v) sp.p[v] = v < V ? src_projs[v] : nullptr;
    const int n = V * B;
    setup_kernel<<<(n + 63) / 64, 64, 0, stream>>>(sp, ref_proj, V, B, rt, conv_w, bn_w, bn_b, bn_mean, bn_var, bn_eps,
                                                   fc_w, fc_b, G, dwp);
    return launch_status();
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
                           mdf_stream_t stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || C <= 0 || G <= 0 || D < 0 || H < 0 || W < 0 || N < 2 || C % G != 0) return MDF_ERR_INVALID_SHAPE;
    if (N > MDF_MAX_VIEWS || C / G > 16) return MDF_ERR_UNSUPPORTED;
    if ((size_t)B * D * H * W == 0) return MDF_OK;   // empty volume
    if (!features || !src_projs || !ref_proj || !depth_hypos || !conv_weight || !bn_weight || !bn_bias || !bn_mean ||
        !bn_var || !fc_weight || !fc_bias || !cost_volume)
        return MDF_ERR_NULL_POINTER;
    const int V = N - 1;
    const bool staged = staged_supported(C, G) && algo != 2;
    if (algo == 1 && !staged) return MDF_ERR_UNSUPPORTED;
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
    int st = run_setup(src_projs, ref_proj, V, B, rt, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, bn_eps,
                       fc_weight, fc_bias, G, dwp, stream);
    if (st != MDF_OK) return st;

    if (!staged) {
        DirectArgs a;
        for (int i = 0; i < MDF_MAX_VIEWS; ++i) a.fea.p[i] = i < N ? features[i] : nullptr;
        a.rt = rt; a.dwp = dwp; a.conv_w = conv_weight; a.hypos = depth_hypos; a.out = cost_volume;
        a.per_pixel = hypos_per_pixel; a.V = V; a.B = B; a.C = C; a.G = G; a.D = D; a.H = H; a.W = W;
        const size_t total = (size_t)B * D * H * W;
        cost_volume_direct_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a);
        return launch_status();
    }

    float* Q = reinterpret_cast<float*>(wsb + ws.q_off);
    float* S = reinterpret_cast<float*>(wsb + ws.s_off);
    {
        FeaPtrs fp;
        for (int i = 0; i < MDF_MAX_VIEWS; ++i) fp.p[i] = i < N ? features[i] : nullptr;
        const int xtiles = (W + 31) / 32;
        const long long blocks = (long long)N * B * H * xtiles;
        if (blocks > INT_MAX) return MDF_ERR_UNSUPPORTED;
        prep_kernel<<<(unsigned)blocks, dim3(32, 8), 32 * (G + 1) * sizeof(float), stream>>>(fp, B, G, H, W, xtiles, Q, S);
        st = launch_status();
        if (st != MDF_OK) return st;
    }
    StagedArgs a;
    a.S = S; a.Q = Q; a.rt = rt; a.dwp = dwp; a.conv_w = conv_weight; a.hypos = depth_hypos; a.out = cost_volume;
    a.per_pixel = hypos_per_pixel; a.V = V; a.B = B; a.D = D; a.H = H; a.W = W;
    a.tiles_x = (W + 31) / 32;
    if (G == 32) {
        a.tiles_y = (H + CfgG32::TH - 1) / CfgG32::TH; a.slabs = (D + CfgG32::SLAB - 1) / CfgG32::SLAB;
        return launch_staged<CfgG32>(a, stream);
    } else if (G == 16) {
        a.tiles_y = (H + CfgG16::TH - 1) / CfgG16::TH; a.slabs = (D + CfgG16::SLAB - 1) / CfgG16::SLAB;
        return launch_staged<CfgG16>(a, stream);
    }
    a.tiles_y = (H + CfgG8::TH - 1) / CfgG8::TH; a.slabs = (D + CfgG8::SLAB - 1) / CfgG8::SLAB;
    return launch_staged<CfgG8>(a, stream);
}

int mdf_cost_volume_fwd(const float* const* features, int N, const float* ref_proj, const float* const* src_projs,
                        const float* depth_hypos, int hypos_per_pixel, const float* conv_weight,
                        const float* bn_weight, const float* bn_bias, const float* bn_mean, const float* bn_var,
                        float bn_eps, const float* fc_weight, const float* fc_bias, int B, int C, int G, int D, int H,
                        int W, float* cost_volume, void* workspace, size_t workspace_bytes, mdf_stream_t stream)
{
    return mdf_cost_volume_fwd_ex(features, N, ref_proj, src_projs, depth_hypos, hypos_per_pixel, conv_weight, bn_weight,
                                  bn_bias, bn_mean, bn_var, bn_eps, fc_weight, fc_bias, B, C, G, D, H, W, cost_volume,
                                  workspace, workspace_bytes, 0, stream);
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
    homo_warp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src_fea, rt, depth_hypos, hypos_per_pixel, B, C, D, H, W, warped);
    return launch_status();
}

size_t mdf_variance_volume_workspace_bytes(int B, int N, int C, int D, int H, int W)
{
    (void)C; (void)D; (void)H; (void)W;
    if (B <= 0 || N < 2) return 0;
    return align_up((size_t)(N - 1) * B * 12 * sizeof(float), 256);
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
    variance_volume_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a);
    return launch_status();
}

}  // extern "C"

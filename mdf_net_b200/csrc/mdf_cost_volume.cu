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
//   prep_kernel             (C/G == 2) source features NCHW -> "pair difference" maps in planar-float4
//                           layout S4[v][b][j][y][x] = (f[2g+1]-f[2g])*log2(e) for g = 4j..4j+3, reference
//                           -> q = tanh((r0-r1)/2).  softmax([a,b]) = [sigmoid(a-b), 1-sigmoid(a-b)] and
//                           bilinear sampling is linear, so gathering the difference map is the same
//                           computation with half the taps and one exp per group.
//   cost_volume_staged      the hot kernel: a CTA owns a tile of reference pixels x a slab of depth
//                           planes; for each source view it finds the bounding box of its samples,
//                           pulls that [G/4][BH][BW] float4 box of S4 into shared memory with ONE TMA
//                           tile load (hardware zero fill = grid_sample's zero padding; neighbouring
//                           lanes read neighbouring 16-byte texels = conflict-free LDS.128 with
//                           immediate offsets), then every thread walks its planes:
//                           4 x LDS.128 per 4 groups -> blend -> sigmoid -> similarity -> view weight.
//                           Samples outside the box (rough depth maps) are served by further staging
//                           rounds; there is no slow global-memory path.
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
    size_t dwp_off;   // 64 floats: [0]=alpha [1]=beta' [2]=fc_w [3]=fc_b [4]=beta(raw) [5]=weight of an out-of-image sample
    size_t q_off;     // [B][G/4][H][W] float4 reference q maps           (staged path)
    size_t s_off;     // [V][B][G/4][H][W] float4 source difference maps  (staged path)
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
        // weight of a (sample, view) pair whose taps all fall outside the source image: every similarity is 0.5
        const float hv = fmaf(fmaxf(dwp[1], 0.0f), fc_w[0], fc_b[0]);
        dwp[5] = 1.0f / (1.0f + expf(-hv));
    }
}

// ------------------------------------------------------------------------------------------------
// prep (C/G == 2): one thread per pixel of one view; blockIdx.y = view * B + b.
//   "planar float4" layout: plane j holds groups 4j..4j+3 of every pixel as one float4
//   view 0 (reference):  Q4[b][j][y][x]      = 2*sigmoid(r[2g]-r[2g+1]) - 1,  g = 4j..4j+3
//   view v>0:            S4[v-1][b][j][y][x] = (f[2g+1]-f[2g]) * log2(e)
// Loads are 128-byte coalesced rows of the NCHW planes, stores are 512-byte coalesced float4 rows.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_kernel(FeaPtrs feas, int B, int G, int HW, float4* __restrict__ Q4, float4* __restrict__ S4)
{
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const int v = blockIdx.y / B, b = blockIdx.y % B;
    const float* __restrict__ f = feas.p[v] + (size_t)b * 2 * G * HW + pix;
    const int J = G / 4;
    if (v == 0) {
        float4* __restrict__ dst = Q4 + (size_t)b * J * HW + pix;
        for (int j = 0; j < J; ++j) {
            float d[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a = __ldg(f + (size_t)(8 * j + 2 * k) * HW), c = __ldg(f + (size_t)(8 * j + 2 * k + 1) * HW);
                d[k] = 2.0f / (1.0f + expf(c - a)) - 1.0f;
            }
            dst[(size_t)j * HW] = make_float4(d[0], d[1], d[2], d[3]);
        }
        return;
    }
    float4* __restrict__ dst = S4 + ((size_t)(v - 1) * B + b) * J * HW + pix;
    for (int j = 0; j < J; ++j) {
        float d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = __ldg(f + (size_t)(8 * j + 2 * k) * HW), c = __ldg(f + (size_t)(8 * j + 2 * k + 1) * HW);
            d[k] = (c - a) * kLog2e;
        }
        dst[(size_t)j * HW] = make_float4(d[0], d[1], d[2], d[3]);
    }
}

// ------------------------------------------------------------------------------------------------
// TMA / mbarrier / shared-memory primitives (inline PTX; SASS: UTMALDG, SYNCS, LDS.128)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// ptxas folds `addr + constant` into the immediate offset of LDS.128
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// ------------------------------------------------------------------------------------------------
// staged kernel configuration
//   G      groups (channels of the difference map): 32 / 16 / 8 at the three stages
//   PT     depth planes walked by one thread (accumulators: PT*G registers)
//   TH     tile height in pixels (tile width is one warp = 32 pixels: 128-byte output rows)
//   PG     plane groups per CTA -> the CTA's slab is PT*PG planes, blockDim = (32, TH, PG)
//   BW,BH  box (texels) of one source difference map staged per TMA load: [G/4][BH][BW] float4
// ------------------------------------------------------------------------------------------------
template <int G_, int PT_, int TH_, int PG_, int BW_, int BH_, int MINB_>
struct StagedCfg {
    static constexpr int G = G_, PT = PT_, TH = TH_, PG = PG_, BW = BW_, BH = BH_, MINB = MINB_;
    static constexpr int J = G / 4;
    static constexpr int THREADS = 32 * TH * PG;
    static constexpr int PLANE_BYTES = BW * BH * 16;
    static constexpr int BOX_BYTES = J * PLANE_BYTES;
    static constexpr int SLAB = PT * PG;
    static constexpr size_t SMEM = BOX_BYTES + 128 /*align slack*/ + 64;
    static_assert(BW * 4 <= 256 && BH <= 256, "TMA box dimensions are limited to 256 elements");
};

struct StagedArgs {
    const float4* Q4;     // [B][G/4][H][W]
    const float* rt;      // [V][B][12]
    const float* dwp;     // folded depth_weight
    const float* conv_w;  // (G,)
    const float* hypos;
    float* out;           // (B,G,D,H,W)
    int per_pixel, V, B, D, H, W, tiles_x, tiles_y, slabs;
};

constexpr int kNone = INT_MAX;

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
cost_volume_staged_kernel(const __grid_constant__ CUtensorMap tmap, const StagedArgs a)
{
    constexpr int G = Cfg::G, J = Cfg::J, PT = Cfg::PT, TH = Cfg::TH, BW = Cfg::BW, BH = Cfg::BH;
    constexpr int PLANE = Cfg::PLANE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t box = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t bar = box + Cfg::BOX_BYTES;
    // red[slot][0] = min x0, red[slot][1] = min y0 of the samples still to be staged
    int* red = reinterpret_cast<int*>(smem_raw + (box - smem_u32(smem_raw)) + Cfg::BOX_BYTES + 16);

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
    const GridNormFast gf = make_grid_norm_fast(H, W);
    const GridNorm& gn = gf.g;

    if (tid == 0) {
        mbar_init(bar, 1);
        red[0] = red[1] = red[2] = red[3] = kNone;
        fence_barrier_init();
    }

    // per-thread constants: hypotheses of my planes, cq_g = conv_w[g] * q_g of my pixel
    float depth[PT];
    uint32_t ok_mask = 0;                       // bit i: plane d0+i exists and my pixel is inside the image
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        const int d = d0 + i;
        depth[i] = 0.0f;
        if (pix_ok && d < D) {
            ok_mask |= 1u << i;
            depth[i] = a.per_pixel ? __ldg(a.hypos + ((size_t)b * D + d) * HW + (size_t)py * W + px)
                                   : __ldg(a.hypos + (size_t)b * D + d);
        }
    }
    const float4* __restrict__ qp = a.Q4 + (size_t)b * J * HW + (size_t)py * W + px;
    float cq[G];
    float ksum = 0.0f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const float4 q = pix_ok ? __ldg(qp + (size_t)j * HW) : make_float4(0.f, 0.f, 0.f, 0.f);
        cq[4 * j + 0] = __ldg(a.conv_w + 4 * j + 0) * q.x;
        cq[4 * j + 1] = __ldg(a.conv_w + 4 * j + 1) * q.y;
        cq[4 * j + 2] = __ldg(a.conv_w + 4 * j + 2) * q.z;
        cq[4 * j + 3] = __ldg(a.conv_w + 4 * j + 3) * q.w;
        ksum += (cq[4 * j + 0] + cq[4 * j + 1]) + (cq[4 * j + 2] + cq[4 * j + 3]);
    }
    ksum *= 0.5f;
    const float alpha = __ldg(a.dwp + 0), betap = __ldg(a.dwp + 1), fcw = __ldg(a.dwp + 2), fcb = __ldg(a.dwp + 3);
    const float w_void = __ldg(a.dwp + 5);       // view weight of a sample with no tap in bounds (similarity 0.5)

    float acc[PT][G];
    float wsum[PT];
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        wsum[i] = 0.0f;
#pragma unroll
        for (int g = 0; g < G; ++g) acc[i][g] = 0.0f;
    }
    uint64_t n_void = 0;                         // 8 bits per plane: views whose sample fell outside the source image
    uint32_t iter = 0;                           // CTA-uniform count of staging rounds (mbarrier phase, reduction slot)
    __syncthreads();

    for (int v = 0; v < a.V; ++v) {
        // ---- 1. sample positions of my planes in view v ----
        float rt[12];
        {
            const float* rp = a.rt + ((size_t)v * a.B + b) * 12;
#pragma unroll
            for (int k = 0; k < 12; ++k) rt[k] = __ldg(rp + k);
        }
        const RotXYZ r = rot_xyz(rt, (float)px, (float)py);
        float ix[PT], iy[PT];
        uint32_t todo = 0;                       // bit i: sample of plane i still has to be gathered
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            sample_position_fast(r, rt, depth[i], gf, ix[i], iy[i]);
            const bool inside = (ix[i] > -1.0f) && (ix[i] < gn.fw) && (iy[i] > -1.0f) && (iy[i] < gn.fh);
            if ((ok_mask >> i) & 1u) {
                if (inside) todo |= 1u << i;
                else n_void += 1ull << (8 * i);
            }
        }

        // ---- 2. staging rounds: box origin = min corner of the samples still to do ----
        bool first = true;
        while (true) {
            int* slot = red + 2 * (iter & 1u);
            // order-preserving integer keys of the (non-negative part of the) positions: floor once, after the min
            int kx = kNone, ky = kNone;
#pragma unroll
            for (int i = 0; i < PT; ++i)
                if ((todo >> i) & 1u) {
                    kx = min(kx, ix[i] < 0.0f ? -1 : __float_as_int(ix[i]));
                    if (first) ky = min(ky, iy[i] < 0.0f ? -1 : __float_as_int(iy[i]));
                }
            kx = __reduce_min_sync(0xffffffffu, kx);
            if (first) ky = __reduce_min_sync(0xffffffffu, ky);
            if (lane == 0 && kx != kNone) {
                atomicMin(slot, kx < 0 ? -1 : (int)__int_as_float(kx));
                if (first) atomicMin(slot + 1, ky < 0 ? -1 : (int)__int_as_float(ky));
            }
            if (tid == 0) { int* other = red + 2 * ((iter + 1u) & 1u); other[0] = kNone; other[1] = kNone; }
            __syncthreads();
            const int ox = slot[0];
            if (ox == kNone) {                   // CTA-uniform: nothing to gather in this view
                __syncthreads();                 // everybody has read the slot before the next view writes it
                break;
            }
            if (!first) {
                // later rounds: y origin over the samples whose column fits, so that at least one sample
                // (the topmost of them) lands inside the box and the loop always makes progress
                const float xlim = (float)(ox + BW - 1);
                int m = kNone;
#pragma unroll
                for (int i = 0; i < PT; ++i)
                    if (((todo >> i) & 1u) && ix[i] < xlim) m = min(m, iy[i] < 0.0f ? -1 : __float_as_int(iy[i]));
                m = __reduce_min_sync(0xffffffffu, m);
                if (lane == 0 && m != kNone) atomicMin(slot + 1, m < 0 ? -1 : (int)__int_as_float(m));
                __syncthreads();
            }
            const int oy = slot[1];

            // one TMA tile load of the [G/4][BH][BW] float4 box (zero filled outside the image)
            if (tid == 0) {
                mbar_expect_tx(bar, Cfg::BOX_BYTES);
                tma_load_3d(box, &tmap, bar, ox * 4, oy, (v * a.B + b) * J);
            }
            mbar_wait(bar, iter & 1u);

            // ---- 3. gather the samples that landed in the box ----
#pragma unroll
            for (int i = 0; i < PT; ++i) {
                if (!((todo >> i) & 1u)) continue;
                float fx0, fy0;
                int x0, y0;
                floor_small(ix[i], fx0, x0);
                floor_small(iy[i], fy0, y0);
                const int rx = x0 - ox, ry = y0 - oy;
                if ((unsigned)rx >= (unsigned)(BW - 1) || (unsigned)ry >= (unsigned)(BH - 1)) continue;   // next round
                todo &= ~(1u << i);
                const float ax = __fsub_rn(__fadd_rn(fx0, 1.0f), ix[i]), bx = __fsub_rn(ix[i], fx0);
                const float ay = __fsub_rn(__fadd_rn(fy0, 1.0f), iy[i]), by = __fsub_rn(iy[i], fy0);
                Taps t;
                t.wnw = __fmul_rn(ax, ay); t.wne = __fmul_rn(bx, ay); t.wsw = __fmul_rn(ax, by); t.wse = __fmul_rn(bx, by);
                const uint32_t addr = box + (uint32_t)(ry * BW + rx) * 16u;
                float p[G];
                float z = -ksum;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float4 nw = lds128(addr + j * PLANE);                 // constant offsets -> LDS.128 [R + imm]
                    const float4 ne = lds128(addr + j * PLANE + 16);
                    const float4 sw = lds128(addr + j * PLANE + BW * 16);
                    const float4 se = lds128(addr + j * PLANE + BW * 16 + 16);
                    p[4 * j + 0] = rcp_approx(1.0f + ex2_approx(blend4(nw.x, ne.x, sw.x, se.x, t)));
                    p[4 * j + 1] = rcp_approx(1.0f + ex2_approx(blend4(nw.y, ne.y, sw.y, se.y, t)));
                    p[4 * j + 2] = rcp_approx(1.0f + ex2_approx(blend4(nw.z, ne.z, sw.z, se.z, t)));
                    p[4 * j + 3] = rcp_approx(1.0f + ex2_approx(blend4(nw.w, ne.w, sw.w, se.w, t)));
                    z = fmaf(cq[4 * j + 0], p[4 * j + 0], z);
                    z = fmaf(cq[4 * j + 1], p[4 * j + 1], z);
                    z = fmaf(cq[4 * j + 2], p[4 * j + 2], z);
                    z = fmaf(cq[4 * j + 3], p[4 * j + 3], z);
                }
                float h = fmaf(z, alpha, betap);              // BatchNorm3d (eval)
                h = fmaxf(h, 0.0f);                           // ReLU
                h = fmaf(h, fcw, fcb);                        // Conv3d(1,1,1)
                const float w = rcp_approx(1.0f + ex2_approx(-kLog2e * h));   // Sigmoid
                wsum[i] += w;
#pragma unroll
                for (int g = 0; g < G; ++g) acc[i][g] = fmaf(w, p[g], acc[i][g]);
            }
            ++iter;
            first = false;
            // the box and the reduction slot are reused: everybody must be done reading; also learn
            // whether any sample is still waiting for another box
            if (!__syncthreads_or(todo != 0u)) break;
        }
    }

    // ---- 4. volume_sum / weight_sum (homoaggregate.py:46), coalesced 128-byte rows ----
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        if (!((ok_mask >> i) & 1u)) continue;
        const float nv = (float)((unsigned)(n_void >> (8 * i)) & 255u);
        const float ws = fmaf(nv, w_void, wsum[i]);
        const float half_void = 0.5f * nv * w_void;          // void samples: similarity 0.5 in every group
        const float rw = __frcp_rn(ws);
        float* op = a.out + (((size_t)b * G) * D + (d0 + i)) * HW + (size_t)py * W + px;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 q = __ldg(qp + (size_t)j * HW);
            op[(size_t)(4 * j + 0) * D * HW] = fmaf(q.x, fmaf(acc[i][4 * j + 0] + half_void, rw, -0.5f), 0.5f);
            op[(size_t)(4 * j + 1) * D * HW] = fmaf(q.y, fmaf(acc[i][4 * j + 1] + half_void, rw, -0.5f), 0.5f);
            op[(size_t)(4 * j + 2) * D * HW] = fmaf(q.z, fmaf(acc[i][4 * j + 2] + half_void, rw, -0.5f), 0.5f);
            op[(size_t)(4 * j + 3) * D * HW] = fmaf(q.w, fmaf(acc[i][4 * j + 3] + half_void, rw, -0.5f), 0.5f);
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
static int launch_staged(const StagedArgs& args, const float* S4, cudaStream_t stream)
{
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) return MDF_ERR_UNSUPPORTED;
    // 3-D view of S4[(v*B+b)*J + j][y][x] (float4 texels): dim0 = 4*W floats, dim1 = H rows, dim2 = planes
    CUtensorMap tmap;
    const cuuint64_t planes = (cuuint64_t)args.V * args.B * Cfg::J;
    const cuuint64_t dims[3] = {(cuuint64_t)args.W * 4, (cuuint64_t)args.H, planes};
    const cuuint64_t strides[2] = {(cuuint64_t)args.W * 16, (cuuint64_t)args.H * args.W * 16};
    const cuuint32_t box[3] = {(cuuint32_t)Cfg::BW * 4, (cuuint32_t)Cfg::BH, (cuuint32_t)Cfg::J};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(S4), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return MDF_ERR_CUDA; }
    auto kern = cost_volume_staged_kernel<Cfg>;
    MDF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    StagedArgs a = args;
    a.tiles_x = (a.W + 31) / 32;
    a.tiles_y = (a.H + Cfg::TH - 1) / Cfg::TH;
    a.slabs = (a.D + Cfg::SLAB - 1) / Cfg::SLAB;
    const long long items = (long long)a.tiles_x * a.tiles_y * a.slabs * a.B;
    if (items <= 0) return MDF_OK;
    if (items > INT_MAX) return MDF_ERR_UNSUPPORTED;
    kern<<<(unsigned)items, dim3(32, Cfg::TH, Cfg::PG), Cfg::SMEM, stream>>>(tmap, a);
    return launch_status();
}

//                    G  PT TH PG  BW  BH MINB
using CfgG32 = StagedCfg<32, 1, 4, 2, 40, 8, 2>;    // 256 thr, box 40 KiB, slab 2 planes
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

    float4* Q4 = reinterpret_cast<float4*>(wsb + ws.q_off);
    float4* S4 = reinterpret_cast<float4*>(wsb + ws.s_off);
    {
        FeaPtrs fp;
        for (int i = 0; i < MDF_MAX_VIEWS; ++i) fp.p[i] = i < N ? features[i] : nullptr;
        const long long HW = (long long)H * W;
        if (HW > INT_MAX - 256 || (long long)N * B > 65535) return MDF_ERR_UNSUPPORTED;
        prep_kernel<<<dim3((unsigned)((HW + 255) / 256), (unsigned)(N * B)), 256, 0, stream>>>(fp, B, G, (int)HW, Q4, S4);
        st = launch_status();
        if (st != MDF_OK) return st;
    }
    StagedArgs a;
    a.Q4 = Q4; a.rt = rt; a.dwp = dwp; a.conv_w = conv_weight; a.hypos = depth_hypos; a.out = cost_volume;
    a.per_pixel = hypos_per_pixel; a.V = V; a.B = B; a.D = D; a.H = H; a.W = W;
    a.tiles_x = a.tiles_y = a.slabs = 0;
    const float* S = reinterpret_cast<const float*>(S4);
    if (G == 32) return launch_staged<CfgG32>(a, S, stream);
    if (G == 16) return launch_staged<CfgG16>(a, S, stream);
    return launch_staged<CfgG8>(a, S, stream);
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

// mdf_staged.cuh -- the hot kernel of the plane-sweep cost volume (C/G == 2) and its layout pass.
//
//   prep_kernel             (overlaps the 1-block setup kernel of mdf_setup.cuh: programmatic dependent launch)
//                           source features NCHW -> "pair difference" maps in planar-float4 layout
//                           S4[v][b][j][y][x] = (f[2g+1]-f[2g])*log2(e) for g = 4j..4j+3, and the reference
//                           view -> q = 2*sigmoid(r[2g]-r[2g+1]) - 1.  softmax([a,b]) = [sigmoid(a-b),
//                           1-sigmoid(a-b)] and bilinear sampling is linear, so gathering the difference
//                           map is the reference's computation with half the taps and one exp per group:
//                               similarity_g = 0.5 + q_g * (sigmoid(warp(S)_g) - 0.5)
//   cost_volume_staged      a CTA owns a tile of reference pixels x a slab of depth planes.  For each
//                           source view it finds the bounding box of its samples, pulls that
//                           [G/4][BH][BW] float4 box of S4 into shared memory with ONE TMA tile load
//                           (hardware zero fill = grid_sample's zero padding; neighbouring lanes read
//                           neighbouring 16-byte texels = conflict-free LDS.128 with immediate offsets),
//                           then every thread walks its planes: 4 x LDS.128 per 4 groups -> blend ->
//                           sigmoid -> similarity -> learned view weight -> weighted mean.  Samples that
//                           fall outside the box (rough depth maps, silhouettes) are served by further
//                           staging rounds; there is no slow global-memory path.  Output stores are
//                           128-byte coalesced rows of the (B,G,D,H,W) volume.
//
// Reference: net/unit/base.py:85-126 (homo_warping), net/unit/homoaggregate.py:25-46 (+16-20).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "mdf_common.cuh"
#include "mdf_host.cuh"
#include "mdf_setup.cuh"
#include "mdf_tma.cuh"

namespace mdf {

struct FeaPtrs { const float* p[MDF_MAX_VIEWS]; };

// ------------------------------------------------------------------------------------------------
// prep: one thread per pixel of one view; blockIdx.y = view * B + b.
// Loads are 128-byte coalesced rows of the NCHW planes, stores are 512-byte coalesced float4 rows.
// ------------------------------------------------------------------------------------------------
struct PrepSetup {          // the per-call scalar work (mdf_setup.cuh) of the launch that precedes the layout pass
    SrcPtrs src_projs;
    const float* ref_proj;
    int V;
    float* rt;
    float* dwp;
    DepthWeightPtrs dw;
};

// Exit path of every block: wait for the setup kernel this launch overlaps with (programmatic dependent
// launch), so that "prep has completed" implies "setup has completed" for whatever follows in the stream.
__device__ __forceinline__ void prep_exit() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One thread handles PX consecutive pixels of one view (PX = 4 when the planes are 16-byte aligned: LDG.128 per
// channel plane, four float4 texels written back to back; PX = 1 otherwise).
template <int PX>
static __global__ void __launch_bounds__(256)
prep_kernel(FeaPtrs feas, int B, int G, int HW, const float* __restrict__ conv_w,
            float4* __restrict__ Q4, float4* __restrict__ CQ4, float4* __restrict__ S4)
{
    const int pix = (blockIdx.x * blockDim.x + threadIdx.x) * PX;
    if (pix >= HW) { prep_exit(); return; }
    const int v = blockIdx.y / B, b = blockIdx.y % B;
    const float* __restrict__ f = feas.p[v] + (size_t)b * 2 * G * HW + pix;
    const int J = G / 4;
    // channel pair (a, c) = (f[2g], f[2g+1]) of PX pixels
    auto load_pair = [&](int g, float (&a)[PX], float (&c)[PX]) {
        if (PX == 4) {
            const float4 va = __ldg(reinterpret_cast<const float4*>(f + (size_t)(2 * g) * HW));
            const float4 vc = __ldg(reinterpret_cast<const float4*>(f + (size_t)(2 * g + 1) * HW));
            a[0] = va.x; a[PX > 1 ? 1 : 0] = va.y; a[PX > 2 ? 2 : 0] = va.z; a[PX > 3 ? 3 : 0] = va.w;
            c[0] = vc.x; c[PX > 1 ? 1 : 0] = vc.y; c[PX > 2 ? 2 : 0] = vc.z; c[PX > 3 ? 3 : 0] = vc.w;
        } else {
            a[0] = __ldg(f + (size_t)(2 * g) * HW);
            c[0] = __ldg(f + (size_t)(2 * g + 1) * HW);
        }
    };
    if (v == 0) {
        // reference view: q_g = 2*sigmoid(r[2g]-r[2g+1]) - 1, cq_g = conv_w[g]*q_g
        float4* __restrict__ qd = Q4 + (size_t)b * J * HW + pix;
        float4* __restrict__ cd = CQ4 + (size_t)b * J * HW + pix;
        for (int j = blockIdx.z; j < J; j += gridDim.z) {
            float d[4][PX];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a[PX], c[PX];
                load_pair(4 * j + k, a, c);
#pragma unroll
                for (int x = 0; x < PX; ++x) d[k][x] = 2.0f / (1.0f + expf(c[x] - a[x])) - 1.0f;
            }
            const float w0 = __ldg(conv_w + 4 * j), w1 = __ldg(conv_w + 4 * j + 1), w2 = __ldg(conv_w + 4 * j + 2), w3 = __ldg(conv_w + 4 * j + 3);
#pragma unroll
            for (int x = 0; x < PX; ++x) {
                qd[(size_t)j * HW + x] = make_float4(d[0][x], d[1][x], d[2][x], d[3][x]);
                cd[(size_t)j * HW + x] = make_float4(w0 * d[0][x], w1 * d[1][x], w2 * d[2][x], w3 * d[3][x]);
            }
        }
        prep_exit();
        return;
    }
    // source views: blockIdx.z strides over the float4 planes (more blocks in flight at the coarse stage)
    float4* __restrict__ dst = S4 + ((size_t)(v - 1) * B + b) * J * HW + pix;
    for (int j = blockIdx.z; j < J; j += gridDim.z) {
        float d[4][PX];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a[PX], c[PX];
            load_pair(4 * j + k, a, c);
#pragma unroll
            for (int x = 0; x < PX; ++x) d[k][x] = (c[x] - a[x]) * kLog2e;
        }
#pragma unroll
        for (int x = 0; x < PX; ++x) dst[(size_t)j * HW + x] = make_float4(d[0][x], d[1][x], d[2][x], d[3][x]);
    }
    prep_exit();
}

static inline unsigned prep_zsplit(int HW, int images, int G);

// setup (1 block: float64 projections + folded depth_weight) and the layout pass, overlapped by programmatic
// dependent launch: the layout pass reads nothing the setup kernel writes, so it starts immediately.
static int launch_setup_and_prep(const PrepSetup& su, const FeaPtrs& fp, int N, int B, int G, int HW,
                                 float4* Q4, float4* CQ4, float4* S4, cudaStream_t stream)
{
    const int n = su.V * B;
    setup_kernel<<<(n + 63) / 64, 64, 0, stream>>>(su.src_projs, su.ref_proj, su.V, B, su.rt, su.dw, G, su.dwp);
    int st = launch_status();
    if (st != MDF_OK) return st;
    // 4 pixels per thread when every feature plane is 16-byte aligned
    bool vec = (HW % 4) == 0;
    for (int i = 0; i < N; ++i) vec = vec && (reinterpret_cast<uintptr_t>(fp.p[i]) & 15) == 0;
    const int px = vec ? 4 : 1;
    const int threads_x = (HW + px - 1) / px;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((threads_x + 255) / 256), (unsigned)(N * B), prep_zsplit(threads_x, N * B, G));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (vec) MDF_CUDA_TRY(cudaLaunchKernelEx(&cfg, prep_kernel<4>, fp, B, G, HW, su.dw.conv_w, Q4, CQ4, S4));
    else MDF_CUDA_TRY(cudaLaunchKernelEx(&cfg, prep_kernel<1>, fp, B, G, HW, su.dw.conv_w, Q4, CQ4, S4));
    return MDF_OK;
}

// planes per source-view block column: split until the grid holds a few waves of 148 SMs x 8 blocks
static inline unsigned prep_zsplit(int threads_x, int images, int G)
{
    const long long blocks = (long long)((threads_x + 255) / 256) * images;
    unsigned z = 1;
    while (z < (unsigned)(G / 4) && blocks * z < 148LL * 8 * 4) z *= 2;
    return z;
}

// ------------------------------------------------------------------------------------------------
// configuration
//   G      groups (channels of the difference map): 32 / 16 / 8 at the three stages
//   PT     depth planes walked by one thread (accumulators: PT*G registers)
//   TH     tile height in warps (a warp covers WX x WY pixels, 32 x 1 by default: 128-byte output rows; tile = WX x TH*WY pixels)
//   PG     plane groups per CTA -> the CTA's slab is PT*PG planes, blockDim = (32, TH, PG)
//   BW,BH  box (texels) of one source difference map staged per TMA load: [G/4][BH][BW] float4
//   MINB   CTAs per SM the register allocation aims at
//   CQS    keep the per-pixel similarity weights cq_g = conv_w[g]*q_g in shared memory instead of registers
// ------------------------------------------------------------------------------------------------
template <int G_, int PT_, int TH_, int PG_, int BW_, int BH_, int MINB_, bool CQS_, bool EARLY_ = true, bool BF_ = false, bool RCP2_ = false, bool TRACE_ = false, int ABL_ = 0, int WX_ = 32>
struct StagedCfg {
    static constexpr int G = G_, PT = PT_, TH = TH_, PG = PG_, BW = BW_, BH = BH_, MINB = MINB_;
    static constexpr bool RCP2 = RCP2_, TRACE = TRACE_;
    // a warp covers WX x WY reference pixels (WX * WY = 32): 32 x 1 rows by default; 8 x 4 / 16 x 2 where the map width is not a
    // multiple of 32 (W = 200 at stage 0 of DTU fills 6.25 warps of 32: one lane in nine idles), see pick_warp_width()
    static constexpr int WX = WX_, WY = 32 / WX_;
    static_assert(WX_ == 32 || WX_ == 16 || WX_ == 8, "warp footprint");
    static constexpr int ABL = ABL_;   // ablation study (wrong results): 1 no MUFU, 2 no tap loads, 4 no stores, 8 cheap positions, 16 no cq loads
    static constexpr bool CQS = CQS_, EARLY = EARLY_, BF = BF_;   // BF: branch-free sample bodies (samples of a thread interleave)
    static constexpr int J = G / 4;
    static constexpr int NCQ = CQS ? 1 : G;
    static constexpr int THREADS = 32 * TH * PG;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int PLANE_BYTES = BW * BH * 16;
    static constexpr int BOX_BYTES = J * PLANE_BYTES;          // multiple of 128
    static constexpr int TILE_BYTES = J * TH * 32 * 16;        // per-pixel float4 planes of the CTA's tile (q, cq)
    static constexpr int RT_BYTES = kMaxSrcViews * 12 * 4;
    static constexpr int SLAB = PT * PG;
    static constexpr int OFF_Q = 2 * BOX_BYTES;                // two boxes: the gather of view v overlaps the load of v+1
    static constexpr int OFF_CQ = OFF_Q + TILE_BYTES;
    static constexpr int OFF_RT = OFF_CQ + TILE_BYTES;
    static constexpr int OFF_BAR = (OFF_RT + RT_BYTES + 15) / 16 * 16;   // 4 mbarriers: box 0, box 1, retry loads, tiles
    static constexpr int OFF_CTL = OFF_BAR + 32;                          // control words, see enum below
    static constexpr size_t SMEM = OFF_CTL + 4 * (16 + 2 * kMaxSrcViews) + 128 /*alignment slack*/;
    static_assert(BW * 2 <= 256 && BH <= 256, "TMA box dimensions are limited to 256 elements");
    static_assert(PT <= 8, "plane bookkeeping uses 8 bits per plane");
};

struct StagedArgs {
    const float* rt;      // [V][B][12]
    const float* dwp;     // folded depth_weight
    const float* vparams; // MODE 2: [V][4] per-view BatchNorm fold alpha_v, beta'_v, weight of an out-of-image sample, -
    double* stats;        // MODE 1: [V][2] sum z, sum z^2 of the kernel's z (= true z - 0.5*sum(conv_w)) over (B,D,H,W)
    const float* hypos;
    float* out;           // (B,G,D,H,W)
    // MODE 3 (backward, phase 1 on the staged gather): upstream gradient and saved forward output (B,G,D,H,W) in;
    // za [V][3][B*D*H*W]: slot 3v <- A'_v = sum_g gout_g sim_vg, slot 3v+1 <- the kernel's z_v;  go [B*D*H*W] <- sum_g gout_g out_g
    const float* gout = nullptr;
    const float* fwd_out = nullptr;
    float* za = nullptr;
    float* go = nullptr;
    GridNorm gn;
    int per_pixel, V, B, D, H, W, tiles_x, tiles_y, slabs;
};

struct StagedMaps {
    CUtensorMap s4;       // source difference maps  [V*B*J][H][W] float4
    CUtensorMap q4;       // reference q              [B*J][H][W]   float4
    CUtensorMap cq4;      // conv_w * q               [B*J][H][W]   float4
};

constexpr int kNone = INT_MAX;

// phase timestamps of a sample of CTAs (tuning builds only: StagedCfg::TRACE)
constexpr int kTraceSlots = 4096, kTraceWords = 32;
#ifdef MDF_TUNING
__device__ long long g_trace[kTraceSlots * kTraceWords];
__device__ unsigned g_trace_count;
#endif

// order-preserving integer key of a sample coordinate in (-1, size): negative -> -1, else its bit pattern
__device__ __forceinline__ int coord_key(float v) { return v < 0.0f ? -1 : __float_as_int(v); }
__device__ __forceinline__ int key_floor(int k) { return k < 0 ? -1 : (int)__int_as_float(k); }

// control words in shared memory.  The bounding-box words exist once per view parity: view v+1 is
// prepared while slower warps may still be preparing view v.
enum { kMinX = 0, kMinY = 1, kCount = 2, kSetStride = 4, kRetry = 8 /* [2][2] */, kLeft = 12, kOrg = 16 /* [V][2] */ };

// MODE 0: eval (one BatchNorm fold for all views).  The training path (mdf_backward.cu) runs the same kernel twice:
// MODE 1 gathers and only accumulates the batch statistics of z per source view (train-mode BatchNorm3d normalises each
// view's z over (B,D,H,W), homoaggregate.py:40 + base.py:50-68), MODE 2 is the forward with per-view folds.
// MODE 3 is phase 1 of the backward on the same gather: the accumulator registers hold gout_g * q_g instead, and every
// (sample, view) leaves its z_v and A'_v = sum_g gout_g sim_vg behind for the per-element pass and the sweep of mdf_backward.cu.
template <class Cfg, int MODE = 0>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
cost_volume_staged_kernel(const __grid_constant__ StagedMaps maps, const StagedArgs a)
{
    constexpr int G = Cfg::G, J = Cfg::J, PT = Cfg::PT, TH = Cfg::TH, BW = Cfg::BW, BH = Cfg::BH, NCQ = Cfg::NCQ;
    constexpr int PLANE = Cfg::PLANE_BYTES, THREADS = Cfg::THREADS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t pad = (128u - (smem_u32(smem_raw) & 127u)) & 127u;
    const uint32_t box0 = smem_u32(smem_raw) + pad;
    const uint32_t bar0 = box0 + Cfg::OFF_BAR;           // +0: box 0, +8: box 1, +16: retry loads, +24: q / cq tiles
    const uint32_t rt_s = box0 + Cfg::OFF_RT;
    volatile int* ctl = reinterpret_cast<volatile int*>(smem_raw + pad + Cfg::OFF_CTL);
    int* ctl_nv = reinterpret_cast<int*>(smem_raw + pad + Cfg::OFF_CTL);

    const int lane = threadIdx.x, ty = threadIdx.y, pg = threadIdx.z;
    const int tid = lane + 32 * (ty + TH * pg);
    long long tr[kTraceWords];
    int trn = 0;
    const bool tracing = Cfg::TRACE && lane == 0 && (blockIdx.x % 8) == 3;
    auto stamp = [&]() { if (Cfg::TRACE) { if (trn < kTraceWords) tr[trn] = clock64(); ++trn; } };
    stamp();                                     // 0: start
    const uint32_t q_s = box0 + Cfg::OFF_Q + (uint32_t)(ty * 32 + lane) * 16u;      // [j][ty][lane] float4
    const uint32_t cq_s = box0 + Cfg::OFF_CQ + (uint32_t)(ty * 32 + lane) * 16u;
    constexpr uint32_t TJ = TH * 32 * 16;                                           // plane stride of the tiles

    int it = blockIdx.x;
    const int tile_x = it % a.tiles_x; it /= a.tiles_x;
    const int tile_y = it % a.tiles_y; it /= a.tiles_y;
    const int slab = it % a.slabs;
    const int b = it / a.slabs;

    const int H = a.H, W = a.W, D = a.D;
    const int px = tile_x * Cfg::WX + (lane % Cfg::WX), py = (tile_y * TH + ty) * Cfg::WY + lane / Cfg::WX;
    const bool pix_ok = (px < W) && (py < H);
    const int d0 = slab * Cfg::SLAB + pg * PT;
    const size_t HW = (size_t)H * W;
    GridNormFast gf;
    gf.g = a.gn;
    gf.r_half_wm1 = refine_rcp(a.gn.half_wm1);
    gf.r_half_hm1 = refine_rcp(a.gn.half_hm1);
    const bool tiny_map = !(a.gn.half_wm1 > 0.25f && a.gn.half_hm1 > 0.25f);   // W or H == 1: the normalisation divides by 0

    if (tid == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(bar0 + 16, 1); mbar_init(bar0 + 24, 1);
        ctl[kMinX] = kNone; ctl[kMinY] = kNone; ctl[kCount] = 0;
        ctl[kSetStride + kMinX] = kNone; ctl[kSetStride + kMinY] = kNone; ctl[kSetStride + kCount] = 0;
        ctl[kRetry + 0] = ctl[kRetry + 1] = ctl[kRetry + 2] = ctl[kRetry + 3] = kNone;
        ctl[kLeft] = 0;
        fence_barrier_init();
        // q and cq = conv_w * q of the tile's pixels: two TMA tile loads (zero fill beyond the image)
        mbar_expect_tx(bar0 + 24, 2 * Cfg::TILE_BYTES);
        // (the tile lands as [j][TH*WY rows][WX] float4: row-major over the tile = (ty * 32 + lane), whatever the warp footprint)
        tma_load_3d(box0 + Cfg::OFF_Q, &maps.q4, bar0 + 24, tile_x * Cfg::WX * 2, tile_y * TH * Cfg::WY, b * J);
        tma_load_3d(box0 + Cfg::OFF_CQ, &maps.cq4, bar0 + 24, tile_x * Cfg::WX * 2, tile_y * TH * Cfg::WY, b * J);
    }
    // rot | trans of every source view of this batch item -> shared memory
    for (int k = tid; k < a.V * 12; k += THREADS) {
        const float val = __ldg(a.rt + ((size_t)(k / 12) * a.B + b) * 12 + (k % 12));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(rt_s + 4u * k), "f"(val) : "memory");
    }

    // per-thread constants: hypotheses of my planes
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
    float alpha = __ldg(a.dwp + 0), betap = __ldg(a.dwp + 1);          // MODE 2: reloaded per view
    const float fcw = __ldg(a.dwp + 2), fcb = __ldg(a.dwp + 3);
    float w_void_view = 0.0f;                    // MODE 2: weight of an out-of-image sample of the view being prepared
    float wvoid[MODE == 2 ? PT : 1];             // MODE 2: sum over the views of those weights, per plane
    float zs1 = 0.0f, zs2 = 0.0f;                // MODE 1: sum z, sum z^2 of the (at most PT) samples this thread gathered for one view
    float a_base[MODE == 3 ? PT : 1], a_void[MODE == 3 ? PT : 1];   // MODE 3: A' = a_base + sum_g (gout_g q_g) p_g; A' of an out-of-image sample
    const size_t total = (size_t)a.B * D * HW;
    const size_t e0 = ((size_t)b * D + d0) * HW + (size_t)py * W + px;      // element of plane d0 (+ i * HW)
    // MODE 1: one row of sums per warp (plain read-modify-write by lane 0: a shared-memory double atomicAdd is a CAS spin loop)
    __shared__ double stats_s[MODE == 1 ? Cfg::WARPS : 1][MODE == 1 ? 2 * kMaxSrcViews : 2];
    double* stats_w = stats_s[MODE == 1 ? ty + TH * pg : 0];
    if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < PT; ++i) wvoid[i] = 0.0f;
    }
    if (MODE == 1)
        for (int k = tid; k < Cfg::WARPS * 2 * kMaxSrcViews; k += THREADS) (&stats_s[0][0])[k] = 0.0;
    // per-view BatchNorm fold (MODE 2) and, after a gather, the hand-over of this warp's statistics (MODE 1)
    auto set_view = [&](int v) {
        if (MODE == 2) { alpha = __ldg(a.vparams + 4 * v); betap = __ldg(a.vparams + 4 * v + 1); }
    };
    auto set_void_view = [&](int v) {
        if (MODE == 2) w_void_view = __ldg(a.vparams + 4 * v + 2);
    };
    auto flush_stats = [&](int v) {
        if (MODE == 1) {
            // a warp's handful of samples (at most 32 * PT) in float, everything beyond that (the CTA, the grid) in double
            float f1 = zs1, f2 = zs2;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { f1 += __shfl_xor_sync(0xffffffffu, f1, o); f2 += __shfl_xor_sync(0xffffffffu, f2, o); }
            if (lane == 0) { stats_w[2 * v] += (double)f1; stats_w[2 * v + 1] += (double)f2; }
            zs1 = 0.0f; zs2 = 0.0f;
        }
    };

    float2 acc[PT][G / 2];                       // pairs of groups: the core runs on packed FFMA2 / FMUL2 / FADD2
    float wsum[PT];
#pragma unroll
    for (int i = 0; i < PT; ++i) {
        wsum[i] = 0.0f;
#pragma unroll
        for (int g = 0; g < G / 2; ++g) acc[i][g] = make_float2(0.0f, 0.0f);
    }
    uint64_t n_void = 0;                         // 8 bits per plane: views whose sample fell outside the source image
    uint32_t left = 0;                           // bit v: some sample of mine did not fit view v's box
    stamp();                                     // 1: hypotheses requested, before the first barrier
    __syncthreads();                             // rt_s, mbarriers and control words are set up
    stamp();                                     // 2: after the first barrier

    // rot | trans of view v from shared memory (issued early, consumed by `positions`)
    auto load_rt = [&](int v, float (&rt)[12]) {
#pragma unroll
        for (int k = 0; k < 12; ++k) rt[k] = lds32(rt_s + (uint32_t)(v * 12 + k) * 4u);
    };
    // sample positions of my planes; returns the mask of samples with at least one tap in bounds
    auto positions = [&](int v, const float (&rt)[12], float (&ix)[PT], float (&iy)[PT], bool count_void) -> uint32_t {
        uint32_t todo = 0;
        const RotXYZ r = rot_xyz(rt, (float)px, (float)py);
        // all planes through the branch-free fast chain first (PT independent dependency chains in one basic
        // block); the IEEE chain only if some lane of the warp met an operand the fast divisions do not cover
        bool exact = !tiny_map;
        if (Cfg::ABL & 8) {
#pragma unroll
            for (int i = 0; i < PT; ++i) { ix[i] = fmaf(depth[i], 0.002f, (float)px); iy[i] = fmaf(r.x, 1e-9f, (float)py + 0.25f); }
        } else {
#pragma unroll
        for (int i = 0; i < PT; ++i) exact = sample_position_try(r, rt, depth[i], gf, ix[i], iy[i]) && exact;
        }
        if (__any_sync(0xffffffffu, !exact)) {
#pragma unroll
            for (int i = 0; i < PT; ++i) sample_position(r, rt, depth[i], gf.g, ix[i], iy[i]);
        }
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            const bool inside = (ix[i] > -1.0f) && (ix[i] < a.gn.fw) && (iy[i] > -1.0f) && (iy[i] < a.gn.fh);
            if ((ok_mask >> i) & 1u) {
                if (inside) todo |= 1u << i;
                else if (count_void) {
                    if (MODE == 2) wvoid[i] += w_void_view;
                    else if (MODE == 3) {
                        // every similarity is 0.5: z (shifted) = 0, A' = 0.5 * sum_g gout_g
                        a.za[(size_t)(3 * v) * total + e0 + (size_t)i * HW] = a_void[MODE == 3 ? i : 0];
                        a.za[(size_t)(3 * v + 1) * total + e0 + (size_t)i * HW] = 0.0f;
                    } else n_void += 1ull << (8 * i);
                }
            }
        }
        return todo;
    };

    // Contribute this warp's corner to the CTA-wide bounding box of view v; the LAST warp to arrive issues
    // the TMA load of the box into buffer v & 1.  No barrier: nobody waits here.  (All warps of the CTA
    // have finished gathering view v-2 from that buffer when the last of them arrives.)
    auto announce = [&](int v, const float (&ix)[PT], const float (&iy)[PT], uint32_t todo) {
        int kx = kNone, ky = kNone;
#pragma unroll
        for (int i = 0; i < PT; ++i)
            if ((todo >> i) & 1u) { kx = min(kx, coord_key(ix[i])); ky = min(ky, coord_key(iy[i])); }
        kx = __reduce_min_sync(0xffffffffu, kx);
        ky = __reduce_min_sync(0xffffffffu, ky);
        if (lane == 0) {
            const int set = kSetStride * (v & 1);
            if (kx != kNone) { atomicMin(ctl_nv + set + kMinX, key_floor(kx)); atomicMin(ctl_nv + set + kMinY, key_floor(ky)); }
            // the mins must be performed before the arrival becomes visible (reductions and atomics with a
            // return value are NOT ordered with each other without it: measured, not assumed)
            __threadfence_block();
            if (atomicAdd(ctl_nv + set + kCount, 1) == Cfg::WARPS - 1) {
                __threadfence_block();
                int ox = ctl[set + kMinX], oy = ctl[set + kMinY];
                if (ox == kNone) { ox = 0; oy = 0; }          // no sample of the whole CTA lands in this view
                ctl[set + kMinX] = kNone; ctl[set + kMinY] = kNone; ctl[set + kCount] = 0;
                ctl[kOrg + 2 * v] = ox; ctl[kOrg + 2 * v + 1] = oy;
                const uint32_t bar = bar0 + 8u * (v & 1);
                mbar_expect_tx(bar, Cfg::BOX_BYTES);
                // the tensor maps count in 8-byte elements: 2 per texel
                tma_load_3d(box0 + (uint32_t)(v & 1) * Cfg::BOX_BYTES, &maps.s4, bar, ox * 2, oy, (v * a.B + b) * J);
            }
        }
    };

    float cq[NCQ];
    float ksum = 0.0f;
    auto ex2_approx = [](float x) -> float { return (Cfg::ABL & 1) ? fmaf(x, 0.5f, 1.0f) : mdf::ex2_approx(x); };
    auto rcp_approx = [](float x) -> float { return (Cfg::ABL & 1) ? x * 0.25f : mdf::rcp_approx(x); };
    // gather the samples of `todo` that lie inside the box with origin (ox, oy); returns the rest
    auto gather = [&](int v, uint32_t box, int ox, int oy, const float (&ix)[PT], const float (&iy)[PT], uint32_t todo) -> uint32_t {
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            const bool act = (todo >> i) & 1u;
            if (!Cfg::BF && !act) continue;
            const float sx = (Cfg::BF && !act) ? 0.0f : ix[i], sy = (Cfg::BF && !act) ? 0.0f : iy[i];
            float fx0, fy0;
            int x0, y0;
            floor_small(sx, fx0, x0);
            floor_small(sy, fy0, y0);
            int rx = x0 - ox, ry = y0 - oy;
            const bool inb = act && (unsigned)rx < (unsigned)(BW - 1) && (unsigned)ry < (unsigned)(BH - 1);
            if (!Cfg::BF && !inb) continue;                       // not in this box
            if (Cfg::BF && !inb) { rx = 0; ry = 0; }              // a harmless cell of the box; the sample's weight is zeroed below
            if (inb) todo &= ~(1u << i);
            const float ax = __fsub_rn(__fadd_rn(fx0, 1.0f), sx), bx = __fsub_rn(sx, fx0);
            const float ay = __fsub_rn(__fadd_rn(fy0, 1.0f), sy), by = __fsub_rn(sy, fy0);
            const float wnw = __fmul_rn(ax, ay), wne = __fmul_rn(bx, ay), wsw = __fmul_rn(ax, by), wse = __fmul_rn(bx, by);
            const float2 Wnw = make_float2(wnw, wnw), Wne = make_float2(wne, wne);
            const float2 Wsw = make_float2(wsw, wsw), Wse = make_float2(wse, wse);
            const uint32_t addr = box + (uint32_t)(ry * BW + rx) * 16u;
            float2 p[G / 2];
            float2 z2 = make_float2(-ksum, 0.0f);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                float4 nw, ne, sw, se;
                if (Cfg::ABL & 2) {
                    const float fa = __uint_as_float(addr + j);
                    nw = make_float4(ax, bx, ay, fa); ne = make_float4(bx, ay, fa, ax); sw = make_float4(ay, fa, ax, bx); se = make_float4(fa, by, bx, ay);
                } else {
                    nw = lds128(addr + j * PLANE);                 // constant offsets -> LDS.128 [R + imm]
                    ne = lds128(addr + j * PLANE + 16);
                    sw = lds128(addr + j * PLANE + BW * 16);
                    se = lds128(addr + j * PLANE + BW * 16 + 16);
                }
                float4 c;
                if (Cfg::CQS && (Cfg::ABL & 16)) c = make_float4(ax, bx, ay, by);
                else if (Cfg::CQS) c = lds128(cq_s + (uint32_t)j * TJ);
                else c = make_float4(cq[(4 * j + 0) % NCQ], cq[(4 * j + 1) % NCQ], cq[(4 * j + 2) % NCQ], cq[(4 * j + 3) % NCQ]);
                // bilinear blend, two groups per instruction; per component the reference's order
                // fma(se,wse, fma(sw,wsw, fma(ne,wne, nw*wnw)))
                float2 t01 = __fmul2_rn(make_float2(nw.x, nw.y), Wnw), t23 = __fmul2_rn(make_float2(nw.z, nw.w), Wnw);
                t01 = __ffma2_rn(make_float2(ne.x, ne.y), Wne, t01); t23 = __ffma2_rn(make_float2(ne.z, ne.w), Wne, t23);
                t01 = __ffma2_rn(make_float2(sw.x, sw.y), Wsw, t01); t23 = __ffma2_rn(make_float2(sw.z, sw.w), Wsw, t23);
                t01 = __ffma2_rn(make_float2(se.x, se.y), Wse, t01); t23 = __ffma2_rn(make_float2(se.z, se.w), Wse, t23);
                // sigmoid(a-b) = 1 / (1 + 2^t).  t is capped so that (1+2^t0)(1+2^t1) cannot become inf * 0;
                // beyond the cap the similarity is below 1e-18 either way.  One MUFU.RCP serves two groups:
                // r = 1/(u0*u1), 1/u0 = r*u1, 1/u1 = r*u0.
                const float2 one2 = make_float2(1.0f, 1.0f);
                float2 p01, p23;
                if (Cfg::RCP2) {
                    // one MUFU.RCP per group: 2^t = inf gives 1/inf = 0, no cap and no cross multiplications (the XU pipe has
                    // room: it is a third busy, the issue slots are what the kernel runs out of)
                    const float2 u01 = __fadd2_rn(make_float2(ex2_approx(t01.x), ex2_approx(t01.y)), one2);
                    const float2 u23 = __fadd2_rn(make_float2(ex2_approx(t23.x), ex2_approx(t23.y)), one2);
                    p01 = make_float2(rcp_approx(u01.x), rcp_approx(u01.y));
                    p23 = make_float2(rcp_approx(u23.x), rcp_approx(u23.y));
                } else {
                    const float2 u01 = __fadd2_rn(make_float2(ex2_approx(fminf(t01.x, 62.0f)), ex2_approx(fminf(t01.y, 62.0f))), one2);
                    const float2 u23 = __fadd2_rn(make_float2(ex2_approx(fminf(t23.x, 62.0f)), ex2_approx(fminf(t23.y, 62.0f))), one2);
                    const float r01 = rcp_approx(u01.x * u01.y), r23 = rcp_approx(u23.x * u23.y);
                    p01 = __fmul2_rn(make_float2(r01, r01), make_float2(u01.y, u01.x));
                    p23 = __fmul2_rn(make_float2(r23, r23), make_float2(u23.y, u23.x));
                }
                p[2 * j] = p01; p[2 * j + 1] = p23;
                z2 = __ffma2_rn(make_float2(c.x, c.y), p01, z2);
                z2 = __ffma2_rn(make_float2(c.z, c.w), p23, z2);
            }
            const float z = (Cfg::BF && !inb) ? 0.0f : z2.x + z2.y;
            if (MODE == 1) { zs1 += z; zs2 = fmaf(z, z, zs2); continue; }     // statistics pass: z is all it wants
            if (MODE == 3) {
                float2 ap2 = make_float2(a_base[MODE == 3 ? i : 0], 0.0f);
#pragma unroll
                for (int g = 0; g < G / 2; ++g) ap2 = __ffma2_rn(acc[i][g], p[g], ap2);
                a.za[(size_t)(3 * v) * total + e0 + (size_t)i * HW] = ap2.x + ap2.y;
                a.za[(size_t)(3 * v + 1) * total + e0 + (size_t)i * HW] = z;
                continue;
            }
            float h = fmaf(z, alpha, betap);              // BatchNorm3d (eval fold, or this view's batch statistics)
            h = fmaxf(h, 0.0f);                           // ReLU
            h = fmaf(h, fcw, fcb);                        // Conv3d(1,1,1)
            float w = rcp_approx(1.0f + ex2_approx(-kLog2e * h));   // Sigmoid
            if (Cfg::BF && !inb) w = 0.0f;
            wsum[i] += w;
            const float2 w2 = make_float2(w, w);
#pragma unroll
            for (int g = 0; g < G / 2; ++g) acc[i][g] = __ffma2_rn(w2, p[g], acc[i][g]);
        }
        return todo;
    };

    if (MODE == 3) {
        // upstream gradient of my planes: acc <- gout_g * q_g, go = sum_g gout_g out_g (homoaggregate.py:46 backward),
        // a_void = 0.5 sum_g gout_g, a_base = a_void - 0.5 sum_g gout_g q_g   (sim = 0.5 + q (p - 0.5))
        mbar_wait(bar0 + 24, 0);                 // q tile
        const size_t gstride = (size_t)D * HW;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            a_base[MODE == 3 ? i : 0] = 0.0f; a_void[MODE == 3 ? i : 0] = 0.0f;
            if (!((ok_mask >> i) & 1u)) continue;
            const size_t o = (((size_t)b * G) * D + (d0 + i)) * HW + (size_t)py * W + px;
            float go = 0.0f, sg = 0.0f, sgq = 0.0f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float4 q = lds128(q_s + (uint32_t)j * TJ);
                const float* gp = a.gout + o + (size_t)(4 * j) * gstride;
                const float* fp = a.fwd_out + o + (size_t)(4 * j) * gstride;
                const float g0 = __ldg(gp), g1 = __ldg(gp + gstride), g2 = __ldg(gp + 2 * gstride), g3 = __ldg(gp + 3 * gstride);
                go = fmaf(g0, __ldg(fp), go); go = fmaf(g1, __ldg(fp + gstride), go);
                go = fmaf(g2, __ldg(fp + 2 * gstride), go); go = fmaf(g3, __ldg(fp + 3 * gstride), go);
                acc[i][2 * j] = make_float2(g0 * q.x, g1 * q.y);
                acc[i][2 * j + 1] = make_float2(g2 * q.z, g3 * q.w);
                sg += (g0 + g1) + (g2 + g3);
                sgq += (acc[i][2 * j].x + acc[i][2 * j].y) + (acc[i][2 * j + 1].x + acc[i][2 * j + 1].y);
            }
            a_void[MODE == 3 ? i : 0] = 0.5f * sg;
            a_base[MODE == 3 ? i : 0] = 0.5f * (sg - sgq);
            a.go[e0 + (size_t)i * HW] = go;
        }
    }

    float ix[PT], iy[PT];
    uint32_t todo;
    {
        float rt[12];
        load_rt(0, rt);
        set_void_view(0);
        todo = positions(0, rt, ix, iy, true);
    }
    stamp();                                     // 3: positions of view 0 done (the hypotheses have arrived)
    announce(0, ix, iy, todo);
    stamp();                                     // 4: announced
    mbar_wait(bar0 + 24, 0);                     // q / cq tiles have landed
    stamp();                                     // 5: tiles landed
    {   // ksum = 0.5 * sum_g cq_g (z accumulates sum_g cq_g (p_g - 0.5)); cq stays in registers unless CQS
        float s4 = 0.0f;
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 c = lds128(cq_s + (uint32_t)j * TJ);
            if (!Cfg::CQS) { cq[(4 * j + 0) % NCQ] = c.x; cq[(4 * j + 1) % NCQ] = c.y; cq[(4 * j + 2) % NCQ] = c.z; cq[(4 * j + 3) % NCQ] = c.w; }
            s4 += (c.x + c.y) + (c.z + c.w);
        }
        ksum = 0.5f * s4;
    }

    // ---- main loop: no block-wide barrier; warps run freely, the box of view v+1 streams in while view v
    //      is gathered (its load is issued before anybody waits for view v) ----
    for (int v = 0; v < a.V; ++v) {
        float nx[PT], ny[PT];
        uint32_t ntodo = 0;
        if (v + 1 < a.V) {
            float rt[12];
            load_rt(v + 1, rt);
            set_void_view(v + 1);
            ntodo = positions(v + 1, rt, nx, ny, true);
            announce(v + 1, nx, ny, ntodo);
        }
        stamp();                                 // 6 + 3v: next view prepared, waiting for this view's box
        mbar_wait(bar0 + 8u * (v & 1), (uint32_t)(v >> 1) & 1u);
        stamp();                                 // 7 + 3v: box landed
        const int ox = ctl[kOrg + 2 * v], oy = ctl[kOrg + 2 * v + 1];
        set_view(v);
        todo = gather(v, box0 + (uint32_t)(v & 1) * Cfg::BOX_BYTES, ox, oy, ix, iy, todo);
        flush_stats(v);
        stamp();                                 // 8 + 3v: gathered
        if (todo != 0u) left |= 1u << v;
#pragma unroll
        for (int i = 0; i < PT; ++i) { ix[i] = nx[i]; iy[i] = ny[i]; }
        todo = ntodo;
    }

    // ---- volume_sum / weight_sum (homoaggregate.py:46), coalesced 128-byte rows ----
    auto epilogue = [&]() {
        const float w_void = __ldg(a.dwp + 5);       // view weight of a sample with no tap in bounds (similarity 0.5)
    #pragma unroll
        for (int i = 0; i < PT; ++i) {
            if (!((ok_mask >> i) & 1u)) continue;
            const float nv = (float)((unsigned)(n_void >> (8 * i)) & 255u);
            // weight of this plane's out-of-image samples: count * one weight (eval), or the sum of the views' own weights
            const float wv_sum = MODE == 2 ? wvoid[MODE == 2 ? i : 0] : nv * w_void;
            const float ws = MODE == 2 ? wv_sum + wsum[i] : fmaf(nv, w_void, wsum[i]);
            const float rw = __frcp_rn(ws);
            // out = 0.5 + q * ((acc + 0.5*wv_sum) / ws - 0.5); void samples have similarity 0.5 in every group
            const float c0 = fmaf(0.5f * wv_sum, rw, -0.5f);
            const float2 rw2 = make_float2(rw, rw), c02 = make_float2(c0, c0), half2 = make_float2(0.5f, 0.5f);
            float* op = a.out + (((size_t)b * G) * D + (d0 + i)) * HW + (size_t)py * W + px;
            const size_t gstride = (size_t)D * HW;
    #pragma unroll
            for (int j = 0; j < J; ++j) {
                const float4 q = lds128(q_s + (uint32_t)j * TJ);
                const float2 o01 = __ffma2_rn(make_float2(q.x, q.y), __ffma2_rn(acc[i][2 * j], rw2, c02), half2);
                const float2 o23 = __ffma2_rn(make_float2(q.z, q.w), __ffma2_rn(acc[i][2 * j + 1], rw2, c02), half2);
                if (!(Cfg::ABL & 4) || a.D < 0) { op[0] = o01.x; op[gstride] = o01.y; op[2 * gstride] = o23.x; op[3 * gstride] = o23.y; }
                op += 4 * gstride;
            }
        }
    };

    // Warps whose samples all fitted write their results before the CTA-wide vote (their stores overlap the tail
    // of the slower warps); the others write after the retry rounds.
    const bool early = MODE != 1 && MODE != 3 && Cfg::EARLY && __all_sync(0xffffffffu, left == 0u);
    if (early) epilogue();
    stamp();                                     // early epilogue done

    // ---- samples that did not fit their view's box (rough depth maps, silhouettes): synchronous staging
    //      rounds.  x origin = min over the samples left; y origin = min over those whose column fits, so the
    //      topmost of them lands inside the box and every round makes progress. ----
    if (__syncthreads_or(left != 0u)) {
        if (left != 0u) atomicOr(ctl_nv + kLeft, (int)left);
        __syncthreads();
        uint32_t views = (uint32_t)ctl[kLeft];
        uint32_t riter = 0;                          // CTA-uniform count of retry loads (mbarrier phase, slot)
        while (views != 0u) {
            const int v = __ffs(views) - 1;
            views &= views - 1u;
            {
                float rt[12];
                load_rt(v, rt);
                todo = positions(v, rt, ix, iy, false);
            }
            set_view(v);
            {   // drop what round 0 already gathered
                const int ox0 = ctl[kOrg + 2 * v], oy0 = ctl[kOrg + 2 * v + 1];
#pragma unroll
                for (int i = 0; i < PT; ++i) {
                    if (!((todo >> i) & 1u)) continue;
                    float f; int x0, y0;
                    floor_small(ix[i], f, x0);
                    floor_small(iy[i], f, y0);
                    if ((unsigned)(x0 - ox0) < (unsigned)(BW - 1) && (unsigned)(y0 - oy0) < (unsigned)(BH - 1)) todo &= ~(1u << i);
                }
            }
            while (__syncthreads_or(todo != 0u)) {
                int* slot = ctl_nv + kRetry + 2 * (riter & 1u);
                int kx = kNone;
#pragma unroll
                for (int i = 0; i < PT; ++i)
                    if ((todo >> i) & 1u) kx = min(kx, coord_key(ix[i]));
                kx = __reduce_min_sync(0xffffffffu, kx);
                if (lane == 0 && kx != kNone) atomicMin(slot, key_floor(kx));
                if (tid == 0) { volatile int* other = ctl + kRetry + 2 * ((riter + 1u) & 1u); other[0] = kNone; other[1] = kNone; }
                __syncthreads();
                const int ox = ctl[kRetry + 2 * (riter & 1u)];
                const float xlim = (float)(ox + BW - 1);
                int m = kNone;
#pragma unroll
                for (int i = 0; i < PT; ++i)
                    if (((todo >> i) & 1u) && ix[i] < xlim) m = min(m, coord_key(iy[i]));
                m = __reduce_min_sync(0xffffffffu, m);
                if (lane == 0 && m != kNone) atomicMin(slot + 1, key_floor(m));
                __syncthreads();
                const int oy = ctl[kRetry + 2 * (riter & 1u) + 1];
                if (tid == 0) {
                    mbar_expect_tx(bar0 + 16, Cfg::BOX_BYTES);
                    tma_load_3d(box0, &maps.s4, bar0 + 16, ox * 2, oy, (v * a.B + b) * J);
                }
                mbar_wait(bar0 + 16, riter & 1u);
                ++riter;
                todo = gather(v, box0, ox, oy, ix, iy, todo);
            }
            flush_stats(v);
        }
    }

    if (MODE == 1) {
        __syncthreads();                             // every warp has handed its sums over
        if (tid < 2 * a.V) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < Cfg::WARPS; ++w) sum += stats_s[MODE == 1 ? w : 0][tid];
            if (sum != 0.0) atomicAdd(a.stats + tid, sum);
        }
        return;
    }
    if (MODE == 3) return;
    if (!early) epilogue();
    stamp();                                     // end
#ifdef MDF_TUNING
    if (tracing) {
        const unsigned slot = atomicAdd(&g_trace_count, 1u);
        if (slot < kTraceSlots) {
            for (int k = 0; k < kTraceWords; ++k) g_trace[slot * kTraceWords + k] = k < trn ? tr[k] : 0;
            g_trace[slot * kTraceWords + kTraceWords - 1] = ((long long)blockIdx.x << 8) | (ty + TH * pg);
            g_trace[slot * kTraceWords + kTraceWords - 2] = trn;
        }
    }
#else
    (void)tracing;
#endif
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int encode_planes(CUtensorMap* tmap, const float* base, int W, int H, long long planes, int box_w, int box_h, int box_planes)
{
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) return MDF_ERR_UNSUPPORTED;
    // 3-D view of float4 planes [plane][y][x] in 8-byte elements: dim0 = 2*W, dim1 = H rows, dim2 = planes.
    // (8-byte elements because a box dimension is limited to 256 elements.)
    const cuuint64_t dims[3] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    const cuuint32_t box[3] = {(cuuint32_t)box_w * 2, (cuuint32_t)box_h, (cuuint32_t)box_planes};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return MDF_ERR_CUDA; }
    return MDF_OK;
}

struct StagedBuffers { const float* S4; const float* Q4; const float* CQ4; };

// optional timing hook of mdf_cost_volume_fwd_ex: two events recorded around the hot kernel alone
struct HotEvents { cudaEvent_t start = nullptr, stop = nullptr; };

template <class Cfg, int MODE = 0>
static int launch_staged(const StagedArgs& args, const StagedBuffers& buf, cudaStream_t stream, const HotEvents& ev = HotEvents())
{
    StagedMaps maps;
    int st = encode_planes(&maps.s4, buf.S4, args.W, args.H, (long long)args.V * args.B * Cfg::J, Cfg::BW, Cfg::BH, Cfg::J);
    if (st == MDF_OK) st = encode_planes(&maps.q4, buf.Q4, args.W, args.H, (long long)args.B * Cfg::J, Cfg::WX, Cfg::TH * Cfg::WY, Cfg::J);
    if (st == MDF_OK) st = encode_planes(&maps.cq4, buf.CQ4, args.W, args.H, (long long)args.B * Cfg::J, Cfg::WX, Cfg::TH * Cfg::WY, Cfg::J);
    if (st != MDF_OK) return st;
    auto kern = cost_volume_staged_kernel<Cfg, MODE>;
    MDF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    StagedArgs a = args;
    a.gn = make_grid_norm(a.H, a.W);
    a.tiles_x = (a.W + Cfg::WX - 1) / Cfg::WX;
    a.tiles_y = (a.H + Cfg::TH * Cfg::WY - 1) / (Cfg::TH * Cfg::WY);
    a.slabs = (a.D + Cfg::SLAB - 1) / Cfg::SLAB;
    const long long items = (long long)a.tiles_x * a.tiles_y * a.slabs * a.B;
    if (items <= 0) return MDF_OK;
    if (items > INT_MAX) return MDF_ERR_UNSUPPORTED;
    if (ev.start) cudaEventRecord(ev.start, stream);
    kern<<<(unsigned)items, dim3(32, Cfg::TH, Cfg::PG), Cfg::SMEM, stream>>>(maps, a);
    if (ev.stop) cudaEventRecord(ev.stop, stream);
    return launch_status();
}

// The product build ships ONE configuration per G (variant 0).  Tuning builds (-DMDF_TUNING) add the shape variants,
// the diagnostic variants (branch-free samples, one reciprocal per group, phase timestamps, ablations) behind
// algo = 16 + k: every one of them was measured within +-3 % of variant 0 or slower (profiles/r02_*).
//                         G  PT TH PG  BW  BH MINB CQS
// (round 2: boxes widened 40 -> 48, 64 -> 80, 64 -> 80 texels: fewer retry rounds at silhouettes and wide search ranges;
//  -3 % at DTU 1600x1152, -5.5 % at 1920x1056 N=7, -5 % with the widest HyposByFit ranges; profiles/r02_box_sweep_*.log)
using CfgG32_0 = StagedCfg<32, 1, 2, 4, 48, 6, 2, true>;     // 256 thr x 2 CTAs; tile 32x2, slab 4 planes; 2 x 36 KiB boxes + 2 x 8 KiB tiles
using CfgG16_0 = StagedCfg<16, 2, 2, 4, 80, 6, 2, false>;    // 256 thr x 2 CTAs; tile 32x2, slab 8 planes; 2 x 30 KiB boxes
using CfgG8_0  = StagedCfg<8, 4, 4, 2, 80, 10, 2, false>;    // 256 thr x 2 CTAs; tile 32x4, slab 8 planes; 2 x 25 KiB boxes

// the same with narrower, taller warp footprints (eval mode only): tile 8 x 8 / 16 x 4 pixels, boxes of the same bytes
using CfgG32_W8  = StagedCfg<32, 1, 2, 4, 24, 12, 2, true, true, false, false, false, 0, 8>;
using CfgG16_W16 = StagedCfg<16, 2, 2, 4, 56, 8, 2, false, true, false, false, false, 0, 16>;
using CfgG16_W8  = StagedCfg<16, 2, 2, 4, 40, 12, 2, false, true, false, false, false, 0, 8>;

// lanes that fall beyond the right edge of the map, per warp width
static inline int pick_warp_width(int W, bool has16, bool has8)
{
    auto waste = [&](int wx) { return (W + wx - 1) / wx * wx - W; };
    int best = 32;
    if (has16 && waste(16) < waste(best)) best = 16;
    if (has8 && waste(8) < waste(best)) best = 8;
    // a narrower footprint only where it recovers more than 3 % of the lanes
    return (waste(32) - waste(best)) * 100 > 3 * W ? best : 32;
}

// the configuration of a stage; MODE 1 / 2 = training (batch statistics / per-view BatchNorm folds)
template <int MODE>
static int launch_staged_default(int G, const StagedArgs& a, const StagedBuffers& S, cudaStream_t stream, const HotEvents& ev = HotEvents())
{
    if (MODE == 0) {
        if (G == 32 && pick_warp_width(a.W, false, true) == 8) return launch_staged<CfgG32_W8, 0>(a, S, stream, ev);
        if (G == 16) {
            const int wx = pick_warp_width(a.W, true, true);
            if (wx == 16) return launch_staged<CfgG16_W16, 0>(a, S, stream, ev);
            if (wx == 8) return launch_staged<CfgG16_W8, 0>(a, S, stream, ev);
        }
    }
    if (G == 32) return launch_staged<CfgG32_0, MODE>(a, S, stream, ev);
    if (G == 16) return launch_staged<CfgG16_0, MODE>(a, S, stream, ev);
    if (G == 8) return launch_staged<CfgG8_0, MODE>(a, S, stream, ev);
    return MDF_ERR_UNSUPPORTED;
}
template <int MODE>
static int launch_staged_train(int G, const StagedArgs& a, const StagedBuffers& S, cudaStream_t stream)
{
    return launch_staged_default<MODE>(G, a, S, stream);
}
// the eval-mode launch for other translation units (defined in mdf_cost_volume.cu: one instantiation of the kernels)
int launch_staged_eval(int G, const StagedArgs& a, const StagedBuffers& S, cudaStream_t stream);

#ifdef MDF_TUNING
//                         G  PT TH PG  BW  BH MINB CQS   EARLY  BF    RCP2   TRACE  ABL
using CfgG32_1 = StagedCfg<32, 1, 4, 2, 40, 7, 2, true>;     // tile 32x4, slab 2 planes
using CfgG32_2 = StagedCfg<32, 1, 4, 2, 40, 7, 2, true, false>;
using CfgG32_3 = StagedCfg<32, 1, 1, 8, 48, 4, 2, true>;     // tile 32x1, slab 8 planes
using CfgG32_4 = StagedCfg<32, 1, 2, 4, 40, 6, 2, true>;     // round 1's default box: 2 x 30 KiB
using CfgG32_5 = StagedCfg<32, 1, 2, 4, 48, 7, 2, true>;     // taller box: 2 x 42 KiB
using CfgG32_9 = StagedCfg<32, 1, 2, 4, 56, 6, 2, true>;     // wider still: 2 x 42 KiB
using CfgG16_5 = StagedCfg<16, 2, 2, 4, 64, 8, 2, false>;    // taller box: 2 x 32 KiB
using CfgG16_7 = StagedCfg<16, 2, 2, 4, 64, 6, 2, false>;    // round 1's default box: 2 x 24 KiB
using CfgG16_9 = StagedCfg<16, 2, 2, 4, 96, 6, 2, false>;    // wider still: 2 x 36 KiB
using CfgG8_9x = StagedCfg<8, 4, 4, 2, 64, 10, 2, false>;    // round 1's default box: 2 x 20 KiB
using CfgG8_9y = StagedCfg<8, 4, 4, 2, 64, 12, 2, false>;    // taller box: 2 x 24 KiB
using CfgG8_9z = StagedCfg<8, 4, 4, 2, 96, 10, 2, false>;    // wider still: 2 x 30 KiB
using CfgG16_1 = StagedCfg<16, 2, 4, 2, 64, 8, 2, false>;    // tile 32x4, slab 4 planes
using CfgG16_2 = StagedCfg<16, 2, 4, 2, 48, 8, 2, false>;    // narrower box
using CfgG16_3 = StagedCfg<16, 2, 1, 8, 80, 4, 2, false>;    // tile 32x1, slab 16 planes
using CfgG8_1  = StagedCfg<8, 4, 8, 1, 64, 14, 2, false>;    // tile 32x8, slab 4 planes
using CfgG8_2  = StagedCfg<8, 2, 4, 2, 64, 10, 3, false>;    // 3 CTAs / SM, slab 4 planes
using CfgG8_3  = StagedCfg<8, 2, 2, 4, 64, 6, 3, false>;     // tile 32x2, slab 8 planes, 3 CTAs / SM
using CfgG16_4 = StagedCfg<16, 2, 2, 4, 64, 6, 2, false, true, true>;          // branch-free samples
using CfgG8_4  = StagedCfg<8, 4, 4, 2, 64, 10, 2, false, true, true>;
using CfgG8_5  = StagedCfg<8, 2, 4, 2, 64, 10, 3, false, true, true>;
using CfgG32_6 = StagedCfg<32, 1, 2, 4, 48, 6, 2, true, true, false, true>;    // one reciprocal per group
using CfgG16_6 = StagedCfg<16, 2, 2, 4, 64, 6, 2, false, true, false, true>;
using CfgG8_6  = StagedCfg<8, 4, 4, 2, 64, 10, 2, false, true, false, true>;
using CfgG8_7  = StagedCfg<8, 2, 4, 2, 64, 10, 3, false, true, false, true>;
template <int ABL> using CfgG32_A = StagedCfg<32, 1, 2, 4, 48, 6, 2, true, true, false, false, false, ABL>;   // ablations
template <int ABL> using CfgG16_A = StagedCfg<16, 2, 2, 4, 80, 6, 2, false, true, false, false, false, ABL>;
template <int ABL> using CfgG8_A  = StagedCfg<8, 4, 4, 2, 80, 10, 2, false, true, false, false, false, ABL>;
using CfgG32_T = StagedCfg<32, 1, 2, 4, 48, 6, 2, true, true, false, false, true>;    // phase timestamps
using CfgG16_T = StagedCfg<16, 2, 2, 4, 80, 6, 2, false, true, false, false, true>;
using CfgG8_T  = StagedCfg<8, 4, 4, 2, 80, 10, 2, false, true, false, false, true>;

// algo = 16 + variant: 0 default, 1-3 tile shapes, 4 / 5 / 7 / 12 / 14 / 16 box shapes (4, 7 at G16, 14 at G8: round 1's defaults),
// 4-5 (G16 4, G8 4-5) branch-free, 6-7 one reciprocal per group, 8-13 ablations
// (no MUFU / no tap loads / no stores / cheap positions / no cq loads / all of them), 15 phase timestamps
static int launch_staged_variant(int G, int variant, const StagedArgs& a, const StagedBuffers& S, cudaStream_t stream, const HotEvents& ev)
{
#define MDF_V(k, Cfg) case k: return launch_staged<Cfg>(a, S, stream, ev)
    if (G == 32) {
        switch (variant) {
            MDF_V(0, CfgG32_0); MDF_V(1, CfgG32_1); MDF_V(2, CfgG32_2); MDF_V(3, CfgG32_3); MDF_V(4, CfgG32_4); MDF_V(5, CfgG32_5); MDF_V(14, CfgG32_9); MDF_V(6, CfgG32_6);
            MDF_V(8, CfgG32_A<1>); MDF_V(9, CfgG32_A<2>); MDF_V(10, CfgG32_A<4>); MDF_V(11, CfgG32_A<8>); MDF_V(12, CfgG32_A<16>);
            MDF_V(13, CfgG32_A<31>); MDF_V(15, CfgG32_T);
        }
    } else if (G == 16) {
        switch (variant) {
            MDF_V(0, CfgG16_0); MDF_V(1, CfgG16_1); MDF_V(2, CfgG16_2); MDF_V(3, CfgG16_3); MDF_V(4, CfgG16_4); MDF_V(5, CfgG16_5); MDF_V(6, CfgG16_6); MDF_V(7, CfgG16_7); MDF_V(14, CfgG16_9);
            MDF_V(8, CfgG16_A<1>); MDF_V(9, CfgG16_A<2>); MDF_V(10, CfgG16_A<4>); MDF_V(11, CfgG16_A<8>); MDF_V(13, CfgG16_A<15>);
            MDF_V(15, CfgG16_T);
        }
    } else if (G == 8) {
        switch (variant) {
            MDF_V(0, CfgG8_0); MDF_V(1, CfgG8_1); MDF_V(2, CfgG8_2); MDF_V(3, CfgG8_3); MDF_V(4, CfgG8_4); MDF_V(5, CfgG8_5);
            MDF_V(6, CfgG8_6); MDF_V(7, CfgG8_7);
            MDF_V(8, CfgG8_A<1>); MDF_V(9, CfgG8_A<2>); MDF_V(10, CfgG8_A<4>); MDF_V(11, CfgG8_A<8>); MDF_V(13, CfgG8_A<15>);
            MDF_V(15, CfgG8_T); MDF_V(14, CfgG8_9x); MDF_V(12, CfgG8_9y); MDF_V(16, CfgG8_9z);
        }
    }
#undef MDF_V
    return MDF_ERR_UNSUPPORTED;
}
#endif  // MDF_TUNING

}  // namespace mdf

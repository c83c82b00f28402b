// mdf_prob_head.cu -- the tail of MDF-Net's 3-D regulariser fused with the regression head, for sm_100a.
//
// Reference chain replaced by ONE launch per stage:
//   x = self.prob(x).squeeze(1)        net/unit/regular.py:43,67 / :110,130   Conv3d(c0, 1, k=3, pad=1, bias=False)
//   F.softmax(x, dim=1)                net/unit/regular.py:69,133
//   depth_regression                   net/unit/regress.py:5-7
//   confidence_regress (+ nearest x2)  net/unit/regress.py:9-25, net/core.py:75-77          (last stage)
//   HyposByFit curve fit               net/unit/depthhypos.py:78-125, 169-215               (stages 0 and 1)
// The 3-D CNN body stays on cuDNN (north_star); only its last layer -- a convolution with ONE output channel, the shape
// implicit-GEMM libraries handle worst -- moves here, so that the logits never exist in memory: the feature volume
// x (B,c0,D,H,W) is read once, the outputs are written once.
//
// Mapping: a warp is 8 x 4 lanes, a lane owns 4 consecutive pixels (LDG.128) -> tile of 32 x 4 pixels; the D output
// planes are split into slabs of DSLAB planes over the warps of the CTA (the coarse stages have few pixels and deep
// columns: splitting D is what fills the machine).  A thread keeps DSLAB*4 accumulators and walks c0 x (DSLAB+2) input
// planes: per plane three row loads (y-1, y, y+1; overlaps are served by L1), the x-1 / x+4 neighbours by two warp
// shuffles (edge lanes of the tile: one predicated scalar load), then up to 27*4 FFMAs with the channel's 27 weights in
// registers.  Zero padding = predicated loads.  The logits meet in shared memory ([d][pixel]); then one thread per pixel
// runs the column tail (softmax, expectation, confidence window, curve fit) on registers, lanes along x.
//
// Bounds: HBM 4*B*c0*D*H*W bytes read (88 / 88 / 118 MB at BASELINE configs[1]) vs 27*c0*D*H*W FMAs per batch item
// (0.60 / 0.60 / 0.80 G): 13 / 13 / 18 us of HBM time against 17 / 17 / 22 us of FP32 issue -- FMA bound by a small margin.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mdf_b200_debug.h"
#include "mdf_common.cuh"
#include "mdf_host.cuh"
#include "mdf_tail.cuh"
#include "mdf_tma.cuh"

namespace mdf {

struct ProbHeadArgs {
    const float* x;        // (B,C,D,H,W) last feature volume of the regulariser
    const float* w;        // (C,3,3,3)   weight of Conv3d(C,1,3) = (1,C,3,3,3) contiguous
    const float* hypos;    // (B,D) or (B,D,H,W)
    float* logits;         // (B,D,H,W) or nullptr (diagnostic / unfused use)
    float* prob;           // (B,D,H,W) or nullptr
    float* depth;          // (B,H,W) or nullptr
    float* conf;           // (B,H*up,W*up) or nullptr
    float* s;              // (B,H,W) or nullptr (FIT != 0)
    int per_pixel, B, C, H, W;
    int conf_n, pad_front, pad_back, up;
};

template <int PX> struct VecLoad;
template <> struct VecLoad<1> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void st(float* p, const float (&v)[1]) { p[0] = v[0]; }
};
template <> struct VecLoad<4> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4])
    {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

constexpr int kMaxProbChannels = 64;

// what a padded row reads (zero padding of the convolution)
__device__ __align__(16) const float g_prob_zeros[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

// Tile geometry: a warp is 8 x 4 lanes, a lane PX consecutive pixels -> tile of 8*PX x 4 pixels; a CTA is NT such tiles
// stacked in y times D/DSLAB warps along the depth axis.
template <int D_, int DSLAB_, int NT_, int PX_>
struct ProbCfg {
    static constexpr int D = D_, DSLAB = DSLAB_, NT = NT_, PX = PX_;
    static constexpr int NWD = D / DSLAB;
    static constexpr int WARPS = NWD * NT, THREADS = 32 * WARPS;
    static constexpr int TW = 8 * PX, TH = 4, TPIX = TW * TH;       // pixels of one tile
    static_assert(D % DSLAB == 0, "slabs must tile the depth axis");
};

// Column tail (one thread = one pixel, sequential sums in the reference's order): softmax (regular.py:69,133),
// expectation (regress.py:7), confidence window (regress.py:9-25 + core.py:75-77), curve fit (depthhypos.py:78-125,
// 169-215).  The column stays in shared memory (cs[d * S]: logits on entry, probabilities on return) and is swept a few
// times with short unrolled loops, so that the tail does not set the register allocation of the convolution.
template <int D, int S, int FIT>
__device__ __forceinline__ void column_tail(float* __restrict__ cs, const ProbHeadArgs& a, int b, int y, int x)
{
    constexpr int CH = 8;                           // planes per batch: their hypotheses are requested together
    static_assert(D % CH == 0, "the configured depths are multiples of 8");
    const size_t HW = (size_t)a.H * a.W, p = (size_t)y * a.W + x;
    const float* __restrict__ hcol = a.per_pixel ? a.hypos + (size_t)b * D * HW + p : a.hypos + (size_t)b * D;
    const size_t hstride = a.per_pixel ? HW : 1;
    const bool need_h = a.depth != nullptr || FIT != 0;
    float m = cs[0];
#pragma unroll 8
    for (int d = 1; d < D; ++d) m = fmaxf(m, cs[d * S]);
    float sum = 0.0f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
        const float e = expf(__fsub_rn(cs[d * S], m));
        cs[d * S] = e;
        sum = __fadd_rn(sum, e);
    }
    float* __restrict__ pc = a.prob ? a.prob + (size_t)b * D * HW + p : nullptr;
    float acc = 0.0f, eidx = 0.0f;
    double hs = 0.0;
    for (int d0 = 0; d0 < D; d0 += CH) {
        float hv[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) hv[i] = need_h ? __ldg(hcol + (size_t)(d0 + i) * hstride) : 0.0f;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const float pr = __fdiv_rn(cs[(d0 + i) * S], sum);
            cs[(d0 + i) * S] = pr;
            if (pc) pc[(size_t)(d0 + i) * HW] = pr;
            eidx = __fadd_rn(eidx, __fmul_rn(pr, (float)(d0 + i)));                    // regress.py:15-17
            acc = __fadd_rn(acc, __fmul_rn(pr, hv[i]));                                // regress.py:7
            if (FIT == 1) hs += (double)hv[i];
        }
    }
    if (a.depth) a.depth[(size_t)b * HW + p] = acc;
    if (FIT == 2) {
        LaplaceSums ls;
        for (int d0 = 0; d0 < D; d0 += CH) {
            float hv[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) hv[i] = __ldg(hcol + (size_t)(d0 + i) * hstride);     // L1 hits
#pragma unroll
            for (int i = 0; i < CH; ++i) ls.add(hv[i], acc, cs[(d0 + i) * S]);
        }
        a.s[(size_t)b * HW + p] = ls.scale();
    } else if (FIT == 1) {
        const double mean = hs / (double)D;
        GaussMoments gm;
        for (int d0 = 0; d0 < D; d0 += CH) {
            float hv[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) hv[i] = __ldg(hcol + (size_t)(d0 + i) * hstride);
#pragma unroll 2
            for (int i = 0; i < CH; ++i) gm.add((double)hv[i] - mean, cs[(d0 + i) * S]);
        }
        a.s[(size_t)b * HW + p] = gm.scale(D);
    }
    if (a.conf) {
        const int Dp = D + a.pad_front + a.pad_back - a.conf_n + 1;
        const int kk = max(0, min((int)eidx, Dp - 1));
        float sw = 0.0f;
        for (int j = 0; j < a.conf_n; ++j) {
            const int d = kk - a.pad_front + j;
            sw = __fadd_rn(sw, (d >= 0 && d < D) ? cs[d * S] : 0.0f);
        }
        const float fn = (float)a.conf_n;
        store_upsampled(a.conf, __fmul_rn(fn, __fdiv_rn(sw, fn)), b, y, x, a.H, a.W, a.up);
    }
}

// The tail's per-pixel hypotheses are the only global loads after the convolution: ask L2 for them up front.
template <int D>
__device__ __forceinline__ void prefetch_hypotheses(const ProbHeadArgs& a, int b, int y, int x)
{
    if (!a.per_pixel || !(a.depth != nullptr || a.s != nullptr) || y >= a.H || x >= a.W) return;
    const size_t HW = (size_t)a.H * a.W;
    const float* hcol = a.hypos + (size_t)b * D * HW + (size_t)y * a.W + x;
#pragma unroll 8
    for (int d = 0; d < D; ++d) asm volatile("prefetch.global.L2 [%0];" ::"l"(hcol + (size_t)d * HW));
}

template <class Cfg, int FIT>
__global__ void __launch_bounds__(Cfg::THREADS)
prob_head_kernel(const ProbHeadArgs a)
{
    constexpr int D = Cfg::D, DSLAB = Cfg::DSLAB, NT = Cfg::NT, PX = Cfg::PX, TPIX = Cfg::TPIX, TW = Cfg::TW;
    __shared__ __align__(16) float w_s[kMaxProbChannels * 28];      // 27 weights per channel, padded to 28
    __shared__ __align__(16) float col_s[NT * D * TPIX];            // logits of the CTA's pixels: [tile][d][pixel]
    const int lane = threadIdx.x, wz = threadIdx.y, t = threadIdx.z;
    const int tid = lane + 32 * (wz + Cfg::NWD * t);
    const int H = a.H, W = a.W, C = a.C;
    for (int i = tid; i < C * 28; i += Cfg::THREADS) {
        const int c = i / 28, k = i % 28;
        w_s[i] = k < 27 ? __ldg(a.w + c * 27 + k) : 0.0f;
    }
    __syncthreads();

    const int b = blockIdx.z;
    const int lx = lane & 7, ly = lane >> 3;
    const int xt = blockIdx.x * TW, yt = (blockIdx.y * NT + t) * Cfg::TH;
    const int x0 = xt + lx * PX, y = yt + ly;
    const int d0 = wz * DSLAB;
    const size_t HW = (size_t)H * W;
    const bool col_ok = x0 < W;                                      // W % PX == 0 (host checked): the whole vector is inside

    float lg[DSLAB][PX];
#pragma unroll
    for (int j = 0; j < DSLAB; ++j)
#pragma unroll
        for (int k = 0; k < PX; ++k) lg[j][k] = 0.0f;

    // Zero padding without predicates in the loop: every row of this lane (y-1, y, y+1) is a (pointer, plane stride)
    // pair; a row outside the image -- or a lane beyond the right edge -- points at a block of zeros with stride 0.
    // The halo value of the tile's edge lanes (x0-1 for lx == 0, x0+PX for lx == 7) is a second such pair; the other
    // lanes get their neighbours by shuffle and skip that load (one loop-invariant predicate).
    const float* __restrict__ zeros = g_prob_zeros;
    const bool edge = lx == 0 || lx == 7;
    const float* rowp[3];
    const float* halop[3];
    uint32_t rows[3], halos[3];                                     // plane strides in floats (0 for the zero block)
    const float* __restrict__ xbase = a.x + (size_t)b * C * D * HW + (ptrdiff_t)d0 * (ptrdiff_t)HW;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        const bool rv = col_ok && (unsigned)yy < (unsigned)H;
        const int hx = lx == 0 ? x0 - 1 : x0 + PX;
        const bool hv = rv && edge && (unsigned)hx < (unsigned)W;
        rowp[ky] = rv ? xbase + (size_t)yy * W + x0 : zeros;
        rows[ky] = rv ? (uint32_t)HW : 0u;
        halop[ky] = hv ? xbase + (size_t)yy * W + hx : zeros;
        halos[ky] = hv ? (uint32_t)HW : 0u;
    }

    // Input planes of a slab: local dl = Q0 .. Q1 (global d0 + dl).  The halo planes -1 and DSLAB do not exist when one
    // warp owns the whole column.  The walk over (channel, plane) is software pipelined: the three rows of the NEXT
    // plane are requested (registers, two plane buffers) before the current plane's 27*PX FFMAs are issued.
    constexpr int Q0 = Cfg::NWD == 1 ? 0 : -1, Q1 = Cfg::NWD == 1 ? DSLAB - 1 : DSLAB, NP = Q1 - Q0 + 1;
    static_assert(NP % 2 == 0, "the two plane buffers alternate statically");
    struct Plane { float v[3][PX]; float h[3]; };
    // Running pointers to the next plane to be requested (planes are requested in the order they are consumed):
    // advanced by one plane per request, by D - NP + 1 planes when the walk wraps to the next channel.
    const ptrdiff_t first = (ptrdiff_t)Q0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) { rowp[ky] += first * (ptrdiff_t)rows[ky]; halop[ky] += first * (ptrdiff_t)halos[ky]; }
    // planes outside the volume (dl = -1 of the first slab, dl = DSLAB of the last) read as zeros
    auto load_next = [&](const int dl /* literal */, const bool wrap /* literal */, Plane& P) {
        const bool maybe_outside = Cfg::NWD > 1 && (dl == -1 || dl == DSLAB);
        const bool pv = !maybe_outside || (unsigned)(d0 + dl) < (unsigned)D;       // warp uniform
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const float* __restrict__ rp = pv ? rowp[ky] : zeros;
            const float* __restrict__ hp = pv ? halop[ky] : zeros;
            VecLoad<PX>::ld(rp, P.v[ky]);
            P.h[ky] = 0.0f;
            if (edge) P.h[ky] = __ldg(hp);
            const ptrdiff_t step = wrap ? (ptrdiff_t)(D - NP + 1) : (ptrdiff_t)1;
            rowp[ky] += step * (ptrdiff_t)rows[ky];
            halop[ky] += step * (ptrdiff_t)halos[ky];
        }
    };
    auto compute_plane = [&](const int dl /* literal after unrolling */, const Plane& P, const float (&wt)[28]) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const float sl = __shfl_up_sync(0xffffffffu, P.v[ky][PX - 1], 1);
            const float sr = __shfl_down_sync(0xffffffffu, P.v[ky][0], 1);
            const float left = lx != 0 ? sl : P.h[ky], right = lx != 7 ? sr : P.h[ky];
            // plane dl feeds the slab-local outputs j = dl - kd + 1
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
                const int j = dl - kd + 1;
                if (j < 0 || j >= DSLAB) continue;
                const float w0 = wt[kd * 9 + ky * 3], w1 = wt[kd * 9 + ky * 3 + 1], w2 = wt[kd * 9 + ky * 3 + 2];
#pragma unroll
                for (int k = 0; k < PX; ++k) {
                    const float tl = k == 0 ? left : P.v[ky][k > 0 ? k - 1 : 0];
                    const float tr = k == PX - 1 ? right : P.v[ky][k < PX - 1 ? k + 1 : 0];
                    lg[j][k] = fmaf(w2, tr, fmaf(w1, P.v[ky][k], fmaf(w0, tl, lg[j][k])));
                }
            }
        }
    };

    Plane P[2];
    load_next(Q0, NP == 1, P[0]);
    for (int c = 0; c < C; ++c) {
        float wt[28];
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(w_s + c * 28 + 4 * q);
            wt[4 * q] = v4.x; wt[4 * q + 1] = v4.y; wt[4 * q + 2] = v4.z; wt[4 * q + 3] = v4.w;
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (q + 1 < NP) load_next(Q0 + q + 1, q + 2 == NP, P[(q + 1) & 1]);
            else if (c + 1 < C) load_next(Q0, false, P[0]);
            compute_plane(Q0 + q, P[q & 1], wt);
        }
    }

    // logits -> shared memory [tile][d][ly][lx*PX ..] (and to global memory when asked for)
    {
        float* cs = col_s + ((size_t)t * D + d0) * TPIX + ly * TW + lx * PX;
        const bool ok = col_ok && y < H;
#pragma unroll
        for (int j = 0; j < DSLAB; ++j) {
            VecLoad<PX>::st(cs + j * TPIX, lg[j]);
            if (a.logits && ok) VecLoad<PX>::st(a.logits + ((size_t)b * D + d0 + j) * HW + (size_t)y * W + x0, lg[j]);
        }
    }
    if (!a.prob && !a.depth && !a.conf && !a.s) return;              // uniform over the grid
    __syncthreads();

    // ---- tail: one thread per pixel column, lanes along x (coalesced hypotheses / prob / depth rows) ----
    for (int col = tid; col < NT * TPIX; col += Cfg::THREADS) {
        const int tt = col / TPIX, pp = col % TPIX;
        const int yy = (blockIdx.y * NT + tt) * Cfg::TH + pp / TW, xx = xt + pp % TW;
        if (yy >= H || xx >= W) continue;
        column_tail<D, TPIX, FIT>(col_s + (size_t)tt * D * TPIX + pp, a, b, yy, xx);
    }
}

// ------------------------------------------------------------------------------------------------
// The fast variant (W % 4 == 0, 16-byte aligned volume): the feature volume comes in by TMA.
//   * x is a 4-D tensor map [B*c0][D][H][W]; one pipeline stage = ONE channel of the CTA's tile with its halo:
//     box [D (+2)][4*NT + 2][40] floats at (xt - 4, yt - 1, -1 | 0, b*c0 + c).  TMA's out-of-bounds zero fill IS the
//     convolution's zero padding (left / right / top / bottom / first and last depth plane): the compute loop has no
//     bounds predicate at all.
//   * a dedicated producer warp keeps NSTAGE channels in flight (full / empty mbarriers per stage); the consumer
//     warps -- the same 8 x 4-lane tiles and depth slabs as above -- read their rows with LDS.128 (a quarter warp reads
//     128 contiguous bytes: conflict free), take the x-1 / x+4 neighbours from the adjacent lanes by shuffle (the two
//     edge lanes of a row: one LDS.32 from the halo columns of the box) and issue 27*4 FFMAs per plane.
//   * the logits meet in shared memory (the stage buffers are reused) and the column tail runs as above.
// ------------------------------------------------------------------------------------------------
template <int D_, int DSLAB_, int NT_, int NSTAGE_, int MINB_>
struct ProbTmaCfg {
    static constexpr int D = D_, DSLAB = DSLAB_, NT = NT_, NSTAGE = NSTAGE_, MINB = MINB_, PX = 4;
    static constexpr int NWD = D / DSLAB;
    static constexpr int CWARPS = NWD * NT, THREADS = 32 * (CWARPS + 1);       // consumers + one producer warp
    static constexpr int TW = 32, TH = 4, TPIX = TW * TH;
    static constexpr int DHALO = NWD > 1 ? 1 : 0;                              // one warp owns the whole column: no depth halo
    static constexpr int BW = 40, BH = TH * NT + 2, BD = D + 2 * DHALO;
    static constexpr int STAGE_BYTES = BD * BH * BW * 4;
    static constexpr int STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
    static constexpr int COL_BYTES = NT * D * TPIX * 4;
    static constexpr int OFF_W = NSTAGE * STAGE_STRIDE;
    static constexpr int OFF_BAR = OFF_W + kMaxProbChannels * 28 * 4;
    static constexpr size_t SMEM = OFF_BAR + 16 * NSTAGE + 128 /* alignment slack */;
    static_assert(D % DSLAB == 0, "slabs must tile the depth axis");
    static_assert(COL_BYTES <= NSTAGE * STAGE_STRIDE, "the logits reuse the stage buffers");
    static_assert(BD <= 256 && BH <= 256, "TMA box dimensions are limited to 256 elements");
};

template <class Cfg, int FIT>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
prob_head_tma_kernel(const __grid_constant__ CUtensorMap xmap, const ProbHeadArgs a)
{
    constexpr int D = Cfg::D, DSLAB = Cfg::DSLAB, NT = Cfg::NT, NSTAGE = Cfg::NSTAGE, PX = 4, TPIX = Cfg::TPIX, TW = Cfg::TW;
    constexpr int BW = Cfg::BW, BH = Cfg::BH, CWARPS = Cfg::CWARPS, NWD = Cfg::NWD;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t pad = (128u - (smem_u32(smem_raw) & 127u)) & 127u;
    uint8_t* base = smem_raw + pad;
    const uint32_t base_s = smem_u32(base);
    float* w_s = reinterpret_cast<float*>(base + Cfg::OFF_W);
    const uint32_t full0 = base_s + Cfg::OFF_BAR, empty0 = full0 + 8 * NSTAGE;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);        // warp uniform, and the compiler knows it (no divergence
                                                                   // handling around the consumers' shuffles)
    const int H = a.H, W = a.W, C = a.C;
    const int b = blockIdx.z;
    const int xt = blockIdx.x * TW, yt0 = blockIdx.y * NT * Cfg::TH;
    const size_t HW = (size_t)H * W;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, CWARPS); }
        fence_barrier_init();
    }
    for (int i = tid; i < C * 28; i += Cfg::THREADS) {
        const int c = i / 28, k = i % 28;
        w_s[i] = k < 27 ? __ldg(a.w + c * 27 + k) : 0.0f;
    }
    __syncthreads();

    // hypotheses of the tile's pixels -> L2, one request per 32-byte sector (they are read by the tail, much later)
    for (int col = tid * 8; col < NT * TPIX; col += Cfg::THREADS * 8)
        prefetch_hypotheses<D>(a, b, yt0 + (col / TPIX) * Cfg::TH + (col % TPIX) / TW, xt + (col % TPIX) % TW);

    float lg[DSLAB][PX];
#pragma unroll
    for (int j = 0; j < DSLAB; ++j)
#pragma unroll
        for (int k = 0; k < PX; ++k) lg[j][k] = 0.0f;
    const int wz = warp % NWD, t = warp / NWD;             // consumers: depth slab, tile
    const int lx = lane & 7, ly = lane >> 3;
    const int d0 = wz * DSLAB;

    if (warp == CWARPS) {
        // ---- producer: one lane keeps NSTAGE channels in flight ----
        if (lane == 0) {
            for (int c = 0; c < C; ++c) {
                const int s = c % NSTAGE, k = c / NSTAGE;
                if (k > 0) mbar_wait(empty0 + 8 * s, (uint32_t)(k - 1) & 1u);       // every consumer warp has left channel c - NSTAGE
                mbar_expect_tx(full0 + 8 * s, Cfg::STAGE_BYTES);
                tma_load_4d(base_s + s * Cfg::STAGE_STRIDE, &xmap, full0 + 8 * s, xt - 4, yt0 - 1, -Cfg::DHALO, b * C + c);
            }
        }
    } else {
        // ---- consumers ----
        // this lane's (row y-1, column x0) of plane d0-1 (or d0) inside a stage, and the halo column of the edge lanes
        const uint32_t lane_off = (uint32_t)((d0 * BH + t * Cfg::TH + ly) * BW + 4 + lx * PX) * 4u;
        // halo column of this lane's row: x0+4 for the last lane of a row, x-1 of the ROW'S FIRST lane for all others (lanes
        // 1..6 never use it: they read the same word as lane 0 -- a broadcast -- so the load needs no predicate or branch;
        // the 8 distinct words of a warp fall into 8 distinct banks)
        const uint32_t row_off = (uint32_t)((d0 * BH + t * Cfg::TH + ly) * BW) * 4u;
        const uint32_t halo_off = row_off + (lx == 7 ? (uint32_t)(4 + 8 * PX) * 4u : 12u);
        constexpr int NP = DSLAB + 2 * Cfg::DHALO;          // input planes of a slab: local q = 0 .. NP-1, dl = q - DHALO
        for (int c = 0; c < C; ++c) {
            const int s = c % NSTAGE;
            float wt[28];
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                const float4 v4 = *reinterpret_cast<const float4*>(w_s + c * 28 + 4 * q);
                wt[4 * q] = v4.x; wt[4 * q + 1] = v4.y; wt[4 * q + 2] = v4.z; wt[4 * q + 3] = v4.w;
            }
            mbar_wait(full0 + 8 * s, (uint32_t)(c / NSTAGE) & 1u);
            const uint32_t sb = base_s + s * Cfg::STAGE_STRIDE;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int dl = q - Cfg::DHALO;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const uint32_t off = (uint32_t)((q * BH + ky) * BW) * 4u;          // literal
                    const float4 v4 = lds128(sb + lane_off + off);
                    const float hv = lds32(sb + halo_off + off);
                    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
                    const float sl = __shfl_up_sync(0xffffffffu, v4.w, 1);
                    const float sr = __shfl_down_sync(0xffffffffu, v4.x, 1);
                    const float left = lx != 0 ? sl : hv, right = lx != 7 ? sr : hv;
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                        const int j = dl - kd + 1;                                      // plane dl feeds slab-local output j
                        if (j < 0 || j >= DSLAB) continue;
                        const float w0 = wt[kd * 9 + ky * 3], w1 = wt[kd * 9 + ky * 3 + 1], w2 = wt[kd * 9 + ky * 3 + 2];
#pragma unroll
                        for (int k = 0; k < PX; ++k) {
                            const float tl = k == 0 ? left : v[k > 0 ? k - 1 : 0];
                            const float tr = k == PX - 1 ? right : v[k < PX - 1 ? k + 1 : 0];
                            lg[j][k] = fmaf(w2, tr, fmaf(w1, v[k], fmaf(w0, tl, lg[j][k])));
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
        }
    }
    __syncthreads();                                        // every stage has been consumed: the buffers become the logits' home

    float* col_s = reinterpret_cast<float*>(base);
    if (warp < CWARPS) {
        const int x0 = xt + lx * PX, y = yt0 + t * Cfg::TH + ly;
        float* cs = col_s + ((size_t)t * D + d0) * TPIX + ly * TW + lx * PX;
        const bool ok = x0 < W && y < H;
#pragma unroll
        for (int j = 0; j < DSLAB; ++j) {
            *reinterpret_cast<float4*>(cs + j * TPIX) = make_float4(lg[j][0], lg[j][1], lg[j][2], lg[j][3]);
            if (a.logits && ok)
                *reinterpret_cast<float4*>(a.logits + ((size_t)b * D + d0 + j) * HW + (size_t)y * W + x0) =
                    make_float4(lg[j][0], lg[j][1], lg[j][2], lg[j][3]);
        }
    }
    if (!a.prob && !a.depth && !a.conf && !a.s) return;     // uniform over the grid
    __syncthreads();
    for (int col = tid; col < NT * TPIX; col += Cfg::THREADS) {
        const int tt = col / TPIX, pp = col % TPIX;
        const int yy = yt0 + tt * Cfg::TH + pp / TW, xx = xt + pp % TW;
        if (yy >= H || xx >= W) continue;
        column_tail<D, TPIX, FIT>(col_s + (size_t)tt * D * TPIX + pp, a, b, yy, xx);
    }
}

template <class Cfg>
static int launch_prob_head_tma(const ProbHeadArgs& a, int fit, cudaStream_t stream)
{
    EncodeTiledFn encode = get_encode_fn();
    if (encode == nullptr) return MDF_ERR_UNSUPPORTED;
    CUtensorMap xmap;
    const cuuint64_t dims[4] = {(cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)Cfg::D, (cuuint64_t)a.B * a.C};
    const cuuint64_t strides[3] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.H * a.W * 4, (cuuint64_t)Cfg::D * a.H * a.W * 4};
    const cuuint32_t box[4] = {(cuuint32_t)Cfg::BW, (cuuint32_t)Cfg::BH, (cuuint32_t)Cfg::BD, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.x), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return MDF_ERR_CUDA; }
    const dim3 grid((unsigned)((a.W + Cfg::TW - 1) / Cfg::TW), (unsigned)((a.H + Cfg::TH * Cfg::NT - 1) / (Cfg::TH * Cfg::NT)), (unsigned)a.B);
    if (grid.y > 65535u || grid.z > 65535u) return MDF_ERR_UNSUPPORTED;
    auto launch = [&](auto kern) -> int {
        MDF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        kern<<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(xmap, a);
        return launch_status();
    };
    if (fit == 1) return launch(prob_head_tma_kernel<Cfg, 1>);
    if (fit == 2) return launch(prob_head_tma_kernel<Cfg, 2>);
    return launch(prob_head_tma_kernel<Cfg, 0>);
}

template <class Cfg>
static int launch_prob_head(const ProbHeadArgs& a, int fit, cudaStream_t stream)
{
    const dim3 block(32, Cfg::NWD, Cfg::NT);
    const dim3 grid((unsigned)((a.W + Cfg::TW - 1) / Cfg::TW), (unsigned)((a.H + Cfg::TH * Cfg::NT - 1) / (Cfg::TH * Cfg::NT)), (unsigned)a.B);
    if (grid.y > 65535u || grid.z > 65535u) return MDF_ERR_UNSUPPORTED;
    if (fit == 1) prob_head_kernel<Cfg, 1><<<grid, block, 0, stream>>>(a);
    else if (fit == 2) prob_head_kernel<Cfg, 2><<<grid, block, 0, stream>>>(a);
    else prob_head_kernel<Cfg, 0><<<grid, block, 0, stream>>>(a);
    return launch_status();
}

// Scalar fallback of the register-pipelined kernel (W % 4 != 0 or misaligned volume):    D  DSLAB NT PX
using Prob48_s = ProbCfg<48, 8, 1, 1>;
using Prob24_s = ProbCfg<24, 8, 2, 1>;
using Prob8_s = ProbCfg<8, 8, 4, 1>;
// TMA variants (algo k of mdf_prob_head_fwd_ex; 0 = default):   D  DSLAB NT NSTAGE MINB
using Tma48_0 = ProbTmaCfg<48, 4, 1, 2, 2>;       // 12 + 1 warps, 2 x 48 KB stages, 2 CTAs / SM
using Tma48_1 = ProbTmaCfg<48, 8, 1, 2, 2>;       // 6 + 1 warps
using Tma48_2 = ProbTmaCfg<48, 6, 1, 3, 1>;       // 8 + 1 warps, 3 stages, 1 CTA / SM
using Tma24_0 = ProbTmaCfg<24, 6, 2, 2, 2>;       // 4 x 2 + 1 warps, 2 x 41.6 KB stages
using Tma24_1 = ProbTmaCfg<24, 4, 1, 3, 3>;       // 6 + 1 warps, tile 32x4, 3 x 25 KB
using Tma24_2 = ProbTmaCfg<24, 8, 2, 2, 2>;       // 3 x 2 + 1 warps
using Tma8_0 = ProbTmaCfg<8, 8, 4, 2, 3>;         // whole column per thread, 4 + 1 warps, 2 x 23 KB
using Tma8_1 = ProbTmaCfg<8, 8, 8, 2, 2>;         // 8 + 1 warps, tile 32x32
using Tma8_2 = ProbTmaCfg<8, 4, 4, 3, 2>;         // 2 x 4 + 1 warps, 3 stages

}  // namespace mdf

using namespace mdf;

extern "C" {

int mdf_prob_head_fwd_ex(const float* x, const float* prob_weight, const float* depth_hypos, int hypos_per_pixel,
                         int B, int C, int D, int H, int W, float* logits, float* prob, float* depth, float* confidence,
                         int conf_n, int conf_pad_front, int conf_pad_back, int conf_upsample, int curve, float* s,
                         int algo, mdf_stream_t stream_)
{
    if (B < 0 || C < 1 || D < 0 || H < 0 || W < 0) return MDF_ERR_INVALID_SHAPE;
    if (curve < 0 || curve > 2) return MDF_ERR_UNSUPPORTED;
    if (C > kMaxProbChannels) return MDF_ERR_UNSUPPORTED;
    if (D != 8 && D != 24 && D != 48) return MDF_ERR_UNSUPPORTED;          // config.py:199; the column lives in registers
    if ((size_t)B * H * W == 0) return MDF_OK;                             // nothing to do (torch hands out NULL for empty tensors)
    if (!logits && !prob && !depth && !confidence && !s) return MDF_ERR_NULL_POINTER;
    if ((curve != 0) != (s != nullptr)) return MDF_ERR_NULL_POINTER;
    if (confidence) {
        if (conf_n <= 0 || conf_upsample <= 0 || conf_pad_front < 0 || conf_pad_back < 0) return MDF_ERR_INVALID_SHAPE;
        if (D + conf_pad_front + conf_pad_back - conf_n + 1 <= 0) return MDF_ERR_INVALID_SHAPE;
    }
    if (!x || !prob_weight) return MDF_ERR_NULL_POINTER;
    if ((depth || curve != 0) && !depth_hypos) return MDF_ERR_NULL_POINTER;
    const void* out = logits ? (const void*)logits : prob ? (const void*)prob : depth ? (const void*)depth
                      : confidence ? (const void*)confidence : (const void*)s;
    const int dev = device_of(out);
    if (dev < 0) return dev;
    const void* ptrs[8];
    int n = 0;
    ptrs[n++] = x; ptrs[n++] = prob_weight;
    if (depth_hypos) ptrs[n++] = depth_hypos;
    if (logits) ptrs[n++] = logits;
    if (prob) ptrs[n++] = prob;
    if (depth) ptrs[n++] = depth;
    if (confidence) ptrs[n++] = confidence;
    if (s) ptrs[n++] = s;
    const int st = check_on_device(dev, ptrs, n);
    if (st != MDF_OK) return st;
    DeviceGuard guard(dev);
    ProbHeadArgs a;
    a.x = x; a.w = prob_weight; a.hypos = depth_hypos; a.logits = logits; a.prob = prob; a.depth = depth;
    a.conf = confidence; a.s = s; a.per_pixel = hypos_per_pixel; a.B = B; a.C = C; a.H = H; a.W = W;
    a.conf_n = conf_n; a.pad_front = conf_pad_front; a.pad_back = conf_pad_back; a.up = conf_upsample;
    cudaStream_t stream = (cudaStream_t)stream_;
    // 4 pixels per thread need W % 4 == 0 and 16-byte aligned planes; otherwise the scalar variant
    bool vec = W % 4 == 0;
    {
        const void* q[] = {x, logits};
        for (const void* p : q)
            if (p && (reinterpret_cast<uintptr_t>(p) & 15)) vec = false;
    }
    if (D == 8) {
        if (!vec) return launch_prob_head<Prob8_s>(a, curve, stream);
        if (algo == 1) return launch_prob_head_tma<Tma8_1>(a, curve, stream);
        if (algo == 2) return launch_prob_head_tma<Tma8_2>(a, curve, stream);
        return launch_prob_head_tma<Tma8_0>(a, curve, stream);
    }
    if (D == 24) {
        if (!vec) return launch_prob_head<Prob24_s>(a, curve, stream);
        if (algo == 1) return launch_prob_head_tma<Tma24_1>(a, curve, stream);
        if (algo == 2) return launch_prob_head_tma<Tma24_2>(a, curve, stream);
        return launch_prob_head_tma<Tma24_0>(a, curve, stream);
    }
    if (!vec) return launch_prob_head<Prob48_s>(a, curve, stream);
    if (algo == 1) return launch_prob_head_tma<Tma48_1>(a, curve, stream);
    if (algo == 2) return launch_prob_head_tma<Tma48_2>(a, curve, stream);
    return launch_prob_head_tma<Tma48_0>(a, curve, stream);
}

int mdf_prob_head_fwd(const float* x, const float* prob_weight, const float* depth_hypos, int hypos_per_pixel,
                      int B, int C, int D, int H, int W, float* logits, float* prob, float* depth, float* confidence,
                      int conf_n, int conf_pad_front, int conf_pad_back, int conf_upsample, int curve, float* s,
                      mdf_stream_t stream)
{
    return mdf_prob_head_fwd_ex(x, prob_weight, depth_hypos, hypos_per_pixel, B, C, D, H, W, logits, prob, depth, confidence,
                                conf_n, conf_pad_front, conf_pad_back, conf_upsample, curve, s, 0, stream);
}

}  // extern "C"

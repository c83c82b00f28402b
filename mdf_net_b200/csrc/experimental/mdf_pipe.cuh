// mdf_pipe.cuh -- EXPERIMENTAL (tuning builds only, -DMDF_TUNING): a pre-planned box pipeline for the hot kernel.
// Measured SLOWER than cost_volume_staged_kernel (211 / 214 / 188 us vs 187 / 207 / 168 us at BASELINE configs[1]):
// profiles/r02_pipe_*.  Kept so that the measurement can be repeated (tools/tune_pipe.py); no product path uses it.
//
// cost_volume_staged_kernel (mdf_staged.cuh) spends half of its instructions and most of its stall time outside the
// gather: every thread lives for one slab (4-16 samples), and before every source view all warps of the CTA meet
// (sample positions -> bounding box by atomics -> the last warp issues one TMA load -> everybody waits).  This kernel
// keeps the gather core and removes the meetings:
//
//   * a CTA owns a tile of 32 x TH reference pixels and walks RG consecutive ROUNDS of PT*PG depth planes with the
//     same threads (prologue, q / cq tiles, rot|trans in shared memory: once per item, not once per slab);
//   * PLAN: before the first round the CTA computes, for every (round, source view), the bounding box of its sample
//     cells from the round's two boundary planes (a pixel's samples move monotonically along its epipolar line as the
//     depth grows, so the planes in between stay inside) -- one pass, two block barriers per item -- and thread 0
//     turns the boxes into a list of TMA loads ("steps"): one [G/4][BH][BW] float4 box per (round, view), a grid of
//     boxes where a footprint does not fit one (silhouettes, wide search ranges);
//   * PIPELINE: the steps stream through a ring of NS shared-memory slots.  Every warp visits the steps in order:
//     wait for the slot's mbarrier, gather its samples that lie in the box, count itself out; the last warp out
//     re-arms the slot with step n+NS.  No block barrier after the plan, no per-view reduction, no vote at the end:
//     warps drift apart and the gather of one overlaps the positions / epilogue stores of the others;
//   * samples no planned box covers (non-monotone hypotheses, more boxes than the plan holds) are gathered by their
//     warp straight from the difference maps in global memory (L2): the result never depends on the plan.
//
// The arithmetic is that of the staged kernel bit for bit (same positions, taps, blend order, sigmoid, weights).
//
// Reference: net/unit/base.py:85-126 (homo_warping), net/unit/homoaggregate.py:25-46 (+16-20).
#pragma once

#include "../mdf_staged.cuh"

namespace mdf {

//   G      groups: 32 / 16 / 8
//   PT     depth planes a thread holds accumulators for (one round = PT*PG planes)
//   TH     tile height (tile width = one warp = 32 pixels)
//   PG     warps along depth; blockDim = (32, TH, PG)
//   BW,BH  box (texels) of one step
//   NS     slots of the ring
//   CQS    conv_w * q stays in shared memory (G32: 32 more registers would spill)
template <int G_, int PT_, int TH_, int PG_, int BW_, int BH_, int NS_, int MINB_, bool CQS_>
struct PipeCfg {
    static constexpr int G = G_, PT = PT_, TH = TH_, PG = PG_, BW = BW_, BH = BH_, NS = NS_, MINB = MINB_;
    static constexpr bool CQS = CQS_;
    static constexpr int J = G / 4;
    static constexpr int NCQ = CQS ? 1 : G;
    static constexpr int THREADS = 32 * TH * PG;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int RS = PT * PG;                           // planes per round
    static constexpr int MAXR = 4;                               // rounds per item (RG <= MAXR)
    static constexpr int MAXRV = 48;                             // (round, view) pairs per item (RG * V <= MAXRV)
    static constexpr int MAXSTEPS = 96;
    static constexpr int MAXBOXES = 8;                           // boxes per (round, view)
    static constexpr int PLANE_BYTES = BW * BH * 16;
    static constexpr int BOX_BYTES = J * PLANE_BYTES;            // multiple of 128
    static constexpr int TILE_BYTES = J * TH * 32 * 16;
    static constexpr int OFF_Q = NS * BOX_BYTES;
    static constexpr int OFF_CQ = OFF_Q + TILE_BYTES;
    static constexpr int OFF_RT = OFF_CQ + TILE_BYTES;           // [V][12] floats, 16-byte aligned rows of 48 bytes
    static constexpr int OFF_BAR = OFF_RT + kMaxSrcViews * 48;   // NS full barriers, tile barrier, plan barrier
    static constexpr int OFF_DONE = OFF_BAR + 8 * (NS + 2);      // NS counters
    static constexpr int OFF_NSTEPS = OFF_DONE + 4 * NS;
    static constexpr int OFF_BBOX = (OFF_NSTEPS + 4 + 15) / 16 * 16;             // [r * V + v] int4 (minx, miny, maxx, maxy)
    static constexpr int OFF_META = OFF_BBOX + MAXRV * 16;                       // [r * V + v] first step | boxes << 16
    static constexpr int OFF_PLAN = (OFF_META + MAXRV * 4 + 15) / 16 * 16;       // [MAXSTEPS] int4 (ox, oy, plane0, -)
    static constexpr size_t SMEM = OFF_PLAN + MAXSTEPS * 16 + 128 /*alignment slack*/;
    static_assert(BW * 2 <= 256 && BH <= 256, "TMA box dimensions are limited to 256 elements");
    static_assert(PT <= 8, "plane bookkeeping uses 8 bits per plane");
    static_assert(BOX_BYTES % 128 == 0, "slots stay 128-byte aligned");
};

struct PipeArgs {
    const float* rt;      // [V][B][12]
    const float* dwp;     // folded depth_weight
    const float* hypos;
    const float4* S4;     // [V][B][J][H][W] float4 difference maps (global fallback)
    float* out;           // (B,G,D,H,W)
    GridNorm gn;
    int per_pixel, V, B, D, H, W, tiles_x, tiles_y, rgroups, RG;
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB)
cost_volume_pipe_kernel(const __grid_constant__ StagedMaps maps, const PipeArgs a)
{
    constexpr int G = Cfg::G, J = Cfg::J, PT = Cfg::PT, TH = Cfg::TH, PG = Cfg::PG, BW = Cfg::BW, BH = Cfg::BH;
    constexpr int NS = Cfg::NS, NCQ = Cfg::NCQ, RS = Cfg::RS, WARPS = Cfg::WARPS, THREADS = Cfg::THREADS;
    constexpr int PLANE = Cfg::PLANE_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t pad = (128u - (smem_u32(smem_raw) & 127u)) & 127u;
    uint8_t* const sm = smem_raw + pad;
    const uint32_t box0 = smem_u32(sm);
    const uint32_t bar_full = box0 + Cfg::OFF_BAR, bar_tile = bar_full + 8u * NS, bar_plan = bar_tile + 8u;
    const uint32_t rt_s = box0 + Cfg::OFF_RT;
    int* const done = reinterpret_cast<int*>(sm + Cfg::OFF_DONE);
    volatile int* const nsteps_p = reinterpret_cast<volatile int*>(sm + Cfg::OFF_NSTEPS);
    int* const bbox = reinterpret_cast<int*>(sm + Cfg::OFF_BBOX);
    volatile int* const meta = reinterpret_cast<volatile int*>(sm + Cfg::OFF_META);
    volatile int4* const plan = reinterpret_cast<volatile int4*>(sm + Cfg::OFF_PLAN);

    const int lane = threadIdx.x, ty = threadIdx.y, pg = threadIdx.z;
    const int tid = lane + 32 * (ty + TH * pg);
    const uint32_t q_s = box0 + Cfg::OFF_Q + (uint32_t)(ty * 32 + lane) * 16u;      // [j][ty][lane] float4
    const uint32_t cq_s = box0 + Cfg::OFF_CQ + (uint32_t)(ty * 32 + lane) * 16u;
    constexpr uint32_t TJ = TH * 32 * 16;

    int it = blockIdx.x;
    const int tile_x = it % a.tiles_x; it /= a.tiles_x;
    const int tile_y = it % a.tiles_y; it /= a.tiles_y;
    const int rgroup = it % a.rgroups;
    const int b = it / a.rgroups;

    const int H = a.H, W = a.W, D = a.D, V = a.V;
    const int px = tile_x * 32 + lane, py = tile_y * TH + ty;
    const bool pix_ok = (px < W) && (py < H);
    const size_t HW = (size_t)H * W;
    const int rounds_total = (D + RS - 1) / RS;
    const int round0 = rgroup * a.RG;
    const int nrounds = min(a.RG, rounds_total - round0);
    GridNormFast gf;
    gf.g = a.gn;
    gf.r_half_wm1 = refine_rcp(a.gn.half_wm1);
    gf.r_half_hm1 = refine_rcp(a.gn.half_hm1);
    const bool tiny_map = !(a.gn.half_wm1 > 0.25f && a.gn.half_hm1 > 0.25f);   // W or H == 1: the normalisation divides by 0

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS + 2; ++s) mbar_init(bar_full + 8u * s, 1);
        fence_barrier_init();
        mbar_expect_tx(bar_tile, 2 * Cfg::TILE_BYTES);
        tma_load_3d(box0 + Cfg::OFF_Q, &maps.q4, bar_tile, tile_x * 64, tile_y * TH, b * J);
        tma_load_3d(box0 + Cfg::OFF_CQ, &maps.cq4, bar_tile, tile_x * 64, tile_y * TH, b * J);
    }
    if (tid < NS) done[tid] = 0;
    for (int k = tid; k < nrounds * V; k += THREADS) {
        int* e = bbox + 4 * k;
        e[0] = kNone; e[1] = kNone; e[2] = -kNone; e[3] = -kNone;
    }
    for (int k = tid; k < V * 12; k += THREADS) {
        const float val = __ldg(a.rt + ((size_t)(k / 12) * a.B + b) * 12 + (k % 12));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(rt_s + 4u * k), "f"(val) : "memory");
    }
    __syncthreads();

    // rot | trans of view v: 3 x LDS.128 (broadcast)
    auto load_rt = [&](int v, float (&rt)[12]) {
        const float4 r0 = lds128(rt_s + (uint32_t)v * 48u), r1 = lds128(rt_s + (uint32_t)v * 48u + 16u), r2 = lds128(rt_s + (uint32_t)v * 48u + 32u);
        rt[0] = r0.x; rt[1] = r0.y; rt[2] = r0.z; rt[3] = r0.w; rt[4] = r1.x; rt[5] = r1.y; rt[6] = r1.z; rt[7] = r1.w;
        rt[8] = r2.x; rt[9] = r2.y; rt[10] = r2.z; rt[11] = r2.w;
    };
    auto load_depth = [&](int d) -> float {
        return a.per_pixel ? __ldg(a.hypos + ((size_t)b * D + d) * HW + (size_t)py * W + px) : __ldg(a.hypos + (size_t)b * D + d);
    };

    // ---- PLAN, part 1: bounding boxes of the sample cells of every (round, view) from the boundary planes
    //      round0*RS + k*RS, k = 0..nrounds (the last one clamped to D-1); warps along depth share the boundaries ----
    for (int k = pg; k <= nrounds; k += PG) {
        const int d = min(D - 1, (round0 + k) * RS);
        const float depth = pix_ok ? load_depth(d) : 0.0f;
        for (int v = 0; v < V; ++v) {
            float rt[12];
            load_rt(v, rt);
            const RotXYZ r = rot_xyz(rt, (float)px, (float)py);
            float ix, iy;
            if (tiny_map || !sample_position_try(r, rt, depth, gf, ix, iy)) sample_position(r, rt, depth, gf.g, ix, iy);
            // cells of interest are -1 .. size-1; positions outside (or non-finite) are clamped onto that range: a pixel
            // whose boundary sample is outside may still enter the image between the boundaries
            ix = fminf(fmaxf(ix, -1.0f), a.gn.fw - 1.0f);
            iy = fminf(fmaxf(iy, -1.0f), a.gn.fh - 1.0f);
            float f; int x0, y0;
            floor_small(ix, f, x0);
            floor_small(iy, f, y0);
            const int lox = __reduce_min_sync(0xffffffffu, pix_ok ? x0 : kNone), loy = __reduce_min_sync(0xffffffffu, pix_ok ? y0 : kNone);
            const int hix = __reduce_max_sync(0xffffffffu, pix_ok ? x0 : -kNone), hiy = __reduce_max_sync(0xffffffffu, pix_ok ? y0 : -kNone);
            if (lane == 0 && lox != kNone) {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {          // boundary k closes round k-1 and opens round k
                    const int r_ = k - 1 + rr;
                    if (r_ < 0 || r_ >= nrounds) continue;
                    int* e = bbox + 4 * (r_ * V + v);
                    atomicMin(e + 0, lox); atomicMin(e + 1, loy); atomicMax(e + 2, hix); atomicMax(e + 3, hiy);
                }
            }
        }
    }
    __syncthreads();

    // ---- PLAN, part 2: thread 0 lays the boxes out as steps and starts the first NS loads ----
    auto issue = [&](int m) {
        const int4 e = make_int4(plan[m].x, plan[m].y, plan[m].z, 0);
        const uint32_t s = (uint32_t)(m % NS);
        mbar_expect_tx(bar_full + 8u * s, Cfg::BOX_BYTES);
        tma_load_3d(box0 + s * Cfg::BOX_BYTES, &maps.s4, bar_full + 8u * s, e.x * 2, e.y, e.z);   // the maps count 8-byte elements
    };
    if (tid == 0) {
        int n = 0;
        for (int r = 0; r < nrounds; ++r)
            for (int v = 0; v < V; ++v) {
                const int* e = bbox + 4 * (r * V + v);
                int cnt = 0;
                if (e[0] != kNone) {
                    const int nbx = (e[2] - e[0]) / (BW - 1) + 1, nby = (e[3] - e[1]) / (BH - 1) + 1;
                    for (int jy = 0; jy < nby; ++jy)
                        for (int jx = 0; jx < nbx; ++jx) {
                            if (cnt >= Cfg::MAXBOXES || n + cnt >= Cfg::MAXSTEPS) break;      // the rest goes through global memory
                            plan[n + cnt].x = e[0] + jx * (BW - 1);
                            plan[n + cnt].y = e[1] + jy * (BH - 1);
                            plan[n + cnt].z = (v * a.B + b) * J;
                            ++cnt;
                        }
                }
                meta[r * V + v] = n | (cnt << 16);
                n += cnt;
            }
        *nsteps_p = n;
        for (int m = 0; m < NS && m < n; ++m) issue(m);
        mbar_arrive(bar_plan);
    }

    const float alpha = __ldg(a.dwp + 0), betap = __ldg(a.dwp + 1);
    const float fcw = __ldg(a.dwp + 2), fcb = __ldg(a.dwp + 3);
    float depth[PT], ndepth[PT];
    auto fetch_depths = [&](int r, float (&dst)[PT]) {
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            const int d = (round0 + r) * RS + pg * PT + i;
            dst[i] = (pix_ok && d < D && r < nrounds) ? load_depth(d) : 0.0f;
        }
    };
    fetch_depths(0, ndepth);

    float2 acc[PT][G / 2];
    float wsum[PT];
    uint64_t n_void = 0;
    uint32_t ok_mask = 0;

    auto positions = [&](const float (&rt)[12], float (&ix)[PT], float (&iy)[PT]) -> uint32_t {
        uint32_t todo = 0;
        const RotXYZ r = rot_xyz(rt, (float)px, (float)py);
        bool exact = !tiny_map;
#pragma unroll
        for (int i = 0; i < PT; ++i) exact = sample_position_try(r, rt, depth[i], gf, ix[i], iy[i]) && exact;
        if (__any_sync(0xffffffffu, !exact)) {
#pragma unroll
            for (int i = 0; i < PT; ++i) sample_position(r, rt, depth[i], gf.g, ix[i], iy[i]);
        }
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            const bool inside = (ix[i] > -1.0f) && (ix[i] < a.gn.fw) && (iy[i] > -1.0f) && (iy[i] < a.gn.fh);
            if ((ok_mask >> i) & 1u) {
                if (inside) todo |= 1u << i;
                else n_void += 1ull << (8 * i);
            }
        }
        return todo;
    };

    float cq[NCQ];
    float ksum = 0.0f;
    // one sample: blend -> sigmoid per group -> z -> view weight -> accumulate.  `tap(j, k)` returns tap k (nw, ne, sw, se)
    // of float4 plane j.
    auto sample = [&](int i, float ixv, float iyv, float fx0, float fy0, auto tap) {
        const float ax = __fsub_rn(__fadd_rn(fx0, 1.0f), ixv), bx = __fsub_rn(ixv, fx0);
        const float ay = __fsub_rn(__fadd_rn(fy0, 1.0f), iyv), by = __fsub_rn(iyv, fy0);
        const float wnw = __fmul_rn(ax, ay), wne = __fmul_rn(bx, ay), wsw = __fmul_rn(ax, by), wse = __fmul_rn(bx, by);
        const float2 Wnw = make_float2(wnw, wnw), Wne = make_float2(wne, wne);
        const float2 Wsw = make_float2(wsw, wsw), Wse = make_float2(wse, wse);
        float2 p[G / 2];
        float2 z2 = make_float2(-ksum, 0.0f);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 nw = tap(j, 0), ne = tap(j, 1), sw = tap(j, 2), se = tap(j, 3);
            float4 c;
            if (Cfg::CQS) c = lds128(cq_s + (uint32_t)j * TJ);
            else c = make_float4(cq[(4 * j + 0) % NCQ], cq[(4 * j + 1) % NCQ], cq[(4 * j + 2) % NCQ], cq[(4 * j + 3) % NCQ]);
            // bilinear blend, two groups per instruction; per component the reference's order
            // fma(se,wse, fma(sw,wsw, fma(ne,wne, nw*wnw)))
            float2 t01 = __fmul2_rn(make_float2(nw.x, nw.y), Wnw), t23 = __fmul2_rn(make_float2(nw.z, nw.w), Wnw);
            t01 = __ffma2_rn(make_float2(ne.x, ne.y), Wne, t01); t23 = __ffma2_rn(make_float2(ne.z, ne.w), Wne, t23);
            t01 = __ffma2_rn(make_float2(sw.x, sw.y), Wsw, t01); t23 = __ffma2_rn(make_float2(sw.z, sw.w), Wsw, t23);
            t01 = __ffma2_rn(make_float2(se.x, se.y), Wse, t01); t23 = __ffma2_rn(make_float2(se.z, se.w), Wse, t23);
            // sigmoid(a-b) = 1 / (1 + 2^t), t capped so that (1+2^t0)(1+2^t1) cannot become inf * 0; one MUFU.RCP
            // serves two groups: r = 1/(u0*u1), 1/u0 = r*u1, 1/u1 = r*u0.
            const float2 one2 = make_float2(1.0f, 1.0f);
            const float2 u01 = __fadd2_rn(make_float2(ex2_approx(fminf(t01.x, 62.0f)), ex2_approx(fminf(t01.y, 62.0f))), one2);
            const float2 u23 = __fadd2_rn(make_float2(ex2_approx(fminf(t23.x, 62.0f)), ex2_approx(fminf(t23.y, 62.0f))), one2);
            const float r01 = rcp_approx(u01.x * u01.y), r23 = rcp_approx(u23.x * u23.y);
            const float2 p01 = __fmul2_rn(make_float2(r01, r01), make_float2(u01.y, u01.x));
            const float2 p23 = __fmul2_rn(make_float2(r23, r23), make_float2(u23.y, u23.x));
            p[2 * j] = p01; p[2 * j + 1] = p23;
            z2 = __ffma2_rn(make_float2(c.x, c.y), p01, z2);
            z2 = __ffma2_rn(make_float2(c.z, c.w), p23, z2);
        }
        const float z = z2.x + z2.y;
        float h = fmaf(z, alpha, betap);              // BatchNorm3d (eval fold)
        h = fmaxf(h, 0.0f);                           // ReLU
        h = fmaf(h, fcw, fcb);                        // Conv3d(1,1,1)
        const float w = rcp_approx(1.0f + ex2_approx(-kLog2e * h));   // Sigmoid
        wsum[i] += w;
        const float2 w2 = make_float2(w, w);
#pragma unroll
        for (int g = 0; g < G / 2; ++g) acc[i][g] = __ffma2_rn(w2, p[g], acc[i][g]);
    };

    // gather the samples of `todo` that lie inside the box with origin (ox, oy); returns the rest
    auto gather = [&](uint32_t box, int ox, int oy, const float (&ix)[PT], const float (&iy)[PT], uint32_t todo) -> uint32_t {
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            if (!((todo >> i) & 1u)) continue;
            float fx0, fy0;
            int x0, y0;
            floor_small(ix[i], fx0, x0);
            floor_small(iy[i], fy0, y0);
            const int rx = x0 - ox, ry = y0 - oy;
            if ((unsigned)rx >= (unsigned)(BW - 1) || (unsigned)ry >= (unsigned)(BH - 1)) continue;   // not in this box
            todo &= ~(1u << i);
            const uint32_t addr = box + (uint32_t)(ry * BW + rx) * 16u;
            sample(i, ix[i], iy[i], fx0, fy0, [&](int j, int k) -> float4 {
                return lds128(addr + j * PLANE + (k & 1) * 16 + (k >> 1) * BW * 16);       // constant offsets -> LDS.128 [R + imm]
            });
        }
        return todo;
    };
    // the same from global memory (L2): samples outside every planned box
    auto gather_global = [&](int v, const float (&ix)[PT], const float (&iy)[PT], uint32_t todo) {
        const float4* __restrict__ base = a.S4 + (size_t)(v * a.B + b) * J * HW;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            if (!((todo >> i) & 1u)) continue;
            float fx0, fy0;
            int x0, y0;
            floor_small(ix[i], fx0, x0);
            floor_small(iy[i], fy0, y0);
            const bool xl = x0 >= 0, xr = x0 + 1 < W, yt = y0 >= 0, yb = y0 + 1 < H;
            const float4* __restrict__ p00 = base + (ptrdiff_t)max(y0, 0) * W + max(x0, 0);
            const int dx = (xl && xr) ? 1 : 0, dy = (yt && yb) ? W : 0;
            sample(i, ix[i], iy[i], fx0, fy0, [&](int j, int k) -> float4 {
                const bool okk = ((k & 1) ? xr : xl) && ((k >> 1) ? yb : yt);
                const float4* q = p00 + (size_t)j * HW + ((k & 1) ? dx : 0) + ((k >> 1) ? dy : 0);
                return okk ? __ldg(q) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            });
        }
    };

    mbar_wait(bar_tile, 0);                      // q / cq tiles have landed
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
    mbar_wait(bar_plan, 0);                      // the plan is laid out
    const int nsteps = *nsteps_p;
    const float w_void = __ldg(a.dwp + 5);       // view weight of a sample with no tap in bounds (similarity 0.5)

    // ---- PIPELINE ----
    int n = 0;                                   // next step (warp uniform; every warp visits every step)
    uint32_t slot = 0, par = 0;                  // n % NS, (n / NS) & 1
    for (int r = 0; r < nrounds; ++r) {
        const int d0 = (round0 + r) * RS + pg * PT;
        ok_mask = 0;
        n_void = 0;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            depth[i] = ndepth[i];
            if (pix_ok && d0 + i < D) ok_mask |= 1u << i;
            wsum[i] = 0.0f;
#pragma unroll
            for (int g = 0; g < G / 2; ++g) acc[i][g] = make_float2(0.0f, 0.0f);
        }
        fetch_depths(r + 1, ndepth);             // in flight during this round
        for (int v = 0; v < V; ++v) {
            float ix[PT], iy[PT];
            uint32_t todo;
            {
                float rt[12];
                load_rt(v, rt);
                todo = positions(rt, ix, iy);
            }
            const int m = meta[r * V + v];
            const int last = (m & 0xffff) + (m >> 16);
            for (; n < last; ++n) {
                // Every warp waits for every step, with or without samples in it: the wait is what keeps a warp from
                // counting itself out of step n + NS before step n has been counted out by all (the counters run on).
                const int ox = plan[n].x, oy = plan[n].y;
                mbar_wait(bar_full + 8u * slot, par);
                if (__any_sync(0xffffffffu, todo != 0u)) todo = gather(box0 + slot * Cfg::BOX_BYTES, ox, oy, ix, iy, todo);
                // count this warp out of the slot; the last one out re-arms it with step n + NS
                __syncwarp();
                if (lane == 0 && n + NS < nsteps) {
                    if ((atomicAdd(done + slot, 1) + 1) % WARPS == 0) issue(n + NS);
                }
                if (++slot == NS) { slot = 0; par ^= 1u; }
            }
            if (__any_sync(0xffffffffu, todo != 0u)) gather_global(v, ix, iy, todo);
        }
        // ---- volume_sum / weight_sum (homoaggregate.py:46), coalesced 128-byte rows ----
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            if (!((ok_mask >> i) & 1u)) continue;
            const float nv = (float)((unsigned)(n_void >> (8 * i)) & 255u);
            const float wv_sum = nv * w_void;
            const float ws = fmaf(nv, w_void, wsum[i]);
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
                op[0] = o01.x; op[gstride] = o01.y; op[2 * gstride] = o23.x; op[3 * gstride] = o23.y;
                op += 4 * gstride;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <class Cfg>
static int launch_pipe(const StagedArgs& sa, const StagedBuffers& buf, int rg_override, cudaStream_t stream, const HotEvents& ev)
{
    StagedMaps maps;
    int st = encode_planes(&maps.s4, buf.S4, sa.W, sa.H, (long long)sa.V * sa.B * Cfg::J, Cfg::BW, Cfg::BH, Cfg::J);
    if (st == MDF_OK) st = encode_planes(&maps.q4, buf.Q4, sa.W, sa.H, (long long)sa.B * Cfg::J, 32, Cfg::TH, Cfg::J);
    if (st == MDF_OK) st = encode_planes(&maps.cq4, buf.CQ4, sa.W, sa.H, (long long)sa.B * Cfg::J, 32, Cfg::TH, Cfg::J);
    if (st != MDF_OK) return st;
    auto kern = cost_volume_pipe_kernel<Cfg>;
    MDF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    PipeArgs a;
    a.rt = sa.rt; a.dwp = sa.dwp; a.hypos = sa.hypos; a.S4 = reinterpret_cast<const float4*>(buf.S4); a.out = sa.out;
    a.per_pixel = sa.per_pixel; a.V = sa.V; a.B = sa.B; a.D = sa.D; a.H = sa.H; a.W = sa.W;
    a.gn = make_grid_norm(a.H, a.W);
    a.tiles_x = (a.W + 31) / 32;
    a.tiles_y = (a.H + Cfg::TH - 1) / Cfg::TH;
    const int rounds = (a.D + Cfg::RS - 1) / Cfg::RS;
    const long long tiles = (long long)a.tiles_x * a.tiles_y * a.B;
    // rounds per item: as many as the plan holds (one box per (round, view) plus headroom), but keep >= 6 waves of items
    int rg = min(min(Cfg::MAXR, rounds), max(1, Cfg::MAXRV / a.V));
    while (rg > 1 && tiles * ((rounds + rg - 1) / rg) < 148LL * Cfg::MINB * 6) --rg;
    if (rg_override > 0) rg = min(min(min(rg_override, Cfg::MAXR), rounds), max(1, Cfg::MAXRV / a.V));
    a.RG = rg;
    a.rgroups = (rounds + rg - 1) / rg;
    const long long items = tiles * a.rgroups;
    if (items <= 0) return MDF_OK;
    if (items > INT_MAX) return MDF_ERR_UNSUPPORTED;
    if (ev.start) cudaEventRecord(ev.start, stream);
    kern<<<(unsigned)items, dim3(32, Cfg::TH, Cfg::PG), Cfg::SMEM, stream>>>(maps, a);
    if (ev.stop) cudaEventRecord(ev.stop, stream);
    return launch_status();
}

//                       G  PT TH PG  BW  BH NS MINB CQS
using PipeG32_0 = PipeCfg<32, 1, 2, 4, 40, 6, 3, 2, true>;     // 3 x 30 KiB slots + 2 x 8 KiB tiles
using PipeG32_1 = PipeCfg<32, 1, 2, 4, 48, 6, 2, 2, true>;     // 2 x 36 KiB slots
using PipeG32_2 = PipeCfg<32, 1, 4, 2, 40, 7, 2, 2, true>;     // tile 32x4: 2 x 35 KiB slots + 2 x 16 KiB tiles
using PipeG16_0 = PipeCfg<16, 2, 2, 4, 64, 6, 4, 2, false>;    // 4 x 24 KiB slots
using PipeG16_1 = PipeCfg<16, 2, 2, 4, 48, 6, 5, 2, false>;    // 5 x 18 KiB slots
using PipeG16_2 = PipeCfg<16, 2, 4, 2, 44, 8, 4, 2, false>;    // tile 32x4: 4 x 22 KiB slots
using PipeG8_0  = PipeCfg<8, 4, 4, 2, 64, 10, 4, 2, false>;    // 4 x 20 KiB slots
using PipeG8_1  = PipeCfg<8, 4, 4, 2, 56, 8, 6, 2, false>;     // 6 x 14 KiB slots
using PipeG8_2  = PipeCfg<8, 4, 8, 1, 56, 14, 3, 2, false>;    // tile 32x8, 4 planes per round: 3 x 24.5 KiB slots

static int launch_pipe_variant(int G, int variant, int rg, const StagedArgs& a, const StagedBuffers& S, cudaStream_t stream, const HotEvents& ev)
{
    if (G == 32) {
        switch (variant) {
            case 0: return launch_pipe<PipeG32_0>(a, S, rg, stream, ev);
            case 1: return launch_pipe<PipeG32_1>(a, S, rg, stream, ev);
            case 2: return launch_pipe<PipeG32_2>(a, S, rg, stream, ev);
        }
    } else if (G == 16) {
        switch (variant) {
            case 0: return launch_pipe<PipeG16_0>(a, S, rg, stream, ev);
            case 1: return launch_pipe<PipeG16_1>(a, S, rg, stream, ev);
            case 2: return launch_pipe<PipeG16_2>(a, S, rg, stream, ev);
        }
    } else if (G == 8) {
        switch (variant) {
            case 0: return launch_pipe<PipeG8_0>(a, S, rg, stream, ev);
            case 1: return launch_pipe<PipeG8_1>(a, S, rg, stream, ev);
            case 2: return launch_pipe<PipeG8_2>(a, S, rg, stream, ev);
        }
    }
    return MDF_ERR_UNSUPPORTED;
}

}  // namespace mdf

// mdf_backward.cu -- train-mode forward and the backward pass of the fused cost volume (C/G == 2).
//
// Reference: VectorAggregate.forward under autograd (net/unit/homoaggregate.py:25-46), with
// depth_weight = Conv3d(G,1,k=1) - BatchNorm3d(1) - ReLU - Conv3d(1,1,k=1) - Sigmoid (:16-20; base.py:50-68).
// Gradients flow to every feature map and to the 5 depth_weight parameters; the sampling grid is built
// under no_grad (base.py:97), so projections and hypotheses get none.  In train mode BatchNorm3d uses the
// statistics of z_v over (B,D,H,W), separately for every source view (the module is called once per view),
// which makes forward and backward two-phase: a statistics sweep, then the sweep that uses them.
//
// Notation per (b,d,y,x) and source view v:
//   t_vg  = bilinear sample of S_v (S = (f[2g+1]-f[2g])*log2e)      p_vg = 1/(1+2^t_vg)
//   sim_vg = 0.5 + q_g (p_vg - 0.5)      z_v = sum_g cw_g sim_vg     h_v = a_v z_v + b_v   (BatchNorm)
//   w_v = sigmoid(fcw*relu(h_v) + fcb)   out_g = sum_v w_v sim_vg / sum_v w_v
//
// The train-mode FORWARD runs the tuned TMA-staged kernel of mdf_staged.cuh twice: in its statistics mode (MODE 1: sum z,
// sum z^2 per source view) and, after the per-view BatchNorm folds (bn_fold_kernel), in its per-view-fold mode (MODE 2).
// The BACKWARD:
//   phase 1a  the same staged gather once more (MODE 3): its accumulator registers hold gout_g q_g, and every (element, view)
//             leaves z_v and A'_v = sum_g gout_g sim_vg behind, every element go = sum_g gout_g out_g;
//   phase 1b  bwd_finalize_kernel, one thread per element: what needs all views of an element (dh_v, w_v / sum w), the batch
//             sums of train-mode BatchNorm's backward, d fc;
//   phase 2   bwd_sweep_kernel, one thread per (pixel, view, slice of 8 groups), walks the depth planes with the taps through
//             L1/L2 and scatters the feature gradient with 128-bit vector reductions (red.global.add.v4.f32) into a
//             difference-gradient map dS4 -- half the atomics of scattering into both channels of a pair, and only when the
//             sample leaves its source cell (the scatter is what bounds the backward);
//   finish    dS4 / dQ4 -> NCHW feature gradients.
// The forward's batch statistics come back in through the ABI, so the backward does not repeat the statistics sweep.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdf_common.cuh"
#include "mdf_host.cuh"
#include "mdf_setup.cuh"
#include "mdf_staged.cuh"   // FeaPtrs, prep_kernel

namespace mdf {

constexpr float kLn2 = 0.6931471805599453f;

struct TrainArgs {
    const float4* S4;     // [V][B][J][H][W]
    const float4* Q4;     // [B][J][H][W]
    const float* rt;      // [V][B][12]
    const float* cw;      // (G,)
    const float* bnv;     // [V][4]: a_v, b_v (h = a z + b), 1/sqrt(var+eps), mean   (written by bn_fold_kernel)
    const float* fc;      // [2]: fcw, fcb  (device copies)
    const float* hypos;
    int per_pixel, V, B, D, H, W;
};

// 1/(1+2^t) with the same MUFU approximations (<= 2 ulp each) as the forward kernel it differentiates
__device__ __forceinline__ float sigm2(float t) { return rcp_approx(1.0f + ex2_approx(fminf(t, 126.0f))); }

// ---- BatchNorm constants per view (1 thread) ---------------------------------------------------------
// training: batch statistics from `stats`; else the running statistics.  Also publishes the batch mean and
// the unbiased variance (momentum update of the running statistics happens on the host side of the ABI).
// `stats` come from the staged kernel's statistics pass: sums of ITS z = sum_g cw_g q_g (p_g - 0.5) = true z - hcw with
// hcw = 0.5 * sum_g cw_g (the variance does not see the shift, the mean gets it back here).  `vparams` [V][4] are the
// folds that kernel's forward pass uses: alpha_v, beta_v + alpha_v * hcw, the weight of an out-of-image sample, hcw.
__global__ void bn_fold_kernel(const double* __restrict__ stats, double count, int V, int training,
                               const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                               const float* __restrict__ bn_mean, const float* __restrict__ bn_var, float eps,
                               const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                               const float* __restrict__ cw, int G,
                               float* __restrict__ bnv, float* __restrict__ fc, float* __restrict__ vparams,
                               float* __restrict__ batch_stats, const float* __restrict__ saved_stats)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    fc[0] = fc_w[0]; fc[1] = fc_b[0];
    double hcw = 0.0;
    for (int g = 0; g < G; ++g) hcw += (double)cw[g];
    hcw *= 0.5;
    for (int v = 0; v < V; ++v) {
        double mean = bn_mean[0], var = bn_var[0];
        if (training && saved_stats) {
            // the forward's batch statistics (mean, unbiased variance), handed back by the caller: no second statistics sweep
            mean = (double)saved_stats[2 * v];
            var = (double)saved_stats[2 * v + 1] * (count > 1.0 ? (count - 1.0) / count : 1.0);
        } else if (training) {
            const double mk = stats[2 * v] / count;
            mean = mk + hcw;
            var = stats[2 * v + 1] / count - mk * mk;
            if (var < 0.0) var = 0.0;
            if (batch_stats && training && !saved_stats) {
                batch_stats[2 * v] = (float)mean;
                batch_stats[2 * v + 1] = (float)(count > 1.0 ? var * count / (count - 1.0) : var);
            }
        }
        const double invstd = 1.0 / sqrt(var + (double)eps);
        const double alpha = invstd * (double)bn_w[0];
        bnv[4 * v + 0] = (float)alpha;
        bnv[4 * v + 1] = (float)((double)bn_b[0] - mean * alpha);
        bnv[4 * v + 2] = (float)invstd;
        bnv[4 * v + 3] = (float)mean;
        const float betap = (float)((double)bnv[4 * v + 1] + alpha * hcw);
        const float hv = fmaf(fmaxf(betap, 0.0f), fc_w[0], fc_b[0]);
        vparams[4 * v + 0] = (float)alpha;
        vparams[4 * v + 1] = betap;
        vparams[4 * v + 2] = 1.0f / (1.0f + expf(-hv));
        vparams[4 * v + 3] = (float)hcw;
    }
}

// ---- backward ----------------------------------------------------------------------------------------
struct BwdArgs {
    TrainArgs t;
    const float* out;       // saved forward output (B,G,D,H,W)
    const float* gout;      // upstream gradient      (B,G,D,H,W)
    const double* bsum;     // [V][2] train: sum dh_v, sum dh_v*zhat_v  (phase 1 result)
    float* za;              // [V][3][B*D*H*W]: dh_v, z_v, w_v / sum w of every element, written by phase 1, read by phase 2
    double count;
    int training;
    float4* dS4;            // [V][B][J][H][W]  zero-initialised
    float4* dQ4;            // [B][J][H][W]     zero-initialised
    double* gparam;         // [4 + G]: d bn_w, d bn_b, d fc_w, d fc_b, d cw[G]   zero-initialised
};

// dL/dh_v of one element given the per-view forward values
__device__ __forceinline__ float dh_of(const BwdArgs& a, float aprime, float go, float rwsum, float w, float h, float* dact_out)
{
    const float dw = (aprime - go) * rwsum;             // d out_g / d w_v = (sim_vg - out_g) / wsum
    const float dact = dw * w * (1.0f - w);             // sigmoid
    if (dact_out) *dact_out = dact;
    return h > 0.0f ? dact * __ldg(a.t.fc) : 0.0f;      // Conv3d(1,1,1) then ReLU
}

// phase 1b: the staged gather (mdf_staged.cuh, MODE 3) left A'_v and its z_v (= true z - hcw) of every (element, view) in za
// and go = sum_g gout_g out_g; one thread per element turns them into what the sweep wants -- dh_v, the true z_v, w_v / sum w --
// and reduces what only needs per-element values: the batch sums sum dh_v, sum dh_v * zhat_v per view (train-mode BatchNorm
// backward; d gamma / d beta in both modes) and d fc.weight, d fc.bias.
__global__ void __launch_bounds__(256)
bwd_finalize_kernel(const BwdArgs a, const float* __restrict__ vparams, const float* __restrict__ go_all, double* __restrict__ bsum)
{
    const TrainArgs& t = a.t;
    const size_t total = (size_t)t.B * t.D * t.H * t.W, idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = idx < total;
    const size_t i = ok ? idx : 0;
    const float fcw = __ldg(t.fc), fcb = __ldg(t.fc + 1);
    // the view weight exactly as the forward kernel evaluates it (mdf_staged.cuh: BatchNorm fold, ReLU, Conv3d(1,1,1), sigmoid on
    // the MUFU): the backward differentiates the function the forward computed
    auto weight_of = [&](int v, float zs, float& h) {
        h = fmaf(zs, __ldg(vparams + 4 * v), __ldg(vparams + 4 * v + 1));
        return rcp_approx(1.0f + ex2_approx(-kLog2e * fmaf(fmaxf(h, 0.0f), fcw, fcb)));
    };
    float wsum = 0.0f;
    {
        const float* zp = a.za + total + i;          // slot 3v+1 of view v
        for (int v = 0; v < t.V; ++v, zp += 3 * total) {
            float h;
            wsum += weight_of(v, *zp, h);
        }
    }
    const float rws = __frcp_rn(wsum);
    const float go = __ldg(go_all + i);
    // the sums: a warp's 32 values in float, across warps and blocks in double; one row per warp in shared memory (plain
    // read-modify-write by lane 0: a shared-memory double atomicAdd is a CAS spin loop), ONE block-wide hand-over at the end
    constexpr int NW = 8, NS = 2 * kMaxSrcViews + 2;
    __shared__ double red[NW][NS];
    for (int k = threadIdx.x; k < NW * NS; k += blockDim.x) (&red[0][0])[k] = 0.0;
    __syncthreads();
    double* red_w = red[threadIdx.x >> 5];
    auto warp_add = [&](float val, int slot) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if ((threadIdx.x & 31) == 0) red_w[slot] += (double)val;
    };
    float gfc0 = 0.0f, gfc1 = 0.0f;
    float* slot = a.za + i;                          // (A'_v, z_v, -) of view v -> (dh_v, true z_v, w_v / sum w)
    for (int v = 0; v < t.V; ++v, slot += 3 * total) {
        float s0 = 0.0f, s1 = 0.0f;
        if (ok) {
            const float zs = slot[total], aprime = slot[0];
            float h, dact;
            const float w = weight_of(v, zs, h);
            const float dh = dh_of(a, aprime, go, rws, w, h, &dact);
            const float z = zs + __ldg(vparams + 4 * v + 3);
            const float zhat = (z - __ldg(t.bnv + 4 * v + 3)) * __ldg(t.bnv + 4 * v + 2);
            s0 = dh; s1 = dh * zhat;
            gfc0 = fmaf(dact, fmaxf(h, 0.0f), gfc0);
            gfc1 += dact;
            slot[0] = dh;
            slot[total] = z;
            slot[2 * total] = w * rws;
        }
        warp_add(s0, 2 * v);
        warp_add(s1, 2 * v + 1);
    }
    warp_add(gfc0, 2 * t.V);
    warp_add(gfc1, 2 * t.V + 1);
    __syncthreads();
    if ((int)threadIdx.x < 2 * t.V + 2) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) sum += red[w][threadIdx.x];
        if (sum != 0.0) atomicAdd((int)threadIdx.x < 2 * t.V ? bsum + threadIdx.x : a.gparam + 2 + (threadIdx.x - 2 * t.V), sum);
    }
}

// predicated 16-byte read-only load at p + OFF bytes (zero when the predicate is off): keeps ONE address register pair per
// row of the cell instead of an index -> pointer computation per tap
template <int OFF>
__device__ __forceinline__ float4 ldg4_if(const float4* p, bool pred)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4+%6];\n\t}"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) : "l"(p), "r"((int)pred), "n"(OFF));
    return v;
}

template <int OFF>
__device__ __forceinline__ void red_add_v4_if(float4* addr, bool pred, float a, float b, float c, float d)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q red.relaxed.gpu.global.add.v4.f32 [%0+%6], {%1, %2, %3, %4};\n\t}"
                 ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "r"((int)pred), "n"(OFF) : "memory");
}

__device__ __forceinline__ void red_add_v4(float4* addr, float4 v)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// phase 2: one thread = one (pixel, source view, slice of GS groups) walking ALL depth planes.
// Why: the scatter of the tap gradients bounds the backward (340 M 16-byte reductions at the BlendedMVS train shape, stage 0:
// the LSU needs ~1.3 cycles per lane for them) -- and consecutive planes of a pixel hit the same or the neighbouring source
// cell (0.1-0.2 px per plane at stages 1-2, ~1 px at stage 0).  So the four tap gradients of the current cell stay in
// registers while the planes are walked and go out as vector reductions only when the cell changes; when it moves by one
// column the two shared taps stay.  d q accumulates over the planes too (one reduction per thread instead of one per element),
// d conv.weight is reduced once per thread.  What a thread cannot see -- the other views of an element -- comes from phase 1
// (dh_v, z_v, w_v / sum w).
template <int G>
__global__ void __launch_bounds__(256, 2)
bwd_sweep_kernel(const BwdArgs a)
{
    constexpr int J = G / 4, GS = 8, JS = GS / 4, NS = G / GS;
    const TrainArgs& t = a.t;
    const size_t HW = (size_t)t.H * t.W;
    const int s = blockIdx.y % NS, v = (blockIdx.y / NS) % t.V, b = blockIdx.y / (NS * t.V);
    const size_t pix0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = pix0 < HW;                      // idle lanes of the last block shadow the last pixel and contribute nothing
    const size_t pix = live ? pix0 : HW - 1;
    const int x = (int)(pix % t.W), y = (int)(pix / t.W);
    const GridNormFast gn = make_grid_norm_fast(t.H, t.W);
    const size_t total = (size_t)t.B * t.D * HW;
    const size_t gstride = (size_t)t.D * HW;
    __shared__ float dcw_s[8][GS];                    // one row per warp (plain stores; a shared-memory double atomicAdd is a CAS spin loop)
    __shared__ float cw_s[GS];                        // the slice's conv weights: read per use (8 registers the plane loop needs more)
    if (threadIdx.x < GS) cw_s[threadIdx.x] = __ldg(t.cw + s * GS + threadIdx.x);
    __syncthreads();
    float rt[12];                                     // registers: the reductions below clobber memory, a pointer would be re-read per plane
#pragma unroll
    for (int k = 0; k < 12; ++k) rt[k] = __ldg(t.rt + ((size_t)v * t.B + b) * 12 + k);
    const RotXYZ r = rot_xyz(rt, (float)x, (float)y);
    float q[GS];
#pragma unroll
    for (int jj = 0; jj < JS; ++jj) {
        const float4 qq = __ldg(t.Q4 + ((size_t)b * J + s * JS + jj) * HW + pix);
        q[4 * jj] = qq.x; q[4 * jj + 1] = qq.y; q[4 * jj + 2] = qq.z; q[4 * jj + 3] = qq.w;
    }
    // dz = alpha (dh - m1 - zhat m2), zhat = (z - mean) invstd  (train-mode BatchNorm backward; eval: m1 = m2 = 0)  =  kA dh + kB z + kC
    float kA, kB, kC;
    {
        const float invstd = __ldg(t.bnv + 4 * v + 2), mean = __ldg(t.bnv + 4 * v + 3), alpha = __ldg(t.bnv + 4 * v);
        const float m1 = a.training ? (float)(a.bsum[2 * v] / a.count) : 0.0f, m2 = a.training ? (float)(a.bsum[2 * v + 1] / a.count) : 0.0f;
        kA = alpha; kB = -alpha * m2 * invstd; kC = -alpha * m1 - kB * mean;
    }
    const float4* Sv = t.S4 + (((size_t)v * t.B + b) * J + s * JS) * HW;
    float4* dSv = a.dS4 + (((size_t)v * t.B + b) * J + s * JS) * HW;
    float dq[GS], dcw[GS];
    float gnw[GS], gne[GS], gsw[GS], gse[GS];         // tap gradients of the current cell (cx, cy)
#pragma unroll
    for (int k = 0; k < GS; ++k) { dq[k] = 0.0f; dcw[k] = 0.0f; gnw[k] = gne[k] = gsw[k] = gse[k] = 0.0f; }
    int cx = INT_MIN, cy = INT_MIN;
    // flush `west` / `east` columns of the cell (cx, cy) with bounds (zero padding: out-of-image taps have no gradient)
    auto flush = [&](bool west, bool east) {
        const bool y0in = live && (unsigned)cy < (unsigned)t.H, y1in = live && (unsigned)(cy + 1) < (unsigned)t.H;
        const bool x0in = (unsigned)cx < (unsigned)t.W, x1in = (unsigned)(cx + 1) < (unsigned)t.W;
        float4* dn = dSv + ((ptrdiff_t)cy * t.W + cx);          // one address per row of the cell, predicated reductions
        float4* ds = dn + t.W;
#pragma unroll
        for (int jj = 0; jj < JS; ++jj, dn += HW, ds += HW) {
            if (west) red_add_v4_if<0>(dn, x0in && y0in, gnw[4 * jj], gnw[4 * jj + 1], gnw[4 * jj + 2], gnw[4 * jj + 3]);
            if (west) red_add_v4_if<0>(ds, x0in && y1in, gsw[4 * jj], gsw[4 * jj + 1], gsw[4 * jj + 2], gsw[4 * jj + 3]);
            if (east) red_add_v4_if<16>(dn, x1in && y0in, gne[4 * jj], gne[4 * jj + 1], gne[4 * jj + 2], gne[4 * jj + 3]);
            if (east) red_add_v4_if<16>(ds, x1in && y1in, gse[4 * jj], gse[4 * jj + 1], gse[4 * jj + 2], gse[4 * jj + 3]);
        }
    };
    // running pointers of the per-plane operands (one 64-bit add per plane instead of the index arithmetic)
    const float* zap = a.za + (size_t)(3 * v) * total + (size_t)b * t.D * HW + pix;
    const float* hyp = t.per_pixel ? t.hypos + (size_t)b * t.D * HW + pix : t.hypos + (size_t)b * t.D;
    const size_t hstep = t.per_pixel ? HW : 1;
    const float* gp = a.gout + (size_t)b * G * gstride + pix + (size_t)(s * GS) * gstride;
    for (int d = 0; d < t.D; ++d, zap += HW, hyp += hstep, gp += HW) {
        const float dh = __ldg(zap), z = __ldg(zap + total), wn = __ldg(zap + 2 * total);
        const float dz = fmaf(kA, dh, fmaf(kB, z, kC));
        const float depth = __ldg(hyp);
        float ix, iy;
        sample_position_fast(r, rt, depth, gn, ix, iy);
        const Taps tp = make_taps(ix, iy, gn.g);
        if (tp.valid && (tp.x0 != cx || tp.y0 != cy)) {
            // the cell moved: write out what leaves the 2x2 footprint, keep the column that stays
            if (tp.y0 == cy && tp.x0 == cx + 1) {
                flush(true, false);
#pragma unroll
                for (int k = 0; k < GS; ++k) { gnw[k] = gne[k]; gsw[k] = gse[k]; gne[k] = 0.0f; gse[k] = 0.0f; }
            } else if (tp.y0 == cy && tp.x0 == cx - 1) {
                flush(false, true);
#pragma unroll
                for (int k = 0; k < GS; ++k) { gne[k] = gnw[k]; gse[k] = gsw[k]; gnw[k] = 0.0f; gsw[k] = 0.0f; }
            } else {
                if (cx != INT_MIN) flush(true, true);
#pragma unroll
                for (int k = 0; k < GS; ++k) gnw[k] = gne[k] = gsw[k] = gse[k] = 0.0f;
            }
            cx = tp.x0; cy = tp.y0;
        }
        // the four taps of the cell: one address per plane, the bounds once (zero padding: a tap outside reads as 0)
        const float4* pn = Sv + ((ptrdiff_t)tp.y0 * t.W + tp.x0);
        const float4* ps = pn + t.W;
        const bool x0in = tp.valid && (unsigned)tp.x0 < (unsigned)t.W, x1in = tp.valid && (unsigned)(tp.x0 + 1) < (unsigned)t.W;
        const bool y0in = (unsigned)tp.y0 < (unsigned)t.H, y1in = (unsigned)(tp.y0 + 1) < (unsigned)t.H;
        const bool inw = x0in && y0in, ine = x1in && y0in, isw = x0in && y1in, ise = x1in && y1in;
#pragma unroll
        for (int jj = 0; jj < JS; ++jj, pn += HW, ps += HW) {
            const float4 nw = ldg4_if<0>(pn, inw), ne = ldg4_if<16>(pn, ine);
            const float4 sw = ldg4_if<0>(ps, isw), se = ldg4_if<16>(ps, ise);
            const float tv[4] = {blend4(nw.x, ne.x, sw.x, se.x, tp), blend4(nw.y, ne.y, sw.y, se.y, tp),
                                 blend4(nw.z, ne.z, sw.z, se.z, tp), blend4(nw.w, ne.w, sw.w, se.w, tp)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int gl = 4 * jj + k;
                const float p = sigm2(tv[k]);
                const float sim = fmaf(q[gl], p - 0.5f, 0.5f);
                const float dsim = fmaf(__ldg(gp + (size_t)gl * gstride), wn, dz * cw_s[gl]);
                dcw[gl] = fmaf(dz, sim, dcw[gl]);
                dq[gl] = fmaf(dsim, p - 0.5f, dq[gl]);
                const float dt = -kLn2 * p * (1.0f - p) * (dsim * q[gl]);      // dp/dt = -ln2 p (1-p)
                if (tp.valid) {
                    gnw[gl] = fmaf(dt, tp.wnw, gnw[gl]); gne[gl] = fmaf(dt, tp.wne, gne[gl]);
                    gsw[gl] = fmaf(dt, tp.wsw, gsw[gl]); gse[gl] = fmaf(dt, tp.wse, gse[gl]);
                }
            }
        }
    }
    if (cx != INT_MIN) flush(true, true);
    if (live) {
        float4* dqp = a.dQ4 + ((size_t)b * J + s * JS) * HW + pix;
#pragma unroll
        for (int jj = 0; jj < JS; ++jj)
            red_add_v4(dqp + (size_t)jj * HW, make_float4(dq[4 * jj], dq[4 * jj + 1], dq[4 * jj + 2], dq[4 * jj + 3]));
    }
    // d conv.weight: warp shuffle -> one row per warp in shared memory -> one global (double) atomic per block and group
#pragma unroll
    for (int k = 0; k < GS; ++k) {
        float c = live ? dcw[k] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((threadIdx.x & 31) == 0) dcw_s[threadIdx.x >> 5][k] = c;
    }
    __syncthreads();
    if (threadIdx.x < GS) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += (double)dcw_s[w][threadIdx.x];
        if (sum != 0.0) atomicAdd(a.gparam + 4 + s * GS + threadIdx.x, sum);
    }
}

// dS4 / dQ4 -> NCHW feature gradients.  One thread per pixel per view; blockIdx.y = view * B + b.
struct GradPtrs { float* p[MDF_MAX_VIEWS]; };

__global__ void __launch_bounds__(256)
bwd_finish_kernel(GradPtrs grads, int B, int G, int HW, const float4* __restrict__ Q4, const float4* __restrict__ dQ4,
                  const float4* __restrict__ dS4)
{
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= HW) return;
    const int v = blockIdx.y / B, b = blockIdx.y % B;
    float* __restrict__ g = grads.p[v];
    if (g == nullptr) return;
    g += (size_t)b * 2 * G * HW + pix;
    const int J = G / 4;
    for (int j = 0; j < J; ++j) {
        float d[4];
        if (v == 0) {
            // q = 2*sigmoid(r0 - r1) - 1:  dq/dr0 = (1 - q^2)/2 = -dq/dr1
            const float4 q = __ldg(Q4 + ((size_t)b * J + j) * HW + pix), dq = __ldg(dQ4 + ((size_t)b * J + j) * HW + pix);
            d[0] = dq.x * 0.5f * (1.0f - q.x * q.x); d[1] = dq.y * 0.5f * (1.0f - q.y * q.y);
            d[2] = dq.z * 0.5f * (1.0f - q.z * q.z); d[3] = dq.w * 0.5f * (1.0f - q.w * q.w);
#pragma unroll
            for (int k = 0; k < 4; ++k) { g[(size_t)(8 * j + 2 * k) * HW] = d[k]; g[(size_t)(8 * j + 2 * k + 1) * HW] = -d[k]; }
        } else {
            // S = (f[2g+1] - f[2g]) * log2e
            const float4 ds = __ldg(dS4 + (((size_t)(v - 1) * B + b) * J + j) * HW + pix);
            d[0] = ds.x * kLog2e; d[1] = ds.y * kLog2e; d[2] = ds.z * kLog2e; d[3] = ds.w * kLog2e;
#pragma unroll
            for (int k = 0; k < 4; ++k) { g[(size_t)(8 * j + 2 * k) * HW] = -d[k]; g[(size_t)(8 * j + 2 * k + 1) * HW] = d[k]; }
        }
    }
}

// d bn.weight = sum_v sum dh_v * zhat_v, d bn.bias = sum_v sum dh_v (the per-view batch sums of phase 1); the rest is in place
__global__ void gparam_to_float_kernel(const double* __restrict__ src, const double* __restrict__ bsum, int V, float* __restrict__ dst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double val = src[i];
    if (i < 2) {
        val = 0.0;
        for (int v = 0; v < V; ++v) val += bsum[2 * v + (1 - i)];
    }
    dst[i] = (float)val;
}

// BatchNorm3d's running statistics after a train-mode forward: the module is applied once per source view, so the reference
// performs V momentum updates in view order (torch.nn.BatchNorm3d: running = (1 - f) running + f batch, unbiased variance;
// f = momentum, or 1 / num_batches_tracked when momentum is None).  One thread; in place.
__global__ void bn_running_update_kernel(const float* __restrict__ batch_stats, int V, float momentum, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, long long* __restrict__ num_batches_tracked)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float rm = running_mean[0], rv = running_var[0];
    long long n = num_batches_tracked ? num_batches_tracked[0] : 0;
    for (int v = 0; v < V; ++v) {
        ++n;
        const float f = momentum >= 0.0f ? momentum : (float)(1.0 / (double)n);
        rm = fmaf(f, batch_stats[2 * v] - rm, rm);
        rv = fmaf(f, batch_stats[2 * v + 1] - rv, rv);
    }
    running_mean[0] = rm; running_var[0] = rv;
    if (num_batches_tracked) num_batches_tracked[0] = n;
}

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
struct TrainWorkspace {
    size_t rt, dwp, q, s, cq, bnv, vparams, fc, za, go, stats, bsum, gparam, dq, ds, total;
};

static TrainWorkspace make_train_workspace(int B, int N, int G, int D, int H, int W)
{
    TrainWorkspace w;
    const size_t V = (size_t)(N - 1), plane = (size_t)B * G * H * W * sizeof(float);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    w.rt = take(V * B * 12 * sizeof(float));
    w.dwp = take(64 * sizeof(float));
    w.q = take(plane);
    w.s = take(V * plane);
    w.cq = take(plane);
    w.bnv = take(kMaxSrcViews * 4 * sizeof(float));
    w.vparams = take(kMaxSrcViews * 4 * sizeof(float));
    w.fc = take(2 * sizeof(float));
    w.za = take(V * 3 * (size_t)B * D * H * W * sizeof(float));      // dh_v, z_v, w_v / sum w of every element (backward phase 1 -> 2)
    w.go = take((size_t)B * D * H * W * sizeof(float));               // sum_g gout_g out_g of every element (phase 1a -> 1b)
    w.stats = take(kMaxSrcViews * 2 * sizeof(double));      // stats | bsum | gparam are contiguous: one memset
    w.bsum = take(kMaxSrcViews * 2 * sizeof(double));
    w.gparam = take((4 + 32) * sizeof(double));
    w.dq = take(plane);
    w.ds = take(V * plane);
    w.total = off;
    return w;
}

struct TrainCall {
    const float* const* features; int N; const float* ref_proj; const float* const* src_projs;
    const float* hypos; int per_pixel;
    const float *conv_w, *bn_w, *bn_b, *bn_mean, *bn_var; float bn_eps; const float *fc_w, *fc_b;
    int training, B, C, G, D, H, W;
};

static int validate(const TrainCall& c, const void* out, void* workspace, size_t workspace_bytes, TrainWorkspace* ws, int* dev_out)
{
    if (c.B < 0 || c.C <= 0 || c.G <= 0 || c.D < 0 || c.H < 0 || c.W < 0 || c.N < 2 || c.C % c.G != 0) return MDF_ERR_INVALID_SHAPE;
    if (c.N > MDF_MAX_VIEWS || c.C != 2 * c.G || !(c.G == 8 || c.G == 16 || c.G == 32)) return MDF_ERR_UNSUPPORTED;
    if ((long long)c.H * c.W > INT_MAX - 256 || (long long)c.N * c.B > 65535 || (long long)(c.N - 1) * c.B > 256) return MDF_ERR_UNSUPPORTED;
    if (!c.features || !c.src_projs || !c.ref_proj || !c.hypos || !c.conv_w || !c.bn_w || !c.bn_b || !c.bn_mean || !c.bn_var ||
        !c.fc_w || !c.fc_b || !out)
        return MDF_ERR_NULL_POINTER;
    *ws = make_train_workspace(c.B, c.N, c.G, c.D, c.H, c.W);
    if (!workspace || ((uintptr_t)workspace & 255) != 0 || workspace_bytes < ws->total) return MDF_ERR_WORKSPACE;
    const int dev = device_of(out);
    if (dev < 0) return dev;
    const void* ptrs[MDF_MAX_VIEWS * 2 + 16];
    int n = 0;
    for (int i = 0; i < c.N; ++i) ptrs[n++] = c.features[i];
    for (int i = 0; i < c.N - 1; ++i) ptrs[n++] = c.src_projs[i];
    const void* more[] = {c.ref_proj, c.hypos, c.conv_w, c.bn_w, c.bn_b, c.bn_mean, c.bn_var, c.fc_w, c.fc_b, workspace};
    for (const void* p : more) ptrs[n++] = p;
    const int st = check_on_device(dev, ptrs, n);
    if (st != MDF_OK) return st;
    *dev_out = dev;
    return MDF_OK;
}

// arguments of the TMA-staged kernel (mdf_staged.cuh) over this call's workspace
static StagedArgs staged_args(const TrainCall& c, uint8_t* wsb, const TrainWorkspace& ws, float* out)
{
    StagedArgs a;
    a.rt = reinterpret_cast<const float*>(wsb + ws.rt); a.dwp = reinterpret_cast<const float*>(wsb + ws.dwp);
    a.vparams = reinterpret_cast<const float*>(wsb + ws.vparams); a.stats = reinterpret_cast<double*>(wsb + ws.stats);
    a.hypos = c.hypos; a.out = out;
    a.per_pixel = c.per_pixel; a.V = c.N - 1; a.B = c.B; a.D = c.D; a.H = c.H; a.W = c.W;
    a.gn = make_grid_norm(c.H, c.W);
    a.tiles_x = a.tiles_y = a.slabs = 0;
    return a;
}

static StagedBuffers staged_buffers(uint8_t* wsb, const TrainWorkspace& ws)
{
    StagedBuffers b;
    b.S4 = reinterpret_cast<const float*>(wsb + ws.s); b.Q4 = reinterpret_cast<const float*>(wsb + ws.q);
    b.CQ4 = reinterpret_cast<const float*>(wsb + ws.cq);
    return b;
}

// prep (S4, Q4, rt) + BatchNorm constants; returns the TrainArgs for the sweeps
template <int G>
static int prepare(const TrainCall& c, uint8_t* wsb, const TrainWorkspace& ws, float* batch_stats, const float* saved_stats, cudaStream_t stream, TrainArgs* out)
{
    const int V = c.N - 1;
    float* rt = reinterpret_cast<float*>(wsb + ws.rt);
    float* dwp = reinterpret_cast<float*>(wsb + ws.dwp);
    float4* Q4 = reinterpret_cast<float4*>(wsb + ws.q);
    float4* S4 = reinterpret_cast<float4*>(wsb + ws.s);
    FeaPtrs fp;
    for (int i = 0; i < MDF_MAX_VIEWS; ++i) fp.p[i] = i < c.N ? c.features[i] : nullptr;
    PrepSetup su;
    for (int v = 0; v < kMaxSrcViews; ++v) su.src_projs.p[v] = v < V ? c.src_projs[v] : nullptr;
    su.ref_proj = c.ref_proj; su.V = V; su.rt = rt; su.dwp = dwp;
    su.dw = {c.conv_w, c.bn_w, c.bn_b, c.bn_mean, c.bn_var, c.fc_w, c.fc_b, c.bn_eps};
    const int HW = c.H * c.W;
    int st = launch_setup_and_prep(su, fp, c.N, c.B, G, HW, Q4, reinterpret_cast<float4*>(wsb + ws.cq), S4, stream);
    if (st != MDF_OK) return st;
    // stats, bsum and gparam are adjacent
    MDF_CUDA_TRY(cudaMemsetAsync(wsb + ws.stats, 0, ws.dq - ws.stats, stream));

    TrainArgs a;
    a.S4 = S4; a.Q4 = Q4; a.rt = rt; a.cw = c.conv_w;
    a.bnv = reinterpret_cast<float*>(wsb + ws.bnv); a.fc = reinterpret_cast<float*>(wsb + ws.fc);
    a.hypos = c.hypos; a.per_pixel = c.per_pixel; a.V = V; a.B = c.B; a.D = c.D; a.H = c.H; a.W = c.W;
    const size_t total = (size_t)c.B * c.D * c.H * c.W;
    double* stats = reinterpret_cast<double*>(wsb + ws.stats);
    if (c.training && !saved_stats) {
        // batch statistics of z per source view: the TMA-staged gather in its statistics mode (mdf_staged.cuh, MODE 1)
        st = launch_staged_train<1>(G, staged_args(c, wsb, ws, nullptr), staged_buffers(wsb, ws), stream);
        if (st != MDF_OK) return st;
    }
    bn_fold_kernel<<<1, 32, 0, stream>>>(stats, (double)total, V, c.training, c.bn_w, c.bn_b, c.bn_mean, c.bn_var, c.bn_eps,
                                         c.fc_w, c.fc_b, c.conv_w, G, reinterpret_cast<float*>(wsb + ws.bnv),
                                         reinterpret_cast<float*>(wsb + ws.fc), reinterpret_cast<float*>(wsb + ws.vparams), batch_stats, saved_stats);
    st = launch_status();
    if (st != MDF_OK) return st;
    *out = a;
    return MDF_OK;
}

template <int G>
static int train_fwd(const TrainCall& c, float* cost_volume, float* batch_stats, uint8_t* wsb, const TrainWorkspace& ws, cudaStream_t stream)
{
    TrainArgs a;
    int st = prepare<G>(c, wsb, ws, batch_stats, nullptr, stream, &a);
    if (st != MDF_OK) return st;
    // the forward itself: the TMA-staged kernel with per-view BatchNorm folds (MODE 2)
    return launch_staged_train<2>(G, staged_args(c, wsb, ws, cost_volume), staged_buffers(wsb, ws), stream);
}

template <int G>
static int train_bwd(const TrainCall& c, const float* cost_volume, const float* grad_out, const float* saved_stats, float* const* grad_features,
                     float* grad_params, uint8_t* wsb, const TrainWorkspace& ws, cudaStream_t stream)
{
    BwdArgs a;
    int st = prepare<G>(c, wsb, ws, nullptr, saved_stats, stream, &a.t);
    if (st != MDF_OK) return st;
    a.out = cost_volume; a.gout = grad_out;
    a.bsum = reinterpret_cast<double*>(wsb + ws.bsum);
    a.count = (double)c.B * c.D * c.H * c.W;
    a.training = c.training;
    a.dS4 = reinterpret_cast<float4*>(wsb + ws.ds); a.dQ4 = reinterpret_cast<float4*>(wsb + ws.dq);
    a.gparam = reinterpret_cast<double*>(wsb + ws.gparam);
    MDF_CUDA_TRY(cudaMemsetAsync(wsb + ws.dq, 0, ws.total - ws.dq, stream));
    const size_t total = (size_t)c.B * c.D * c.H * c.W;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    a.za = reinterpret_cast<float*>(wsb + ws.za);
    {
        // phase 1 on the TMA-staged gather (MODE 3 of the hot kernel), then the per-element pass
        StagedArgs sa = staged_args(c, wsb, ws, nullptr);
        sa.gout = grad_out; sa.fwd_out = cost_volume; sa.za = a.za; sa.go = reinterpret_cast<float*>(wsb + ws.go);
        st = launch_staged_train<3>(G, sa, staged_buffers(wsb, ws), stream);
        if (st != MDF_OK) return st;
        bwd_finalize_kernel<<<blocks, 256, 0, stream>>>(a, reinterpret_cast<const float*>(wsb + ws.vparams), sa.go,
                                                        reinterpret_cast<double*>(wsb + ws.bsum));
        st = launch_status();
        if (st != MDF_OK) return st;
    }
    {
        const int HWi = c.H * c.W;
        bwd_sweep_kernel<G><<<dim3((unsigned)((HWi + 255) / 256), (unsigned)(c.B * (c.N - 1) * (G / 8))), 256, 0, stream>>>(a);
        st = launch_status();
        if (st != MDF_OK) return st;
    }
    GradPtrs gp;
    for (int i = 0; i < MDF_MAX_VIEWS; ++i) gp.p[i] = (grad_features && i < c.N) ? grad_features[i] : nullptr;
    const int HW = c.H * c.W;
    bwd_finish_kernel<<<dim3((unsigned)((HW + 255) / 256), (unsigned)(c.N * c.B)), 256, 0, stream>>>(
        gp, c.B, G, HW, a.t.Q4, a.dQ4, a.dS4);
    st = launch_status();
    if (st != MDF_OK) return st;
    if (grad_params) {
        gparam_to_float_kernel<<<1, 64, 0, stream>>>(a.gparam, a.bsum, c.N - 1, grad_params, 4 + G);
        st = launch_status();
    }
    return st;
}

}  // namespace mdf

using namespace mdf;

extern "C" {

size_t mdf_cost_volume_train_workspace_bytes(int B, int N, int C, int G, int D, int H, int W)
{
    if (B <= 0 || N < 2 || C != 2 * G || G <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
    return make_train_workspace(B, N, G, D, H, W).total;
}

int mdf_cost_volume_train_fwd(const float* const* features, int N, const float* ref_proj, const float* const* src_projs,
                              const float* depth_hypos, int hypos_per_pixel, const float* conv_weight, const float* bn_weight,
                              const float* bn_bias, const float* bn_mean, const float* bn_var, float bn_eps,
                              const float* fc_weight, const float* fc_bias, int training, int B, int C, int G, int D, int H,
                              int W, float* cost_volume, float* batch_stats, void* workspace, size_t workspace_bytes,
                              mdf_stream_t stream)
{
    const TrainCall c = {features, N, ref_proj, src_projs, depth_hypos, hypos_per_pixel, conv_weight, bn_weight, bn_bias, bn_mean,
                         bn_var, bn_eps, fc_weight, fc_bias, training, B, C, G, D, H, W};
    if (B >= 0 && D >= 0 && H >= 0 && W >= 0 && C > 0 && G > 0 && N >= 2 && C % G == 0 && (size_t)B * D * H * W == 0) return MDF_OK;
    TrainWorkspace ws;
    int dev = 0;
    int st = validate(c, cost_volume, workspace, workspace_bytes, &ws, &dev);
    if (st != MDF_OK) return st;
    if (training && batch_stats && device_of(batch_stats) != dev) return MDF_ERR_NOT_DEVICE;
    DeviceGuard guard(dev);
    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    cudaStream_t s = (cudaStream_t)stream;
    if (G == 32) return train_fwd<32>(c, cost_volume, batch_stats, wsb, ws, s);
    if (G == 16) return train_fwd<16>(c, cost_volume, batch_stats, wsb, ws, s);
    return train_fwd<8>(c, cost_volume, batch_stats, wsb, ws, s);
}

int mdf_bn_running_update(const float* batch_stats, int V, float momentum, float* running_mean, float* running_var,
                          long long* num_batches_tracked, mdf_stream_t stream)
{
    if (V < 0 || V > kMaxSrcViews) return MDF_ERR_INVALID_SHAPE;
    if (V == 0) return MDF_OK;
    if (!batch_stats || !running_mean || !running_var) return MDF_ERR_NULL_POINTER;
    const int dev = device_of(running_mean);
    if (dev < 0) return dev;
    const void* ptrs[4] = {batch_stats, running_mean, running_var, num_batches_tracked};
    const int st = check_on_device(dev, ptrs, num_batches_tracked ? 4 : 3);
    if (st != MDF_OK) return st;
    DeviceGuard guard(dev);
    bn_running_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(batch_stats, V, momentum, running_mean, running_var, num_batches_tracked);
    return launch_status();
}

int mdf_cost_volume_bwd(const float* const* features, int N, const float* ref_proj, const float* const* src_projs,
                        const float* depth_hypos, int hypos_per_pixel, const float* conv_weight, const float* bn_weight,
                        const float* bn_bias, const float* bn_mean, const float* bn_var, float bn_eps, const float* fc_weight,
                        const float* fc_bias, int training, int B, int C, int G, int D, int H, int W, const float* cost_volume,
                        const float* grad_out, const float* batch_stats, float* const* grad_features, float* grad_params,
                        void* workspace, size_t workspace_bytes, mdf_stream_t stream)
{
    const TrainCall c = {features, N, ref_proj, src_projs, depth_hypos, hypos_per_pixel, conv_weight, bn_weight, bn_bias, bn_mean,
                         bn_var, bn_eps, fc_weight, fc_bias, training, B, C, G, D, H, W};
    if (!grad_out) return MDF_ERR_NULL_POINTER;
    TrainWorkspace ws;
    int dev = 0;
    int st = validate(c, cost_volume, workspace, workspace_bytes, &ws, &dev);
    if (st != MDF_OK) return st;
    if ((size_t)B * D * H * W == 0) return MDF_ERR_INVALID_SHAPE;
    {
        const void* ptrs[MDF_MAX_VIEWS + 3];
        int n = 0;
        ptrs[n++] = grad_out;
        if (batch_stats) ptrs[n++] = batch_stats;
        if (grad_params) ptrs[n++] = grad_params;
        if (grad_features)
            for (int i = 0; i < N; ++i)
                if (grad_features[i]) ptrs[n++] = grad_features[i];
        st = check_on_device(dev, ptrs, n);
        if (st != MDF_OK) return st;
    }
    DeviceGuard guard(dev);
    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    cudaStream_t s = (cudaStream_t)stream;
    if (G == 32) return train_bwd<32>(c, cost_volume, grad_out, training ? batch_stats : nullptr, grad_features, grad_params, wsb, ws, s);
    if (G == 16) return train_bwd<16>(c, cost_volume, grad_out, training ? batch_stats : nullptr, grad_features, grad_params, wsb, ws, s);
    return train_bwd<8>(c, cost_volume, grad_out, training ? batch_stats : nullptr, grad_features, grad_params, wsb, ws, s);
}

}  // extern "C"

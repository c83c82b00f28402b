// mdf_head.cu -- depth-regression / photometric-confidence head of MDF-Net for sm_100a.
//
// Replaces, with one launch per stage, what the reference does with ~12 ATen launches:
//   F.softmax(x, dim=1)          net/unit/regular.py:67-69,130-133   (tail of the 3-D regulariser)
//   depth_regression             net/unit/regress.py:5-7
//   confidence_regress           net/unit/regress.py:9-25  (+ nearest x2 upsample, net/core.py:75-77)
//
// Two kernels: softmax_regress_reg_kernel<D> for the configured depths 8 / 24 / 48 (column in registers) and the
// generic softmax_regress_kernel<DS> below for any other D.
//
// Layout: logits / prob are (B,D,H,W) with W fastest, so for a fixed depth plane consecutive lanes
// read consecutive pixels.  A warp owns 32/DS consecutive pixels and DS interleaved slices of the
// depth axis (lane = slice * (32/DS) + pixel): every load instruction covers DS full 32-byte
// sectors, and the max / sum / expectation over D finish with warp-shuffle reductions across the
// DS slices.  DS = 4 is used for the deep stages (D = 48, 24: few pixels, long columns), DS = 1
// whenever the confidence is requested: trunc(sum_d p_d * d) is a discrete decision
// (regress.py:15-18) and must be summed in the reference's sequential order.
//
// HBM traffic is the algorithmic minimum: logits are read from DRAM once (the second and third
// sweep of a column hit L1: a CTA's working set is 256 * D * 4 B <= 48 KiB), prob / depth /
// confidence are written once.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdf_common.cuh"
#include "mdf_host.cuh"
#include "mdf_tail.cuh"

namespace mdf {

struct HeadArgs {
    const float* logits;   // (B,D,H,W) logits (softmax path) or probabilities (regression-only paths)
    const float* hypos;    // (B,D) or (B,D,H,W)
    float* prob;           // (B,D,H,W) or nullptr
    float* depth;          // (B,H,W)   or nullptr
    float* conf;           // (B,H*up,W*up) or nullptr
    float* s;              // (B,H,W) fitted scale of the column (HyposByFit) or nullptr
    int per_pixel, B, D, H, W;
    int conf_n, pad_front, pad_back, up;
};

// ------------------------------------------------------------------------------------------------
// fused softmax + expectation (+ confidence).  DS = depth slices per warp (1, 2 or 4).
// ------------------------------------------------------------------------------------------------
template <int DS, int FIT>
__global__ void __launch_bounds__(256)
softmax_regress_kernel(const HeadArgs a)
{
    constexpr int PW = 32 / DS;                    // pixels per warp
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    const int slice = lane / PW;
    const size_t HW = (size_t)a.H * a.W;
    const size_t pix = (size_t)warp * PW + (lane % PW);      // over B*H*W
    const bool ok = pix < (size_t)a.B * HW;
    const int b = ok ? (int)(pix / HW) : 0;
    const size_t p = ok ? pix % HW : 0;
    const int D = a.D;
    const int dq = (D + DS - 1) / DS;              // planes per slice
    const int d_lo = slice * dq, d_hi = min(D, d_lo + dq);
    const float* __restrict__ col = a.logits + (size_t)b * D * HW + p;

    // sweep 1: max over D
    float m = -INFINITY;
    if (ok)
        for (int d = d_lo; d < d_hi; ++d) m = fmaxf(m, __ldg(col + (size_t)d * HW));
#pragma unroll
    for (int o = PW; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));

    // sweep 2: sum of exp(x - max), sequential in d inside a slice (ATen's order when DS == 1)
    float sum = 0.0f;
    if (ok)
        for (int d = d_lo; d < d_hi; ++d) sum = __fadd_rn(sum, expf(__fsub_rn(__ldg(col + (size_t)d * HW), m)));
#pragma unroll
    for (int o = PW; o < 32; o <<= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));

    // sweep 3: probabilities, expectation of depth and of the plane index
    float acc = 0.0f, eidx = 0.0f;
    if (ok) {
        float* __restrict__ pcol = a.prob ? a.prob + (size_t)b * D * HW + p : nullptr;
        const float* __restrict__ hcol = a.per_pixel ? a.hypos + (size_t)b * D * HW + p : a.hypos + (size_t)b * D;
        const size_t hstride = a.per_pixel ? HW : 1;
        for (int d = d_lo; d < d_hi; ++d) {
            const float pr = __fdiv_rn(expf(__fsub_rn(__ldg(col + (size_t)d * HW), m)), sum);
            if (pcol) pcol[(size_t)d * HW] = pr;
            acc = __fadd_rn(acc, __fmul_rn(pr, __ldg(hcol + (size_t)d * hstride)));    // regress.py:7
            eidx = __fadd_rn(eidx, __fmul_rn(pr, (float)d));                           // regress.py:15-17
        }
    }
#pragma unroll
    for (int o = PW; o < 32; o <<= 1) {
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        eidx = __fadd_rn(eidx, __shfl_xor_sync(0xffffffffu, eidx, o));
    }
    if (FIT != 0) {
        // sweep 4 (the column is in L1): the curve fit of HyposByFit against the depth just regressed
        const float* __restrict__ hcol = a.per_pixel ? a.hypos + (size_t)b * D * HW + p : a.hypos + (size_t)b * D;
        const size_t hstride = a.per_pixel ? HW : 1;
        auto prob_of = [&](int d) { return __fdiv_rn(expf(__fsub_rn(__ldg(col + (size_t)d * HW), m)), sum); };
        float sv;
        if (FIT == 2) {
            LaplaceSums ls;
            if (ok)
                for (int d = d_lo; d < d_hi; ++d) ls.add(__ldg(hcol + (size_t)d * hstride), acc, prob_of(d));
            ls.reduce_slices<PW>();
            sv = ls.scale();
        } else {
            double hs = 0.0;
            if (ok)
                for (int d = d_lo; d < d_hi; ++d) hs += (double)__ldg(hcol + (size_t)d * hstride);
            const double mean = reduce_slices_f64<PW>(hs) / (double)D;
            GaussMoments gm;
            if (ok)
                for (int d = d_lo; d < d_hi; ++d) gm.add((double)__ldg(hcol + (size_t)d * hstride) - mean, prob_of(d));
            gm.reduce_slices<PW>();
            sv = gm.scale(D);
        }
        if (ok && slice == 0) a.s[pix] = sv;
    }
    if (!ok || slice != 0) return;
    if (a.depth) a.depth[pix] = acc;
    if (a.conf) {
        auto prob_at = [&](int d) { return __fdiv_rn(expf(__fsub_rn(__ldg(col + (size_t)d * HW), m)), sum); };
        const float c = window_confidence(prob_at, eidx, D, a.conf_n, a.pad_front, a.pad_back);
        const int y = (int)(p / a.W), x = (int)(p % a.W);
        store_upsampled(a.conf, c, b, y, x, a.H, a.W, a.up);
    }
}

// ------------------------------------------------------------------------------------------------
// Register-resident variant for the configured depths (D = 8, 24, 48; config.py:199).  A warp owns 32/DS
// pixels x DS slices of the depth axis (lane = slice * (32/DS) + pixel); every lane keeps its D/DS logits in
// registers (independent loads in flight, logits read exactly once), max / sum / expectations finish with
// shfl.xor across the slices.  DS = 1 (one thread per pixel, the reference's sequential summation order) is
// used whenever the confidence -- a discrete decision on trunc(sum p*d) -- is produced; DS = 4 gives the deep,
// small stages (D = 48 / 24: 29 K / 115 K pixels) four times the threads.
// ------------------------------------------------------------------------------------------------
template <int D, int DS, int FIT>
__global__ void __launch_bounds__(128)
softmax_regress_reg_kernel(const HeadArgs a)
{
    constexpr int PW = 32 / DS, DQ = D / DS;
    static_assert(D % DS == 0, "slices must divide the depth");
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int slice = lane / PW;
    const size_t HW = (size_t)a.H * a.W;
    const size_t pix = warp * PW + (lane % PW);
    const bool ok = pix < (size_t)a.B * HW;
    const int b = ok ? (int)(pix / HW) : 0;
    const size_t p = ok ? pix % HW : 0;
    const int d0 = slice * DQ;
    const float* __restrict__ col = a.logits + ((size_t)b * D + d0) * HW + p;
    const float* __restrict__ hcol = a.per_pixel ? a.hypos + ((size_t)b * D + d0) * HW + p : a.hypos + (size_t)b * D + d0;
    const size_t hstride = a.per_pixel ? HW : 1;
    // logits and hypotheses of the column: 2*DQ independent loads in flight, one round trip to DRAM
    float e[DQ], hv[DQ];
#pragma unroll
    for (int d = 0; d < DQ; ++d) e[d] = ok ? __ldg(col + (size_t)d * HW) : 0.0f;
#pragma unroll
    for (int d = 0; d < DQ; ++d) hv[d] = ok ? __ldg(hcol + (size_t)d * hstride) : 0.0f;
    float m = e[0];
#pragma unroll
    for (int d = 1; d < DQ; ++d) m = fmaxf(m, e[d]);
#pragma unroll
    for (int o = PW; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.0f;
#pragma unroll
    for (int d = 0; d < DQ; ++d) { e[d] = expf(__fsub_rn(e[d], m)); sum = __fadd_rn(sum, e[d]); }
#pragma unroll
    for (int o = PW; o < 32; o <<= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
    float* __restrict__ pcol = a.prob ? a.prob + ((size_t)b * D + d0) * HW + p : nullptr;
    float acc = 0.0f, eidx = 0.0f;
    if (ok) {
#pragma unroll
        for (int d = 0; d < DQ; ++d) {
            e[d] = __fdiv_rn(e[d], sum);
            if (pcol) pcol[(size_t)d * HW] = e[d];
            acc = __fadd_rn(acc, __fmul_rn(e[d], hv[d]));                                // regress.py:7
            eidx = __fadd_rn(eidx, __fmul_rn(e[d], (float)(d0 + d)));                    // regress.py:15-17
        }
    }
#pragma unroll
    for (int o = PW; o < 32; o <<= 1) {
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        eidx = __fadd_rn(eidx, __shfl_xor_sync(0xffffffffu, eidx, o));
    }
    if (FIT != 0) {
        // the curve fit of HyposByFit on the register-resident column, against the depth just regressed
        float sv;
        if (FIT == 2) {
            LaplaceSums ls;
            if (ok) {
#pragma unroll
                for (int d = 0; d < DQ; ++d) ls.add(hv[d], acc, e[d]);
            }
            ls.reduce_slices<PW>();
            sv = ls.scale();
        } else {
            double hs = 0.0;
#pragma unroll
            for (int d = 0; d < DQ; ++d) hs += (double)hv[d];
            const double mean = reduce_slices_f64<PW>(hs) / (double)D;
            GaussMoments gm;
            if (ok) {
#pragma unroll
                for (int d = 0; d < DQ; ++d) gm.add((double)hv[d] - mean, e[d]);
            }
            gm.reduce_slices<PW>();
            sv = gm.scale(D);
        }
        if (ok && slice == 0) a.s[pix] = sv;
    }
    if (!ok || slice != 0) return;
    if (a.depth) a.depth[pix] = acc;
    if (DS == 1 && a.conf) {
        // window sum around trunc(eidx): the column lives in registers, so select with a compile-time loop
        const int Dp = D + a.pad_front + a.pad_back - a.conf_n + 1;
        const int k = max(0, min((int)eidx, Dp - 1));
        const int lo = k - a.pad_front, hi = lo + a.conf_n;       // window [lo, hi)
        float s = 0.0f;
#pragma unroll
        for (int d = 0; d < DQ; ++d)
            if (d >= lo && d < hi) s = __fadd_rn(s, e[d]);
        const float fn = (float)a.conf_n;
        store_upsampled(a.conf, __fmul_rn(fn, __fdiv_rn(s, fn)), b, (int)(p / a.W), (int)(p % a.W), a.H, a.W, a.up);
    }
}

template <int D, int DS>
static int launch_reg(const HeadArgs& a, int fit, size_t npix, cudaStream_t stream)
{
    const size_t blocks = (npix * DS + 127) / 128;
    if (blocks > 0x7fffffffu) return MDF_ERR_UNSUPPORTED;
    if (fit == 1) softmax_regress_reg_kernel<D, DS, 1><<<(unsigned)blocks, 128, 0, stream>>>(a);
    else if (fit == 2) softmax_regress_reg_kernel<D, DS, 2><<<(unsigned)blocks, 128, 0, stream>>>(a);
    else softmax_regress_reg_kernel<D, DS, 0><<<(unsigned)blocks, 128, 0, stream>>>(a);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------
// depth_regression / confidence_regress on a given probability volume (the reference's split API).
// One thread per pixel, sequential over D: bit-identical to the reference's accumulation order.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
regress_kernel(const HeadArgs a)
{
    const size_t HW = (size_t)a.H * a.W;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (size_t)a.B * HW) return;
    const int b = (int)(pix / HW);
    const size_t p = pix % HW;
    const int D = a.D;
    const float* __restrict__ col = a.logits + (size_t)b * D * HW + p;
    if (a.depth) {
        const float* __restrict__ hcol = a.per_pixel ? a.hypos + (size_t)b * D * HW + p : a.hypos + (size_t)b * D;
        const size_t hstride = a.per_pixel ? HW : 1;
        float acc = 0.0f;
        for (int d = 0; d < D; ++d) acc = __fadd_rn(acc, __fmul_rn(__ldg(col + (size_t)d * HW), __ldg(hcol + (size_t)d * hstride)));
        a.depth[pix] = acc;
    }
    if (a.conf) {
        float eidx = 0.0f;
        for (int d = 0; d < D; ++d) eidx = __fadd_rn(eidx, __fmul_rn(__ldg(col + (size_t)d * HW), (float)d));
        auto prob_at = [&](int d) { return __ldg(col + (size_t)d * HW); };
        const float c = window_confidence(prob_at, eidx, D, a.conf_n, a.pad_front, a.pad_back);
        store_upsampled(a.conf, c, b, (int)(p / a.W), (int)(p % a.W), a.H, a.W, a.up);
    }
}

static int check_head(const HeadArgs& a, bool need_hypos)
{
    if (a.B < 0 || a.D < 0 || a.H < 0 || a.W < 0) return MDF_ERR_INVALID_SHAPE;
    if (a.conf) {
        if (a.conf_n <= 0 || a.up <= 0 || a.pad_front < 0 || a.pad_back < 0) return MDF_ERR_INVALID_SHAPE;
        if (a.D + a.pad_front + a.pad_back - a.conf_n + 1 <= 0) return MDF_ERR_INVALID_SHAPE;
    }
    if ((size_t)a.B * a.H * a.W == 0) return MDF_OK;
    if (a.D == 0) return MDF_ERR_INVALID_SHAPE;         // softmax / gather over an empty axis
    if (!a.logits || (need_hypos && !a.hypos)) return MDF_ERR_NULL_POINTER;
    return MDF_OK;
}

static int head_device(const HeadArgs& a)
{
    const void* out = a.depth ? (const void*)a.depth : a.conf ? (const void*)a.conf : a.prob ? (const void*)a.prob : (const void*)a.s;
    const int dev = device_of(out);
    if (dev < 0) return dev;
    const void* ptrs[6];
    int n = 0;
    ptrs[n++] = a.logits;
    if (a.s) ptrs[n++] = a.s;
    if (a.hypos) ptrs[n++] = a.hypos;
    if (a.prob) ptrs[n++] = a.prob;
    if (a.depth) ptrs[n++] = a.depth;
    if (a.conf) ptrs[n++] = a.conf;
    const int st = check_on_device(dev, ptrs, n);
    return st != MDF_OK ? st : dev;
}

}  // namespace mdf

using namespace mdf;

extern "C" {

int mdf_softmax_regress_fit_fwd(const float* logits, const float* depth_hypos, int hypos_per_pixel, int B, int D, int H,
                                int W, float* prob, float* depth, float* confidence, int conf_n, int conf_pad_front,
                                int conf_pad_back, int conf_upsample, int curve, float* s, mdf_stream_t stream_)
{
    HeadArgs a;
    a.logits = logits; a.hypos = depth_hypos; a.prob = prob; a.depth = depth; a.conf = confidence; a.s = s;
    a.per_pixel = hypos_per_pixel; a.B = B; a.D = D; a.H = H; a.W = W;
    a.conf_n = conf_n; a.pad_front = conf_pad_front; a.pad_back = conf_pad_back; a.up = conf_upsample;
    if (curve < 0 || curve > 2) return MDF_ERR_UNSUPPORTED;
    if (B >= 0 && D >= 0 && H >= 0 && W >= 0 && (size_t)B * H * W == 0) return MDF_OK;     // empty: torch hands out NULL pointers
    if (!prob && !depth && !confidence && !s) return MDF_ERR_NULL_POINTER;
    if ((curve != 0) != (s != nullptr)) return MDF_ERR_NULL_POINTER;      // s is produced iff a curve is named
    int st = check_head(a, depth != nullptr || curve != 0);
    if (st != MDF_OK) return st;
    const size_t npix = (size_t)B * H * W;
    if (npix == 0) return MDF_OK;
    const int dev = head_device(a);
    if (dev < 0) return dev;
    DeviceGuard guard(dev);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (D == 8 || D == 24 || D == 48) {
        if (confidence) {                                    // sequential order for the discrete index
            if (D == 8) return launch_reg<8, 1>(a, curve, npix, stream);
            if (D == 24) return launch_reg<24, 1>(a, curve, npix, stream);
            return launch_reg<48, 1>(a, curve, npix, stream);
        }
        if (D == 8) return launch_reg<8, 1>(a, curve, npix, stream);
        if (D == 24) return launch_reg<24, 4>(a, curve, npix, stream);
        return launch_reg<48, 4>(a, curve, npix, stream);
    }
    // other depths: depth slices per warp; sequential order whenever the discrete confidence index is produced
    const int ds = (confidence || D < 16) ? 1 : 4;
    const size_t threads = npix * ds;
    const size_t blocks = (threads + 255) / 256;
    if (blocks > 0x7fffffffu) return MDF_ERR_UNSUPPORTED;
    const unsigned nb = (unsigned)blocks;
    if (ds == 1) {
        if (curve == 1) softmax_regress_kernel<1, 1><<<nb, 256, 0, stream>>>(a);
        else if (curve == 2) softmax_regress_kernel<1, 2><<<nb, 256, 0, stream>>>(a);
        else softmax_regress_kernel<1, 0><<<nb, 256, 0, stream>>>(a);
    } else {
        if (curve == 1) softmax_regress_kernel<4, 1><<<nb, 256, 0, stream>>>(a);
        else if (curve == 2) softmax_regress_kernel<4, 2><<<nb, 256, 0, stream>>>(a);
        else softmax_regress_kernel<4, 0><<<nb, 256, 0, stream>>>(a);
    }
    return launch_status();
}

int mdf_softmax_regress_fwd(const float* logits, const float* depth_hypos, int hypos_per_pixel, int B, int D, int H,
                            int W, float* prob, float* depth, float* confidence, int conf_n, int conf_pad_front,
                            int conf_pad_back, int conf_upsample, mdf_stream_t stream)
{
    return mdf_softmax_regress_fit_fwd(logits, depth_hypos, hypos_per_pixel, B, D, H, W, prob, depth, confidence, conf_n,
                                       conf_pad_front, conf_pad_back, conf_upsample, 0, nullptr, stream);
}

static int run_regress(HeadArgs& a, mdf_stream_t stream_)
{
    int st = check_head(a, a.depth != nullptr);
    if (st != MDF_OK) return st;
    const size_t npix = (size_t)a.B * a.H * a.W;
    if (npix == 0) return MDF_OK;
    const int dev = head_device(a);
    if (dev < 0) return dev;
    DeviceGuard guard(dev);
    const size_t blocks = (npix + 255) / 256;
    if (blocks > 0x7fffffffu) return MDF_ERR_UNSUPPORTED;
    regress_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(a);
    return launch_status();
}

int mdf_depth_regression_fwd(const float* prob, const float* depth_hypos, int hypos_per_pixel, int B, int D, int H, int W,
                             float* depth, mdf_stream_t stream)
{
    if (!depth) return MDF_ERR_NULL_POINTER;
    HeadArgs a;
    a.logits = prob; a.hypos = depth_hypos; a.prob = nullptr; a.depth = depth; a.conf = nullptr; a.s = nullptr;
    a.per_pixel = hypos_per_pixel; a.B = B; a.D = D; a.H = H; a.W = W;
    a.conf_n = 0; a.pad_front = 0; a.pad_back = 0; a.up = 1;
    return run_regress(a, stream);
}

int mdf_confidence_fwd(const float* prob, int B, int D, int H, int W, int n, int pad_front, int pad_back, int upsample,
                       float* confidence, mdf_stream_t stream)
{
    if (!confidence) return MDF_ERR_NULL_POINTER;
    HeadArgs a;
    a.logits = prob; a.hypos = nullptr; a.prob = nullptr; a.depth = nullptr; a.conf = confidence; a.s = nullptr;
    a.per_pixel = 0; a.B = B; a.D = D; a.H = H; a.W = W;
    a.conf_n = n; a.pad_front = pad_front; a.pad_back = pad_back; a.up = upsample;
    return run_regress(a, stream);
}

}  // extern "C"

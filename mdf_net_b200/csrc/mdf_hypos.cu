// mdf_hypos.cu -- next-stage depth hypotheses: HyposByFit of MDF-Net (net/unit/depthhypos.py) for sm_100a.
//
// The reference (depthhypos.py:40-76) fits, per pixel, a curve to the probability column the regulariser
// produced -- "gauss1" after stage 0 (:169-215), "laplace" after stage 1 (:78-125); "gauss0" (:127-167) is available but
// not wired by config.py --, upsamples the fitted
// scale s and the regressed depth x2 (bilinear, align_corners=False), turns s into a search range with
// prob_thresh, clamps it, and spreads `ndepths` hypotheses over it.  It does so with ~40 ATen launches, a
// per-pixel batched 3x3 torch.inverse, and Python loops over the depth planes and the batch.
//
//   hypos_fit_kernel       one thread per pixel, one sweep over the D probabilities (coalesced planes).
//                          laplace: two float sums, as the reference.  gauss1: the normal equations of
//                          ln p ~ c2 x^2 + c1 x + c0 have entries up to 935^4 * 48 -- in float32 the reference's own
//                          result is 2e-4 (median) to 3e-2 (max) away from a float64 evaluation of the same
//                          formula -- so the moments are accumulated in float64 around the column mean (c2 does
//                          not depend on the shift) and c2 comes from Cramer's rule: 1e-7 of the exact value.
//   hypos_generate_kernel  one thread per output pixel: bilinear x2 taps of s and depth, range, clamps, then the
//                          D' hypotheses as coalesced rows of the (B,D',2H,2W) tensor, written once.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdf_common.cuh"
#include "mdf_host.cuh"

namespace mdf {

struct FitArgs {
    const float* prob;     // (B,D,H,W)
    const float* hypos;    // (B,D) or (B,D,H,W)
    const float* depth;    // (B,H,W)
    float* s;              // (B,H,W)
    int per_pixel, B, D, H, W;
};

template <int MODE>   // 1 = gauss1, 2 = laplace, 3 = gauss0
__global__ void __launch_bounds__(128)
hypos_fit_kernel(const FitArgs a)
{
    const size_t HW = (size_t)a.H * a.W;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (size_t)a.B * HW) return;
    const int b = (int)(pix / HW);
    const size_t p = pix % HW;
    const float* __restrict__ pcol = a.prob + (size_t)b * a.D * HW + p;
    const float* __restrict__ hcol = a.per_pixel ? a.hypos + (size_t)b * a.D * HW + p : a.hypos + (size_t)b * a.D;
    const size_t hs = a.per_pixel ? HW : 1;
    if (MODE == 2) {
        // x = |hypo - depth|, y = ln max(p, 1e-40); s = 1 / |sum(x y) / sum(x x)|      (depthhypos.py:116-123)
        const float dep = __ldg(a.depth + pix);
        float sxy = 0.0f, sxx = 0.0f;
        for (int d = 0; d < a.D; ++d) {
            const float x = fabsf(__fsub_rn(__ldg(hcol + (size_t)d * hs), dep));
            const float y = logf(fmaxf(__ldg(pcol + (size_t)d * HW), 1e-40f));
            sxy = __fadd_rn(sxy, __fmul_rn(x, y));
            sxx = __fadd_rn(sxx, __fmul_rn(x, x));
        }
        a.s[pix] = __fdiv_rn(1.0f, fabsf(__fdiv_rn(sxy, sxx)));
        return;
    }
    if (MODE == 3) {
        // gauss0: least squares of ln p on [x, 1] with x = (hypo - depth)^2 in float32 as the reference forms it (:153);
        // s = |-1 / b0|, b0 = S_xz / S_xx.  The 2x2 normal equations have entries up to 510^4 * 48: like gauss1, the sums are
        // taken in float64 around the column mean of x (the reference's own float32 result is its noise, not the target).
        const float dep = __ldg(a.depth + pix);
        double xm = 0.0;
        for (int d = 0; d < a.D; ++d) {
            const float df = __fsub_rn(__ldg(hcol + (size_t)d * hs), dep);
            xm += (double)__fmul_rn(df, df);
        }
        xm /= (double)a.D;
        double sxx = 0.0, sxz = 0.0;
        for (int d = 0; d < a.D; ++d) {
            const float df = __fsub_rn(__ldg(hcol + (size_t)d * hs), dep);
            const double u = (double)__fmul_rn(df, df) - xm;
            const double z = log((double)fmaxf(__ldg(pcol + (size_t)d * HW), 1e-40f));
            sxx += u * u; sxz += u * z;
        }
        a.s[pix] = (float)fabs(-sxx / sxz);
        return;
    }
    // gauss1: least squares of ln p on [x^2, x, 1]; s = |-1 / c2|                      (depthhypos.py:189-213)
    double mean = 0.0;
    for (int d = 0; d < a.D; ++d) mean += (double)__ldg(hcol + (size_t)d * hs);
    mean /= (double)a.D;
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, r0 = 0, r1 = 0, r2 = 0;
    for (int d = 0; d < a.D; ++d) {
        const double u = (double)__ldg(hcol + (size_t)d * hs) - mean;
        const double z = log((double)fmaxf(__ldg(pcol + (size_t)d * HW), 1e-40f));
        const double u2 = u * u;
        m0 += 1.0; m1 += u; m2 += u2; m3 += u2 * u; m4 += u2 * u2;
        r0 += z; r1 += u * z; r2 += u2 * z;
    }
    // [[m4 m3 m2][m3 m2 m1][m2 m1 m0]] [c2 c1 c0]^T = [r2 r1 r0]^T
    const double k1 = m2 * m0 - m1 * m1, k2 = m3 * m0 - m1 * m2, k3 = m3 * m1 - m2 * m2;
    const double det = m4 * k1 - m3 * k2 + m2 * k3;
    const double num = r2 * k1 - m3 * (r1 * m0 - m1 * r0) + m2 * (r1 * m1 - m2 * r0);
    a.s[pix] = (float)fabs(-det / num);             // |-1 / c2|, c2 = num / det
}

struct GenArgs {
    const float* depth;    // (B,H,W)
    const float* s;        // (B,H,W)
    const float* range;    // (B,2)
    float* out;            // (B,ND,Ho,Wo)
    float log_thresh;
    int mode, upsample, B, H, W, ND;
};

// F.interpolate(scale_factor=2, mode='bilinear', align_corners=False): ATen's area_pixel_compute_source_index
// (src = 0.5*(dst+0.5) - 0.5, clamped at 0) and the h0*(w0*a + w1*b) + h1*(w0*c + w1*d) blend
__device__ __forceinline__ float up2(const float* __restrict__ m, int H, int W, int Y, int X)
{
    const float sy = fmaxf(__fmaf_rn((float)Y + 0.5f, 0.5f, -0.5f), 0.0f), sx = fmaxf(__fmaf_rn((float)X + 0.5f, 0.5f, -0.5f), 0.0f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = __fsub_rn(sy, (float)y0), lx = __fsub_rn(sx, (float)x0);
    const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
    const float top = __fadd_rn(__fmul_rn(hx, __ldg(m + (size_t)y0 * W + x0)), __fmul_rn(lx, __ldg(m + (size_t)y0 * W + x1)));
    const float bot = __fadd_rn(__fmul_rn(hx, __ldg(m + (size_t)y1 * W + x0)), __fmul_rn(lx, __ldg(m + (size_t)y1 * W + x1)));
    return __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
}

__global__ void __launch_bounds__(256)
hypos_generate_kernel(const GenArgs a)
{
    const int Ho = a.upsample ? 2 * a.H : a.H, Wo = a.upsample ? 2 * a.W : a.W;
    const size_t HWo = (size_t)Ho * Wo;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (size_t)a.B * HWo) return;
    const int b = (int)(pix / HWo);
    const int Y = (int)((pix % HWo) / Wo), X = (int)(pix % Wo);
    const size_t HW = (size_t)a.H * a.W;
    // depth_res.clamp(max = (depth_max.max() - depth_min.min()) / 2)   (depthhypos.py:58)
    float gmin = __ldg(a.range), gmax = __ldg(a.range + 1);
    for (int i = 1; i < a.B; ++i) { gmin = fminf(gmin, __ldg(a.range + 2 * i)); gmax = fmaxf(gmax, __ldg(a.range + 2 * i + 1)); }
    const float dmin = __ldg(a.range + 2 * b), dmax = __ldg(a.range + 2 * b + 1);
    const float sv = a.upsample ? up2(a.s + (size_t)b * HW, a.H, a.W, Y, X) : __ldg(a.s + (size_t)b * HW + (size_t)Y * a.W + X);
    const float dv = a.upsample ? up2(a.depth + (size_t)b * HW, a.H, a.W, Y, X) : __ldg(a.depth + (size_t)b * HW + (size_t)Y * a.W + X);
    float res = a.mode == 1 ? __fsqrt_rn(__fmul_rn(__fmul_rn(-1.0f, sv), a.log_thresh))      // sqrt(-1*s*log(thresh))  :54
                            : fabsf(__fmul_rn(sv, a.log_thresh));                              // |s*log(thresh)|        :56
    res = fminf(fmaxf(res, 1e-6f), __fdiv_rn(__fsub_rn(gmax, gmin), 2.0f));                    // :57
    res = fminf(res, __fmul_rn(__fsub_rn(dmax, dmin), 0.2f));                                  // :59-60
    const float interval = __fdiv_rn(res, (float)(a.ND - 1));                                  // :63
    const float base = __fsub_rn(dv, __fmul_rn(0.5f, res));                                    // :64
    float* __restrict__ op = a.out + (size_t)b * a.ND * HWo + (size_t)Y * Wo + X;
    for (int d = 0; d < a.ND; ++d) {
        float h = __fadd_rn(base, __fmul_rn(interval, (float)d));                              // :65-66
        h = __fadd_rn(dmin, fmaxf(__fsub_rn(h, dmin), 0.0f));                                  // :70-71
        h = __fadd_rn(dmax, fminf(__fsub_rn(h, dmax), 0.0f));                                  // :72-73
        op[(size_t)d * HWo] = h;
    }
}

}  // namespace mdf

using namespace mdf;

extern "C" {

int mdf_hypos_fit_fwd(const float* prob, const float* depth_hypos, int hypos_per_pixel, const float* depth, int curve,
                      int B, int D, int H, int W, float* s, mdf_stream_t stream)
{
    if (B < 0 || D < 0 || H < 0 || W < 0) return MDF_ERR_INVALID_SHAPE;
    if (curve != 1 && curve != 2 && curve != 3) return MDF_ERR_UNSUPPORTED;
    const size_t npix = (size_t)B * H * W;
    if (npix == 0) return MDF_OK;
    if (D < 1) return MDF_ERR_INVALID_SHAPE;
    if (!prob || !depth_hypos || !depth || !s) return MDF_ERR_NULL_POINTER;
    const int dev = device_of(s);
    if (dev < 0) return dev;
    const void* ptrs[] = {prob, depth_hypos, depth};
    int st = check_on_device(dev, ptrs, 3);
    if (st != MDF_OK) return st;
    DeviceGuard guard(dev);
    FitArgs a;
    a.prob = prob; a.hypos = depth_hypos; a.depth = depth; a.s = s;
    a.per_pixel = hypos_per_pixel; a.B = B; a.D = D; a.H = H; a.W = W;
    const size_t blocks = (npix + 127) / 128;
    if (blocks > 0x7fffffffu) return MDF_ERR_UNSUPPORTED;
    if (curve == 1) hypos_fit_kernel<1><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(a);
    else if (curve == 3) hypos_fit_kernel<3><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(a);
    else hypos_fit_kernel<2><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

int mdf_hypos_generate_fwd(const float* depth, const float* s, const float* depth_range, int curve, float prob_thresh,
                           int upsample, int B, int H, int W, int ndepths, float* depth_hypos, mdf_stream_t stream)
{
    if (B < 0 || H < 0 || W < 0 || ndepths < 2) return MDF_ERR_INVALID_SHAPE;
    if (curve != 1 && curve != 2 && curve != 3) return MDF_ERR_UNSUPPORTED;
    const size_t npix = (size_t)B * H * W * (upsample ? 4 : 1);
    if (npix == 0) return MDF_OK;
    if (!depth || !s || !depth_range || !depth_hypos) return MDF_ERR_NULL_POINTER;
    const int dev = device_of(depth_hypos);
    if (dev < 0) return dev;
    const void* ptrs[] = {depth, s, depth_range};
    int st = check_on_device(dev, ptrs, 3);
    if (st != MDF_OK) return st;
    DeviceGuard guard(dev);
    GenArgs a;
    a.depth = depth; a.s = s; a.range = depth_range; a.out = depth_hypos;
    a.log_thresh = logf(prob_thresh);
    a.mode = curve == 2 ? 2 : 1;                    // gauss0 and gauss1 share the range formula (depthhypos.py:53-54)
    a.upsample = upsample ? 1 : 0; a.B = B; a.H = H; a.W = W; a.ND = ndepths;
    const size_t blocks = (npix + 255) / 256;
    if (blocks > 0x7fffffffu) return MDF_ERR_UNSUPPORTED;
    hypos_generate_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    return launch_status();
}

}  // extern "C"

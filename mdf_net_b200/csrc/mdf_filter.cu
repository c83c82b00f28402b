// mdf_filter.cu -- geometric-consistency filter of MDF-Net's post-processing for sm_100a.
//
// Reference: tools/filter/dynamic_filter_gpu.py -- reproject_with_depth (:184-237), check_geometric_consistency
// (:161-182), the per-view aggregation of filter() (:57-100); bilinear_sampler = grid_sample(bilinear, zeros,
// align_corners=True) on pixel coordinates (tools/filter/data_io.py:117-131).  Per (reference view, source view) the
// reference launches ~60 ATen kernels and three cuSOLVER / cuBLAS calls over full-resolution maps and materialises a
// dozen (3|4, H*W) intermediates; here ONE launch per reference view walks all source views per pixel:
//   project the pixel with its depth into the source view -> bilinear tap of the source depth map (the same
//   project / gather pattern as the cost volume) -> project back -> reprojection error and relative depth error ->
//   the 9 dynamic thresholds (i/thre1 px, i/thre2, i = 2..10) -> counts -> geometric mask, averaged depth, photometric
//   and final masks.  Each map is read once (the gathers hit L2), each output written once.
//
//   geo_setup_kernel   (1 block) the matrices of every source view in float64, rounded once: inverse(K_ref),
//                      E_src @ inverse(E_ref), K_src, inverse(K_src), E_ref @ inverse(E_src), K_ref      (:195,200,216,221)
//   geo_filter_kernel  one thread per reference pixel, loop over the source views.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdf_common.cuh"
#include "mdf_host.cuh"

namespace mdf {

constexpr int kGeoMat = 64;      // floats per source view in the workspace: M_rs[16] M_sr[16] Ksinv[9] Ks[9] pad

struct GeoSrcPtrs { const float* p[MDF_MAX_FILTER_VIEWS]; };

__device__ static bool invert_n(const double* a_in, int n, double* inv)      // Gauss-Jordan with partial pivoting
{
    double a[16], b[16];
    for (int i = 0; i < n * n; ++i) { a[i] = a_in[i]; b[i] = (i / n == i % n) ? 1.0 : 0.0; }
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int r = k + 1; r < n; ++r)
            if (fabs(a[r * n + k]) > fabs(a[p * n + k])) p = r;
        if (a[p * n + k] == 0.0) return false;
        if (p != k)
            for (int c = 0; c < n; ++c) {
                double t = a[k * n + c]; a[k * n + c] = a[p * n + c]; a[p * n + c] = t;
                t = b[k * n + c]; b[k * n + c] = b[p * n + c]; b[p * n + c] = t;
            }
        const double piv = 1.0 / a[k * n + k];
        for (int c = 0; c < n; ++c) { a[k * n + c] *= piv; b[k * n + c] *= piv; }
        for (int r = 0; r < n; ++r) {
            if (r == k) continue;
            const double f = a[r * n + k];
            for (int c = 0; c < n; ++c) { a[r * n + c] -= f * a[k * n + c]; b[r * n + c] -= f * b[k * n + c]; }
        }
    }
    for (int i = 0; i < n * n; ++i) inv[i] = b[i];
    return true;
}

static __global__ void geo_setup_kernel(const float* __restrict__ ref_K, const float* __restrict__ ref_E,
                                        const float* __restrict__ src_K, const float* __restrict__ src_E, int S,
                                        float* __restrict__ ref_mats /* Kr_inv[9] Kr[9] */, float* __restrict__ mats)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > S) return;
    double Er[16], Erinv[16];
    for (int i = 0; i < 16; ++i) Er[i] = (double)ref_E[i];
    const bool ok_r = invert_n(Er, 4, Erinv);
    if (s == S) {                                   // the extra thread: the reference camera's own matrices
        double Kr[9], Krinv[9];
        for (int i = 0; i < 9; ++i) Kr[i] = (double)ref_K[i];
        const bool ok = invert_n(Kr, 3, Krinv);
        for (int i = 0; i < 9; ++i) { ref_mats[i] = ok ? (float)Krinv[i] : nanf(""); ref_mats[9 + i] = ref_K[i]; }
        return;
    }
    double Es[16], Esinv[16], Ks[9], Ksinv[9];
    for (int i = 0; i < 16; ++i) Es[i] = (double)src_E[16 * s + i];
    for (int i = 0; i < 9; ++i) Ks[i] = (double)src_K[9 * s + i];
    const bool ok = ok_r && invert_n(Es, 4, Esinv) && invert_n(Ks, 3, Ksinv);
    float* m = mats + (size_t)kGeoMat * s;
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double rs = 0.0, sr = 0.0;
            for (int k = 0; k < 4; ++k) { rs += Es[r * 4 + k] * Erinv[k * 4 + c]; sr += Er[r * 4 + k] * Esinv[k * 4 + c]; }
            m[r * 4 + c] = ok ? (float)rs : nanf("");
            m[16 + r * 4 + c] = ok ? (float)sr : nanf("");
        }
    for (int i = 0; i < 9; ++i) { m[32 + i] = ok ? (float)Ksinv[i] : nanf(""); m[41 + i] = src_K[9 * s + i]; }
}

struct GeoArgs {
    const float* ref_depth;     // (H,W)
    GeoSrcPtrs src;             // S x (H,W)
    const float* confidence;    // (H,W) or nullptr
    const float* ref_mats;      // Kr_inv[9] Kr[9]
    const float* mats;          // S x kGeoMat
    uint16_t* bits;             // (S,H,W) or nullptr: bit i-2 <-> threshold i
    float* depth_reprojected;   // (S,H,W) or nullptr
    float* depth_averaged;      // (H,W) or nullptr
    uint8_t* geo;               // (H,W) or nullptr
    uint8_t* photo;
    uint8_t* fin;
    float photo_threshold;
    float thr1[9], thr2[9];     // i / thre1, i / thre2 for i = 2..10 (the host's float divisions are the same IEEE quotients)
    int nconditions, S, H, W;
};

__device__ __forceinline__ void mat3v(const float* __restrict__ m, float x, float y, float z, float (&o)[3])
{
#pragma unroll
    for (int r = 0; r < 3; ++r) o[r] = __fmaf_rn(m[r * 3 + 2], z, __fmaf_rn(m[r * 3 + 1], y, __fmul_rn(m[r * 3], x)));
}
__device__ __forceinline__ void mat34v(const float* __restrict__ m, const float (&v)[3], float (&o)[3])
{
#pragma unroll
    for (int r = 0; r < 3; ++r)
        o[r] = __fadd_rn(__fmaf_rn(m[r * 4 + 2], v[2], __fmaf_rn(m[r * 4 + 1], v[1], __fmul_rn(m[r * 4], v[0]))), m[r * 4 + 3]);
}

// grid_sample(bilinear, zeros, align_corners=True) through bilinear_sampler's normalisation (data_io.py:121-125)
__device__ __forceinline__ float sample_depth_ac(const float* __restrict__ img, int H, int W, float px, float py,
                                                 float r_wm1, float r_hm1)
{
    const float gx = __fsub_rn(div_shared(__fmul_rn(2.0f, px), (float)(W - 1), r_wm1), 1.0f);
    const float gy = __fsub_rn(div_shared(__fmul_rn(2.0f, py), (float)(H - 1), r_hm1), 1.0f);
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
    const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
    if (!(ix > -1.0f && ix < (float)W && iy > -1.0f && iy < (float)H)) return 0.0f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float ax = __fsub_rn(__fadd_rn(fx, 1.0f), ix), bx = __fsub_rn(ix, fx);
    const float ay = __fsub_rn(__fadd_rn(fy, 1.0f), iy), by = __fsub_rn(iy, fy);
    const bool x0in = (unsigned)x0 < (unsigned)W, x1in = (unsigned)(x0 + 1) < (unsigned)W;
    const bool y0in = (unsigned)y0 < (unsigned)H, y1in = (unsigned)(y0 + 1) < (unsigned)H;
    const float* p = img + (ptrdiff_t)y0 * W + x0;
    const float nw = (x0in && y0in) ? __ldg(p) : 0.0f, ne = (x1in && y0in) ? __ldg(p + 1) : 0.0f;
    const float sw = (x0in && y1in) ? __ldg(p + W) : 0.0f, se = (x1in && y1in) ? __ldg(p + W + 1) : 0.0f;
    return __fmaf_rn(se, __fmul_rn(bx, by), __fmaf_rn(sw, __fmul_rn(ax, by), __fmaf_rn(ne, __fmul_rn(bx, ay), __fmul_rn(nw, __fmul_rn(ax, ay)))));
}

static __global__ void __launch_bounds__(256)
geo_filter_kernel(const GeoArgs a)
{
    extern __shared__ float mat_s[];                 // ref_mats[18] (+pad to 20), then S x kGeoMat
    for (int i = threadIdx.x; i < 20 + a.S * kGeoMat; i += blockDim.x)
        mat_s[i] = i < 18 ? __ldg(a.ref_mats + i) : (i < 20 ? 0.0f : __ldg(a.mats + (i - 20)));
    __syncthreads();
    const uint32_t HW = (uint32_t)a.H * (uint32_t)a.W;           // < 2^31 (host checked): 32-bit index arithmetic
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int y = (int)(p / (uint32_t)a.W), x = (int)(p % (uint32_t)a.W);
    const float fx = (float)x, fy = (float)y;
    const float d = __ldg(a.ref_depth + p);
    const float* Krinv = mat_s;
    const float* Kr = mat_s + 9;
    const float r_wm1 = refine_rcp((float)(a.W - 1)), r_hm1 = refine_rcp((float)(a.H - 1)), r_d = refine_rcp(d);
    // The thresholds grow with i, so a source view passes all of them from some i on: one count per view (how many it
    // fails, 0..9) and a histogram of those counts, 6 bits per bin (S <= 32), instead of 9 mask bits and 9 counters.
    unsigned long long hist = 0;
    int nvalid = 0;
    float dsum = 0.0f;
    for (int s = 0; s < a.S; ++s) {
        const float* m = mat_s + 20 + s * kGeoMat;
        float v[3], q[3], k[3];
        mat3v(Krinv, __fmul_rn(fx, d), __fmul_rn(fy, d), d, v);                       // :196-198
        mat34v(m, v, q);                                                              // :200-201
        mat3v(m + 41, q[0], q[1], q[2], k);                                           // :203
        const float rk = refine_rcp(k[2]);
        const float xs = div_shared(k[0], k[2], rk), ys = div_shared(k[1], k[2], rk);  // :204
        const float ds = sample_depth_ac(a.src.p[s], a.H, a.W, xs, ys, r_wm1, r_hm1);  // :212
        mat3v(m + 32, __fmul_rn(xs, ds), __fmul_rn(ys, ds), ds, v);                   // :216-217
        mat34v(m + 16, v, q);                                                         // :219-220
        const float drep = q[2];                                                      // :222
        mat3v(Kr, q[0], q[1], q[2], k);                                               // :223
        const float rk2 = refine_rcp(k[2]);
        const float xr = div_shared(k[0], k[2], rk2), yr = div_shared(k[1], k[2], rk2); // :224
        const float dx = __fsub_rn(xr, fx), dy = __fsub_rn(yr, fy);
        const float dist = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));   // :170
        const float rel = div_shared(fabsf(__fsub_rn(drep, d)), d, r_d);              // :173-174
        int nfail = 0;                                                                // :175-177, i = 2..10
#pragma unroll
        for (int i = 0; i < 9; ++i)
            if (!(dist < a.thr1[i] && rel < a.thr2[i])) ++nfail;
        const uint32_t b = (0x1FFu << nfail) & 0x1FFu;        // bit i-2 <-> the mask of threshold i
        hist += 1ull << (6 * nfail);
        const bool last = nfail < 9;                          // passes i = 10
        if (a.bits) a.bits[(size_t)s * HW + p] = (uint16_t)b;
        if (a.depth_reprojected) a.depth_reprojected[(size_t)s * HW + p] = last ? drep : 0.0f;      // :180
        if (last) { ++nvalid; dsum = __fadd_rn(dsum, drep); }
    }
    int geo = 0, cum = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {                             // views that pass threshold i+2 = views failing at most i of them
        cum += (int)((hist >> (6 * i)) & 63ull);
        geo += cum >= i + 2;                                  // filter():86-88
    }
    const bool g = a.S > 0 && geo >= a.nconditions;
    const bool ph = a.confidence ? __ldg(a.confidence + p) > a.photo_threshold : true;
    if (a.depth_averaged) a.depth_averaged[p] = __fdiv_rn(__fadd_rn(dsum, d), (float)(nvalid + 1));  // :93
    if (a.geo) a.geo[p] = g;
    if (a.photo) a.photo[p] = ph;
    if (a.fin) a.fin[p] = g && ph;
}

}  // namespace mdf

using namespace mdf;

extern "C" {

size_t mdf_geo_filter_workspace_bytes(int S) { return (size_t)(20 + (S > 0 ? S : 0) * kGeoMat) * sizeof(float) + 256; }

int mdf_geo_filter_fwd(const float* ref_depth, const float* ref_intrinsics, const float* ref_extrinsics,
                       const float* const* src_depths, const float* src_intrinsics, const float* src_extrinsics,
                       int S, int H, int W, const float* confidence, float photo_threshold, int nconditions, float thre1,
                       float thre2, uint16_t* src_bits, float* depth_reprojected, float* depth_averaged, uint8_t* geo_mask,
                       uint8_t* photo_mask, uint8_t* final_mask, void* workspace, size_t workspace_bytes, mdf_stream_t stream_)
{
    if (S < 0 || H < 0 || W < 0) return MDF_ERR_INVALID_SHAPE;
    if (S > MDF_MAX_FILTER_VIEWS) return MDF_ERR_UNSUPPORTED;
    // the kernel counts how many of the nine thresholds i / thre (i = 2..10) a view fails: they must grow with i
    if (!(thre1 > 0.0f && thre2 > 0.0f && thre1 < 3.0e38f && thre2 < 3.0e38f)) return MDF_ERR_UNSUPPORTED;
    const size_t HW = (size_t)H * W;
    if (HW == 0) return MDF_OK;
    if (!ref_depth || !ref_intrinsics || !ref_extrinsics || !workspace) return MDF_ERR_NULL_POINTER;
    if (S > 0 && (!src_depths || !src_intrinsics || !src_extrinsics)) return MDF_ERR_NULL_POINTER;
    if (!src_bits && !depth_reprojected && !depth_averaged && !geo_mask && !photo_mask && !final_mask) return MDF_ERR_NULL_POINTER;
    if (workspace_bytes < mdf_geo_filter_workspace_bytes(S)) return MDF_ERR_WORKSPACE;
    const int dev = device_of(ref_depth);
    if (dev < 0) return dev;
    GeoArgs a;
    {
        const void* ptrs[16];
        int n = 0;
        ptrs[n++] = ref_intrinsics; ptrs[n++] = ref_extrinsics; ptrs[n++] = workspace;
        if (S > 0) { ptrs[n++] = src_intrinsics; ptrs[n++] = src_extrinsics; }
        if (confidence) ptrs[n++] = confidence;
        if (src_bits) ptrs[n++] = src_bits;
        if (depth_reprojected) ptrs[n++] = depth_reprojected;
        if (depth_averaged) ptrs[n++] = depth_averaged;
        if (geo_mask) ptrs[n++] = geo_mask;
        if (photo_mask) ptrs[n++] = photo_mask;
        if (final_mask) ptrs[n++] = final_mask;
        int st = check_on_device(dev, ptrs, n);
        if (st != MDF_OK) return st;
        for (int s = 0; s < S; ++s) {
            if (!src_depths[s]) return MDF_ERR_NULL_POINTER;
            const void* q = src_depths[s];
            st = check_on_device(dev, &q, 1);
            if (st != MDF_OK) return st;
            a.src.p[s] = src_depths[s];
        }
    }
    DeviceGuard guard(dev);
    cudaStream_t stream = (cudaStream_t)stream_;
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    geo_setup_kernel<<<(S + 1 + 31) / 32, 32, 0, stream>>>(ref_intrinsics, ref_extrinsics, src_intrinsics, src_extrinsics, S, ws, ws + 20);
    int st = launch_status();
    if (st != MDF_OK) return st;
    a.ref_depth = ref_depth; a.confidence = confidence; a.ref_mats = ws; a.mats = ws + 20;
    a.bits = src_bits; a.depth_reprojected = depth_reprojected; a.depth_averaged = depth_averaged;
    a.geo = geo_mask; a.photo = photo_mask; a.fin = final_mask;
    a.photo_threshold = photo_threshold; a.nconditions = nconditions;
    for (int i = 0; i < 9; ++i) { a.thr1[i] = (float)(i + 2) / thre1; a.thr2[i] = (float)(i + 2) / thre2; }
    a.S = S; a.H = H; a.W = W;
    const size_t blocks = (HW + 255) / 256;
    if (HW > 0x7fffffffu) return MDF_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(20 + S * kGeoMat) * sizeof(float);
    geo_filter_kernel<<<(unsigned)blocks, 256, smem, stream>>>(a);
    return launch_status();
}

}  // extern "C"

/*
 * mdf_oracle.c -- CPU restatement of MDF-Net's plane-sweep cost-volume path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (mdf_net_b200/) may
 * link, import or execute this file; it exists so that tests/, smoke() and
 * bench.py's cpu_baseline leg have an independent checker for the CUDA
 * kernels.  It is pinned against outputs of the unmodified reference (torch
 * CPU) by tests/golden/ (see tests/golden/make_golden.py).
 *
 * What is restated (file:line into the reference checkout):
 *   homo_warping                 net/unit/base.py:85-126
 *   VectorAggregate.forward      net/unit/homoaggregate.py:25-46 (+ depth_weight :16-20,
 *                                ConvBNReLU3D net/unit/base.py:50-68, eval-mode BN)
 *   homo_aggregate_by_variance   net/unit/homoaggregate.py:49-69
 *   softmax tail of the 3-D CNN  net/unit/regular.py:67-69,130-133
 *   depth_regression             net/unit/regress.py:5-7
 *   confidence_regress           net/unit/regress.py:9-25 (+ nearest x2, net/core.py:75-77)
 *   prob conv of the 3-D CNN     net/unit/regular.py:43,67 / :110,130 (Conv3d(c0,1,3,pad=1,bias=False))
 *
 * The arithmetic of the reference lives in PyTorch (third party; the reference
 * pins torch==1.7.1, this image has 2.11.0).  The pieces of ATen that are
 * restated here are its published grid_sample algorithm
 * (ATen/native/GridSampler.h:27-35 unnormalize with align_corners=False,
 *  ATen/native/cuda/GridSampler.cuh:150-170,220 bilinear taps with zero padding)
 * plus the rounding behaviour observed on torch CPU in this image:
 *   - `tensor / python_float` is a true float32 division,
 *   - the K=3 matmul `rot @ xyz` is  fma(r2,1, fma(r1,y, r0*x)),
 *   - unnormalize is fused:          ix = fma(xn + 1, W/2, -0.5),
 *   - the 4-tap blend is             fma(se,wse, fma(sw,wsw, fma(ne,wne, nw*wnw))).
 * Every other elementwise op of the reference rounds separately, so this file
 * must be compiled with -ffp-contract=off (the Makefile does).
 *
 * Build twice: REAL=float (the oracle) and REAL=double (-DMDF_ORACLE_F64; a
 * higher-precision evaluation of the same formulae used to measure the fp32
 * noise floor of the reference itself).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef MDF_ORACLE_F64
typedef double REAL;
#define R_FMA fma
#define R_EXP exp
#define R_SQRT sqrt
#define R_FLOOR floor
#define SYM(name) name##_f64
#else
typedef float REAL;
#define R_FMA fmaf
#define R_EXP expf
#define R_SQRT sqrtf
#define R_FLOOR floorf
#define SYM(name) name##_f32
#endif

#define MDF_OK 0
#define MDF_EINVAL (-1)

/* ------------------------------------------------------------------------- */
/* proj = src_proj @ inverse(ref_proj)            (base.py:98)                */
/* LU with partial pivoting (what LAPACK getrf does for torch.inverse), then  */
/* A^-1 by solving against the identity; 4x4 product as an fma chain.         */
/* Output: 12 numbers per batch item: rot (3x3 row-major) then trans (3).     */
/* ------------------------------------------------------------------------- */
static int invert4(const REAL *a_in, REAL *inv)
{
    REAL a[16];
    int piv[4];
    memcpy(a, a_in, sizeof(a));
    for (int k = 0; k < 4; ++k) {
        int p = k;
        REAL best = fabs((double)a[k * 4 + k]);
        for (int r = k + 1; r < 4; ++r) {
            REAL v = fabs((double)a[r * 4 + k]);
            if (v > best) { best = v; p = r; }
        }
        piv[k] = p;
        if (p != k)
            for (int c = 0; c < 4; ++c) { REAL t = a[k * 4 + c]; a[k * 4 + c] = a[p * 4 + c]; a[p * 4 + c] = t; }
        if (a[k * 4 + k] == 0) return MDF_EINVAL;
        for (int r = k + 1; r < 4; ++r) {
            a[r * 4 + k] = a[r * 4 + k] / a[k * 4 + k];
            for (int c = k + 1; c < 4; ++c)
                a[r * 4 + c] = a[r * 4 + c] - a[r * 4 + k] * a[k * 4 + c];
        }
    }
    for (int col = 0; col < 4; ++col) {
        REAL b[4] = {0, 0, 0, 0};
        b[col] = 1;
        for (int k = 0; k < 4; ++k)
            if (piv[k] != k) { REAL t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
        for (int r = 1; r < 4; ++r)
            for (int c = 0; c < r; ++c) b[r] = b[r] - a[r * 4 + c] * b[c];
        for (int r = 3; r >= 0; --r) {
            for (int c = r + 1; c < 4; ++c) b[r] = b[r] - a[r * 4 + c] * b[c];
            b[r] = b[r] / a[r * 4 + r];
        }
        for (int r = 0; r < 4; ++r) inv[r * 4 + col] = b[r];
    }
    return MDF_OK;
}

int SYM(mdf_oracle_compose_proj)(const REAL *src_proj, const REAL *ref_proj, int B, REAL *rot_trans)
{
    for (int b = 0; b < B; ++b) {
        REAL inv[16], p[16];
        if (invert4(ref_proj + 16 * b, inv) != MDF_OK) return MDF_EINVAL;
        const REAL *s = src_proj + 16 * b;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) {
                REAL acc = s[r * 4 + 0] * inv[0 * 4 + c];
                for (int k = 1; k < 4; ++k) acc = R_FMA(s[r * 4 + k], inv[k * 4 + c], acc);
                p[r * 4 + c] = acc;
            }
        REAL *o = rot_trans + 12 * b;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) o[r * 3 + c] = p[r * 4 + c];
            o[9 + r] = p[r * 4 + 3];
        }
    }
    return MDF_OK;
}

/* ------------------------------------------------------------------------- */
/* Sample position of reference pixel (x,y) at depth `depth` in the source     */
/* view, in the *pixel* units grid_sample ends up using.  base.py:102-119 +    */
/* ATen unnormalize (align_corners=False).                                     */
/* ------------------------------------------------------------------------- */
typedef struct { REAL ix, iy; } sample_pos_t;

static inline sample_pos_t sample_position(const REAL *rt, int x, int y, REAL depth, int H, int W)
{
    const REAL fx = (REAL)x, fy = (REAL)y, one = 1;
    /* rot_xyz = rot @ [x, y, 1]  (base.py:110) */
    REAL rx = R_FMA(rt[2], one, R_FMA(rt[1], fy, rt[0] * fx));
    REAL ry = R_FMA(rt[5], one, R_FMA(rt[4], fy, rt[3] * fx));
    REAL rz = R_FMA(rt[8], one, R_FMA(rt[7], fy, rt[6] * fx));
    /* * depth (base.py:112), + trans (base.py:114) */
    REAL X = rx * depth, Y = ry * depth, Z = rz * depth;
    X = X + rt[9]; Y = Y + rt[10]; Z = Z + rt[11];
    /* two true divisions (base.py:115) */
    REAL px = X / Z, py = Y / Z;
    /* normalise with the align_corners=True formula (base.py:117-118) */
    REAL xn = px / (REAL)((W - 1) / 2.0) - one;
    REAL yn = py / (REAL)((H - 1) / 2.0) - one;
    /* ...but grid_sample runs with align_corners=False (base.py:122-123):
       ix = ((xn + 1) * W - 1) / 2, evaluated fused as observed on torch CPU */
    sample_pos_t s;
    s.ix = R_FMA(xn + one, (REAL)W / 2, (REAL)-0.5);
    s.iy = R_FMA(yn + one, (REAL)H / 2, (REAL)-0.5);
    return s;
}

/* positions of every (d,y,x) for one projection; pins the CUDA coordinate chain bit for bit */
int SYM(mdf_oracle_sample_positions)(const REAL *rot_trans, const REAL *hypos, int per_pixel, int D, int H, int W,
                                     REAL *ix, REAL *iy);

/* 4 bilinear taps, zero padding.  Returns 0 when no tap can be in bounds. */
typedef struct { int x0, y0; REAL wnw, wne, wsw, wse; int m_nw, m_ne, m_sw, m_se; } taps_t;

static inline int make_taps(sample_pos_t s, int H, int W, taps_t *t)
{
    /* NaN / inf / far-out positions: every tap is out of bounds -> zeros
       (GridSampler.cuh:140-147 maps non-finite coordinates to -100). */
    if (!(s.ix > -1 && s.ix < W && s.iy > -1 && s.iy < H)) return 0;
    REAL fx0 = R_FLOOR(s.ix), fy0 = R_FLOOR(s.iy);
    REAL fx1 = fx0 + 1, fy1 = fy0 + 1;
    t->x0 = (int)fx0; t->y0 = (int)fy0;
    t->wnw = (fx1 - s.ix) * (fy1 - s.iy);
    t->wne = (s.ix - fx0) * (fy1 - s.iy);
    t->wsw = (fx1 - s.ix) * (s.iy - fy0);
    t->wse = (s.ix - fx0) * (s.iy - fy0);
    int xin0 = t->x0 >= 0 && t->x0 < W, xin1 = t->x0 + 1 >= 0 && t->x0 + 1 < W;
    int yin0 = t->y0 >= 0 && t->y0 < H, yin1 = t->y0 + 1 >= 0 && t->y0 + 1 < H;
    t->m_nw = xin0 && yin0; t->m_ne = xin1 && yin0; t->m_sw = xin0 && yin1; t->m_se = xin1 && yin1;
    return 1;
}

static inline REAL blend(const REAL *plane, int W, const taps_t *t)
{
    const REAL *p = plane + (ptrdiff_t)t->y0 * W + t->x0;
    REAL nw = t->m_nw ? p[0] : 0, ne = t->m_ne ? p[1] : 0;
    REAL sw = t->m_sw ? p[W] : 0, se = t->m_se ? p[W + 1] : 0;
    return R_FMA(se, t->wse, R_FMA(sw, t->wsw, R_FMA(ne, t->wne, nw * t->wnw)));
}

static inline REAL hypo_at(const REAL *hypos, int per_pixel, int b, int d, int y, int x, int D, int H, int W)
{
    return per_pixel ? hypos[(((size_t)b * D + d) * H + y) * W + x] : hypos[(size_t)b * D + d];
}

/* ------------------------------------------------------------------------- */
/* homo_warping with a precomposed projection (rot|trans, 12 per batch item). */
/* out: (B, C, D, H, W)                                                       */
/* ------------------------------------------------------------------------- */
int SYM(mdf_oracle_homo_warp)(const REAL *src_fea, const REAL *rot_trans, const REAL *hypos, int per_pixel,
                              int B, int C, int D, int H, int W, REAL *out)
{
    if (B < 0 || C < 0 || D < 0 || H < 0 || W < 0) return MDF_EINVAL;
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
        const REAL *rt = rot_trans + 12 * b;
#pragma omp parallel for collapse(2) schedule(static)
        for (int d = 0; d < D; ++d)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    REAL depth = hypo_at(hypos, per_pixel, b, d, y, x, D, H, W);
                    taps_t t;
                    int ok = make_taps(sample_position(rt, x, y, depth, H, W), H, W, &t);
                    for (int c = 0; c < C; ++c) {
                        const REAL *plane = src_fea + ((size_t)b * C + c) * HW;
                        out[((((size_t)b * C + c) * D + d) * H + y) * W + x] = ok ? blend(plane, W, &t) : 0;
                    }
                }
    }
    return MDF_OK;
}

int SYM(mdf_oracle_sample_positions)(const REAL *rot_trans, const REAL *hypos, int per_pixel, int D, int H, int W,
                                     REAL *ix, REAL *iy)
{
    for (int d = 0; d < D; ++d)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t i = ((size_t)d * H + y) * W + x;
                sample_pos_t s = sample_position(rot_trans, x, y, per_pixel ? hypos[i] : hypos[d], H, W);
                ix[i] = s.ix; iy[i] = s.iy;
            }
    return MDF_OK;
}

/* softmax over n values with stride `st` (max-subtracted, as ATen does) */
static inline void softmax_n(const REAL *in, int n, REAL *out)
{
    REAL m = in[0];
    for (int k = 1; k < n; ++k) m = in[k] > m ? in[k] : m;
    REAL sum = 0;
    for (int k = 0; k < n; ++k) { out[k] = R_EXP(in[k] - m); sum = sum + out[k]; }
    for (int k = 0; k < n; ++k) out[k] = out[k] / sum;
}

/* ------------------------------------------------------------------------- */
/* VectorAggregate.forward, eval mode (homoaggregate.py:25-46).               */
/*   features: V = N pointers, each (B,C,H,W); features[0] is the reference.    */
/*   rot_trans: (N-1) pointers, each 12*B (from compose_proj).                 */
/*   depth-weight parameters (homoaggregate.py:16-20, state-dict names):       */
/*     cw[G]  = depth_weight.0.conv.weight      bn = {weight,bias,mean,var}    */
/*     fc_w, fc_b = depth_weight.1.{weight,bias}                               */
/*   out: (B,G,D,H,W)                                                          */
/* ------------------------------------------------------------------------- */
/* rows [y0, y1) only (bench.py's bounded CPU sample); the other rows of `out` are left untouched */
int SYM(mdf_oracle_vector_aggregate_rows)(const REAL *const *features, const REAL *const *rot_trans, int N,
                                          const REAL *hypos, int per_pixel,
                                          const REAL *cw, REAL bn_w, REAL bn_b, REAL bn_mean, REAL bn_var, REAL bn_eps,
                                          REAL fc_w, REAL fc_b,
                                          int B, int C, int G, int D, int H, int W, int y0, int y1, REAL *out)
{
    if (N < 2 || G <= 0 || C % G != 0 || C / G > 16 || y0 < 0 || y1 > H) return MDF_EINVAL;
    const int cpg = C / G;
    const size_t HW = (size_t)H * W;
    /* eval-mode BatchNorm3d(1) folded the way ATen's CPU kernel applies it */
    const REAL invstd = 1 / R_SQRT(bn_var + bn_eps);
    const REAL alpha = invstd * bn_w;
    const REAL beta = bn_b - bn_mean * alpha;
    for (int b = 0; b < B; ++b) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int d = 0; d < D; ++d)
            for (int y = y0; y < y1; ++y) {
                REAL *vol = (REAL *)malloc(sizeof(REAL) * G);
                REAL *vsum = (REAL *)malloc(sizeof(REAL) * G);
                for (int x = 0; x < W; ++x) {
                    REAL depth = hypo_at(hypos, per_pixel, b, d, y, x, D, H, W);
                    REAL wsum = 0;
                    for (int g = 0; g < G; ++g) vsum[g] = 0;
                    for (int v = 1; v < N; ++v) {
                        taps_t t;
                        int ok = make_taps(sample_position(rot_trans[v - 1] + 12 * b, x, y, depth, H, W), H, W, &t);
                        REAL z = 0;
                        for (int g = 0; g < G; ++g) {
                            REAL rv[16], sv[16], rp[16], sp[16];
                            for (int k = 0; k < cpg; ++k) {
                                size_t ch = ((size_t)b * C + (size_t)g * cpg + k) * HW;
                                rv[k] = features[0][ch + (size_t)y * W + x];
                                sv[k] = ok ? blend(features[v] + ch, W, &t) : 0;
                            }
                            softmax_n(rv, cpg, rp);   /* homoaggregate.py:32 */
                            softmax_n(sv, cpg, sp);   /* homoaggregate.py:38 */
                            REAL dot = 0;
                            for (int k = 0; k < cpg; ++k) dot = dot + sp[k] * rp[k]; /* :39 */
                            vol[g] = dot;
                            z = z + cw[g] * dot;      /* Conv3d(G,1,k=1), no bias */
                        }
                        REAL a = z * alpha + beta;    /* BatchNorm3d eval */
                        a = a > 0 ? a : 0;            /* ReLU */
                        a = a * fc_w + fc_b;          /* Conv3d(1,1,k=1) */
                        REAL w = 1 / (1 + R_EXP(-a)); /* Sigmoid */
                        wsum = wsum + w;              /* :41 */
                        for (int g = 0; g < G; ++g) vsum[g] = vsum[g] + w * vol[g]; /* :42 */
                    }
                    for (int g = 0; g < G; ++g)
                        out[((((size_t)b * G + g) * D + d) * H + y) * W + x] = vsum[g] / wsum; /* :46 */
                }
                free(vol); free(vsum);
            }
    }
    return MDF_OK;
}

int SYM(mdf_oracle_vector_aggregate)(const REAL *const *features, const REAL *const *rot_trans, int N,
                                     const REAL *hypos, int per_pixel,
                                     const REAL *cw, REAL bn_w, REAL bn_b, REAL bn_mean, REAL bn_var, REAL bn_eps,
                                     REAL fc_w, REAL fc_b,
                                     int B, int C, int G, int D, int H, int W, REAL *out)
{
    return SYM(mdf_oracle_vector_aggregate_rows)(features, rot_trans, N, hypos, per_pixel, cw, bn_w, bn_b, bn_mean,
                                                 bn_var, bn_eps, fc_w, fc_b, B, C, G, D, H, W, 0, H, out);
}

/* ------------------------------------------------------------------------- */
/* homo_aggregate_by_variance (homoaggregate.py:49-69).  out: (B,C,D,H,W)     */
/* ------------------------------------------------------------------------- */
int SYM(mdf_oracle_variance_aggregate)(const REAL *const *features, const REAL *const *rot_trans, int N,
                                       const REAL *hypos, int per_pixel,
                                       int B, int C, int D, int H, int W, REAL *out)
{
    if (N < 2 || C <= 0) return MDF_EINVAL;
    const size_t HW = (size_t)H * W;
    const REAL nv = (REAL)N;
    for (int b = 0; b < B; ++b) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int d = 0; d < D; ++d)
            for (int y = 0; y < H; ++y) {
                REAL *s1 = (REAL *)malloc(sizeof(REAL) * C), *s2 = (REAL *)malloc(sizeof(REAL) * C);
                REAL *wv = (REAL *)malloc(sizeof(REAL) * C), *sm = (REAL *)malloc(sizeof(REAL) * C);
                for (int x = 0; x < W; ++x) {
                    REAL depth = hypo_at(hypos, per_pixel, b, d, y, x, D, H, W);
                    for (int c = 0; c < C; ++c) {
                        REAL r = features[0][((size_t)b * C + c) * HW + (size_t)y * W + x];
                        s1[c] = r; s2[c] = r * r;       /* :56 (raw reference feature) */
                    }
                    for (int v = 1; v < N; ++v) {
                        taps_t t;
                        int ok = make_taps(sample_position(rot_trans[v - 1] + 12 * b, x, y, depth, H, W), H, W, &t);
                        for (int c = 0; c < C; ++c)
                            wv[c] = ok ? blend(features[v] + ((size_t)b * C + c) * HW, W, &t) : 0;
                        softmax_n(wv, C, sm);           /* :60 softmax over all channels */
                        for (int c = 0; c < C; ++c) { s1[c] = s1[c] + sm[c]; s2[c] = s2[c] + sm[c] * sm[c]; }
                    }
                    for (int c = 0; c < C; ++c) {
                        REAL mean = s1[c] / nv;
                        out[((((size_t)b * C + c) * D + d) * H + y) * W + x] = s2[c] / nv - mean * mean; /* :66 */
                    }
                }
                free(s1); free(s2); free(wv); free(sm);
            }
    }
    return MDF_OK;
}

/* ------------------------------------------------------------------------- */
/* Head.  softmax over D (regular.py:69,133), depth_regression (regress.py:5-7) */
/* and confidence_regress (regress.py:9-25) + nearest x`up` (core.py:75-77).   */
/* ------------------------------------------------------------------------- */
int SYM(mdf_oracle_softmax_depth)(const REAL *logits, int B, int D, int H, int W, REAL *prob)
{
    const size_t HW = (size_t)H * W;
    if (D <= 0) return MDF_EINVAL;
    for (int b = 0; b < B; ++b)
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < HW; ++p) {
            const REAL *in = logits + (size_t)b * D * HW + p;
            REAL *o = prob + (size_t)b * D * HW + p;
            REAL m = in[0];
            for (int d = 1; d < D; ++d) m = in[d * HW] > m ? in[d * HW] : m;
            REAL sum = 0;
            for (int d = 0; d < D; ++d) { o[d * HW] = R_EXP(in[d * HW] - m); sum = sum + o[d * HW]; }
            for (int d = 0; d < D; ++d) o[d * HW] = o[d * HW] / sum;
        }
    return MDF_OK;
}

int SYM(mdf_oracle_depth_regression)(const REAL *prob, const REAL *hypos, int per_pixel,
                                     int B, int D, int H, int W, REAL *depth)
{
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b)
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < HW; ++p) {
            REAL acc = 0;
            for (int d = 0; d < D; ++d) {
                REAL h = per_pixel ? hypos[((size_t)b * D + d) * HW + p] : hypos[(size_t)b * D + d];
                acc = acc + prob[((size_t)b * D + d) * HW + p] * h;
            }
            depth[(size_t)b * HW + p] = acc;
        }
    return MDF_OK;
}

/* window sum S[k] = n * avg_pool(pad_D(prob, pad_front, pad_back), n)[k],
   gathered at k = trunc(sum_d prob[d] * d); result replicated up x up. */
int SYM(mdf_oracle_confidence)(const REAL *prob, int B, int D, int H, int W,
                               int n, int pad_front, int pad_back, int up, REAL *conf)
{
    const size_t HW = (size_t)H * W;
    const int Dp = D + pad_front + pad_back - n + 1;
    if (n <= 0 || up <= 0 || Dp <= 0) return MDF_EINVAL;
    int status = MDF_OK;
    for (int b = 0; b < B; ++b)
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const REAL *p = prob + (size_t)b * D * HW + (size_t)y * W + x;
                REAL e = 0;
                for (int d = 0; d < D; ++d) e = e + p[d * HW] * (REAL)d;
                long k = (long)e;                    /* .long() truncates */
                REAL c;
                if (k < 0 || k >= Dp) { c = 0; status = MDF_EINVAL; } /* torch.gather would raise */
                else {
                    REAL s = 0;
                    for (int j = 0; j < n; ++j) {
                        long dd = k - pad_front + j;
                        s = s + ((dd >= 0 && dd < D) ? p[dd * HW] : 0);
                    }
                    c = (REAL)n * (s / (REAL)n);
                }
                for (int uy = 0; uy < up; ++uy)
                    for (int ux = 0; ux < up; ++ux)
                        conf[((size_t)b * H * up + (size_t)y * up + uy) * ((size_t)W * up) + (size_t)x * up + ux] = c;
            }
    return status;
}

/* ------------------------------------------------------------------------- */
/* Depth hypotheses of the next stage: HyposByFit (net/unit/depthhypos.py).    */
/*   stage 0     : uniform hypotheses                       depthhypos.py:31-38 */
/*   fit         : per-pixel curve fit of the probability column               */
/*                 "gauss1"  depthhypos.py:169-215  s = |-1/b0|, ln p ~ b0 x^2 + b1 x + b2 */
/*                 "laplace" depthhypos.py:78-125   s = 1/|sum(x y)/sum(x x)|, x = |hypo - depth| */
/*                 "gauss0"  depthhypos.py:127-167  s = |-1/b0|, ln p ~ b0 x + b1, x = (hypo - depth)^2 */
/*   generate    : bilinear x2 upsampling of s and depth (F.interpolate, align_corners=False), */
/*                 search range from s and prob_thresh, clamps, D' hypotheses   depthhypos.py:48-76 */
/* The gauss1 normal equations (entries up to 935^4 * 48) are hopeless in float32: the reference's */
/* own float32 result is ~2e-4 (median) / 3e-2 (max) away from a float64 evaluation.  The restatement */
/* therefore solves them in long double after centring x (the quadratic coefficient is invariant */
/* under shifts), for both REAL types: it is the yardstick, not a rounding-for-rounding copy. */
/* ------------------------------------------------------------------------- */
int SYM(mdf_oracle_hypos_fit)(const REAL *prob, const REAL *hypos, int per_pixel, const REAL *depth,
                              int mode /* 1 = gauss1, 2 = laplace, 3 = gauss0 */, int B, int D, int H, int W, REAL *s_out)
{
    const size_t HW = (size_t)H * W;
    if (mode != 1 && mode != 2 && mode != 3) return MDF_EINVAL;
    for (int b = 0; b < B; ++b)
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < HW; ++p) {
            if (mode == 2) {
                REAL sxy = 0, sxx = 0;
                for (int d = 0; d < D; ++d) {
                    REAL pr = prob[((size_t)b * D + d) * HW + p];
                    if (pr < (REAL)1e-40) pr = (REAL)1e-40;
                    REAL h = per_pixel ? hypos[((size_t)b * D + d) * HW + p] : hypos[(size_t)b * D + d];
                    REAL x = (REAL)fabs((double)(h - depth[(size_t)b * HW + p]));
                    REAL y = (REAL)log((double)pr);
                    sxy = sxy + x * y; sxx = sxx + x * x;
                }
                REAL bb = (REAL)fabs((double)(sxy / sxx));
                s_out[(size_t)b * HW + p] = 1 / bb;
            } else if (mode == 3) {
                /* regression of z = ln p on [x, 1] with x = (hypo - depth)^2 formed in REAL as the reference does
                 * (depthhypos.py:153); the 2x2 normal equations (entries up to 510^4 * 48) are solved in long double
                 * after centring x: the slope is b0 = S_xz / S_xx */
                long double xm = 0, zm = 0;
                for (int d = 0; d < D; ++d) {
                    REAL h = per_pixel ? hypos[((size_t)b * D + d) * HW + p] : hypos[(size_t)b * D + d];
                    REAL df = h - depth[(size_t)b * HW + p];
                    xm += (long double)(REAL)(df * df);
                }
                xm /= D;
                long double sxx = 0, sxz = 0;
                for (int d = 0; d < D; ++d) {
                    REAL pr = prob[((size_t)b * D + d) * HW + p];
                    if (pr < (REAL)1e-40) pr = (REAL)1e-40;
                    REAL h = per_pixel ? hypos[((size_t)b * D + d) * HW + p] : hypos[(size_t)b * D + d];
                    REAL df = h - depth[(size_t)b * HW + p];
                    long double u = (long double)(REAL)(df * df) - xm;
                    sxx += u * u; sxz += u * logl((long double)pr);
                }
                (void)zm;
                s_out[(size_t)b * HW + p] = (REAL)fabsl(-1.0L / (sxz / sxx));
            } else {
                long double mean = 0;
                for (int d = 0; d < D; ++d) mean += per_pixel ? hypos[((size_t)b * D + d) * HW + p] : hypos[(size_t)b * D + d];
                mean /= D;
                long double m[5] = {0, 0, 0, 0, 0}, r[3] = {0, 0, 0};    /* sum u^k, sum u^k z */
                for (int d = 0; d < D; ++d) {
                    REAL pr = prob[((size_t)b * D + d) * HW + p];
                    if (pr < (REAL)1e-40) pr = (REAL)1e-40;
                    long double z = logl((long double)pr);
                    long double u = (long double)(per_pixel ? hypos[((size_t)b * D + d) * HW + p] : hypos[(size_t)b * D + d]) - mean;
                    long double uk = 1;
                    for (int k = 0; k < 5; ++k) { m[k] += uk; if (k < 3) r[k] += uk * z; uk *= u; }
                }
                /* normal equations for [c2, c1, c0]: [[m4 m3 m2][m3 m2 m1][m2 m1 m0]] c = [r2 r1 r0]; Cramer for c2 */
                long double a11 = m[4], a12 = m[3], a13 = m[2], a22 = m[2], a23 = m[1], a33 = m[0];
                long double det = a11 * (a22 * a33 - a23 * a23) - a12 * (a12 * a33 - a23 * a13) + a13 * (a12 * a23 - a22 * a13);
                long double d2 = r[2] * (a22 * a33 - a23 * a23) - a12 * (r[1] * a33 - a23 * r[0]) + a13 * (r[1] * a23 - a22 * r[0]);
                long double c2 = d2 / det;
                s_out[(size_t)b * HW + p] = (REAL)fabsl(-1.0L / c2);
            }
        }
    return MDF_OK;
}

/* F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) of one (H,W) map at fine pixel (Y,X) */
static inline REAL up2(const REAL *m, int H, int W, int Y, int X)
{
    REAL sy = ((REAL)Y + (REAL)0.5) * (REAL)0.5 - (REAL)0.5, sx = ((REAL)X + (REAL)0.5) * (REAL)0.5 - (REAL)0.5;
    if (sy < 0) sy = 0;
    if (sx < 0) sx = 0;
    int y0 = (int)sy, x0 = (int)sx;
    int y1 = y0 + (y0 < H - 1), x1 = x0 + (x0 < W - 1);
    REAL ly = sy - (REAL)y0, lx = sx - (REAL)x0, hy = 1 - ly, hx = 1 - lx;
    return hy * (hx * m[(size_t)y0 * W + x0] + lx * m[(size_t)y0 * W + x1]) + ly * (hx * m[(size_t)y1 * W + x0] + lx * m[(size_t)y1 * W + x1]);
}

/* depth (B,H,W), s (B,H,W), depth_range (B,2) -> hypotheses (B, ND, 2H, 2W) (upsample != 0) or (B, ND, H, W) */
int SYM(mdf_oracle_hypos_generate)(const REAL *depth, const REAL *s, const REAL *depth_range, int mode, REAL prob_thresh,
                                   int upsample, int B, int H, int W, int ND, REAL *out)
{
    if (mode != 1 && mode != 2 && mode != 3) return MDF_EINVAL;
    const int Ho = upsample ? 2 * H : H, Wo = upsample ? 2 * W : W;
    REAL gmax = depth_range[1], gmin = depth_range[0];
    for (int b = 1; b < B; ++b) {
        if (depth_range[2 * b + 1] > gmax) gmax = depth_range[2 * b + 1];
        if (depth_range[2 * b] < gmin) gmin = depth_range[2 * b];
    }
    const REAL cap_all = (gmax - gmin) / 2;
    const REAL lt = (REAL)log((double)prob_thresh);
    for (int b = 0; b < B; ++b) {
        const REAL dmin = depth_range[2 * b], dmax = depth_range[2 * b + 1];
        const REAL cap_b = (dmax - dmin) * (REAL)0.2;
#pragma omp parallel for schedule(static)
        for (int Y = 0; Y < Ho; ++Y)
            for (int X = 0; X < Wo; ++X) {
                REAL sv = upsample ? up2(s + (size_t)b * H * W, H, W, Y, X) : s[((size_t)b * H + Y) * W + X];
                REAL dv = upsample ? up2(depth + (size_t)b * H * W, H, W, Y, X) : depth[((size_t)b * H + Y) * W + X];
                REAL res = mode != 2 ? R_SQRT(-1 * sv * lt) : (REAL)fabs((double)(sv * lt));    /* gauss0 | gauss1 : laplace */
                if (res < (REAL)1e-6) res = (REAL)1e-6;          /* clamp(min, max): NaN propagates like torch */
                if (res > cap_all) res = cap_all;
                if (res > cap_b) res = cap_b;
                const REAL interval = res / (REAL)(ND - 1);
                const REAL base = dv - (REAL)0.5 * res;
                for (int d = 0; d < ND; ++d) {
                    REAL h = base + interval * (REAL)d;
                    REAL delta = h - dmin; if (delta < 0) delta = 0; h = dmin + delta;
                    delta = h - dmax; if (delta > 0) delta = 0; h = dmax + delta;
                    out[(((size_t)b * ND + d) * Ho + Y) * Wo + X] = h;
                }
            }
    }
    return MDF_OK;
}

/* ---- last layer of the regulariser: x = self.prob(x).squeeze(1)  (net/unit/regular.py:43,67 and :110,130) --------
 * nn.Conv3d(c0, 1, 3, stride=1, padding=1, bias=False): cross-correlation with zero padding,
 *   logits[b][d][y][x] = sum_{c,kd,ky,kx} w[c][kd][ky][kx] * x[b][c][d+kd-1][y+ky-1][x+kx-1].
 * torch's CPU convolution (oneDNN / slow_conv3d) does not document its summation order; this restatement sums in
 * (c, kd, ky, kx) order with one rounding per multiply-add pair (separate mul and add), and the float64 build is the
 * yardstick for what any float32 order can claim. */
int SYM(mdf_oracle_prob_conv)(const REAL *x, const REAL *w, int B, int C, int D, int H, int W, REAL *logits)
{
    if (B < 0 || C < 1 || D < 0 || H < 0 || W < 0) return MDF_EINVAL;
    const size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d)
            for (int y = 0; y < H; ++y)
                for (int xx = 0; xx < W; ++xx) {
                    REAL acc = 0;
                    for (int c = 0; c < C; ++c)
                        for (int kd = 0; kd < 3; ++kd) {
                            const int dz = d + kd - 1;
                            if (dz < 0 || dz >= D) continue;
                            for (int ky = 0; ky < 3; ++ky) {
                                const int yy = y + ky - 1;
                                if (yy < 0 || yy >= H) continue;
                                for (int kx = 0; kx < 3; ++kx) {
                                    const int xs = xx + kx - 1;
                                    if (xs < 0 || xs >= W) continue;
                                    acc += w[((c * 3 + kd) * 3 + ky) * 3 + kx] * x[(((size_t)b * C + c) * D + dz) * HW + (size_t)yy * W + xs];
                                }
                            }
                        }
                    logits[((size_t)b * D + d) * HW + (size_t)y * W + xx] = acc;
                }
    return MDF_OK;
}

/* ---- geometric-consistency filter (post-processing), tools/filter/dynamic_filter_gpu.py ---------------------------
 * reproject_with_depth (:184-237), check_geometric_consistency (:161-182), the per-view aggregation of filter() (:57-100)
 * and bilinear_sampler (tools/filter/data_io.py:117-131: grid_sample, bilinear, zero padding, align_corners=True).
 * Matrices: A^-1 by LU with partial pivoting (torch.inverse), products as fma chains over K = 3 / 4, every other
 * elementwise op rounded separately.  cuBLAS / cuSOLVER (the reference runs this on the GPU) do not document their
 * summation order; the decisions are thresholds on quantities that are smooth in the inputs, so implementations agree
 * except within float32 noise of a threshold. */
static int invert3(const REAL *k, REAL *inv)
{
    REAL a[16] = {k[0], k[1], k[2], 0, k[3], k[4], k[5], 0, k[6], k[7], k[8], 0, 0, 0, 0, 1}, o[16];
    if (invert4(a, o) != MDF_OK) return MDF_EINVAL;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) inv[r * 3 + c] = o[r * 4 + c];
    return MDF_OK;
}

static void matmul4(const REAL *a, const REAL *b, REAL *o)
{
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            REAL acc = a[r * 4 + 0] * b[0 * 4 + c];
            for (int k = 1; k < 4; ++k) acc = R_FMA(a[r * 4 + k], b[k * 4 + c], acc);
            o[r * 4 + c] = acc;
        }
}

static inline void mat3v(const REAL *m, REAL x, REAL y, REAL z, REAL *o)
{
    for (int r = 0; r < 3; ++r) o[r] = R_FMA(m[r * 3 + 2], z, R_FMA(m[r * 3 + 1], y, m[r * 3 + 0] * x));
}

static inline void mat34v(const REAL *m, const REAL *v, REAL *o)     /* rows 0-2 of a 4x4 times [v;1] */
{
    for (int r = 0; r < 3; ++r) o[r] = R_FMA(m[r * 4 + 3], (REAL)1, R_FMA(m[r * 4 + 2], v[2], R_FMA(m[r * 4 + 1], v[1], m[r * 4 + 0] * v[0])));
}

/* grid_sample(bilinear, zeros, align_corners=True) of one plane at pixel coordinates (px, py) through
 * bilinear_sampler's normalise / ATen's unnormalise pair */
static inline REAL sample_depth_ac(const REAL *img, int H, int W, REAL px, REAL py)
{
    const REAL gx = (REAL)2 * px / (REAL)(W - 1) - (REAL)1, gy = (REAL)2 * py / (REAL)(H - 1) - (REAL)1;
    const REAL ix = ((gx + (REAL)1) / (REAL)2) * (REAL)(W - 1), iy = ((gy + (REAL)1) / (REAL)2) * (REAL)(H - 1);
    if (!(ix > (REAL)-1 && ix < (REAL)W && iy > (REAL)-1 && iy < (REAL)H)) return 0;       /* also NaN / inf */
    const REAL fx = R_FLOOR(ix), fy = R_FLOOR(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const REAL ax = (fx + 1) - ix, bx = ix - fx, ay = (fy + 1) - iy, by = iy - fy;
    const REAL nw = (x0 >= 0 && y0 >= 0 && x0 < W && y0 < H) ? img[(size_t)y0 * W + x0] : 0;
    const REAL ne = (x0 + 1 >= 0 && y0 >= 0 && x0 + 1 < W && y0 < H) ? img[(size_t)y0 * W + x0 + 1] : 0;
    const REAL sw = (x0 >= 0 && y0 + 1 >= 0 && x0 < W && y0 + 1 < H) ? img[(size_t)(y0 + 1) * W + x0] : 0;
    const REAL se = (x0 + 1 >= 0 && y0 + 1 >= 0 && x0 + 1 < W && y0 + 1 < H) ? img[(size_t)(y0 + 1) * W + x0 + 1] : 0;
    return R_FMA(se, bx * by, R_FMA(sw, ax * by, R_FMA(ne, bx * ay, nw * (ax * ay))));
}

/* Outputs (any may be NULL): per source view the 9 dynamic masks as bits 0-8 of bits[s][p] (bit i-2 <-> threshold i,
 * :176-179) and depth_reprojected (zeroed where the last mask fails, :180); fused: depth_est_averaged (:93),
 * geo / photo / final masks (:86-98). */
int SYM(mdf_oracle_geo_filter)(const REAL *ref_depth, const REAL *ref_K, const REAL *ref_E, const REAL *const *src_depths,
                               const REAL *src_K, const REAL *src_E, int S, int H, int W, const REAL *confidence,
                               REAL photo_threshold, int nconditions, REAL thre1, REAL thre2,
                               uint16_t *bits, REAL *depth_reprojected, REAL *depth_averaged,
                               uint8_t *geo_mask, uint8_t *photo_mask, uint8_t *final_mask)
{
    if (S < 0 || H < 0 || W < 0) return MDF_EINVAL;
    const size_t HW = (size_t)H * W;
    REAL *mats = (REAL *)malloc(sizeof(REAL) * (size_t)(S > 0 ? S : 1) * 64);
    REAL Ainv[9], Erinv[16];
    if (!mats) return MDF_EINVAL;
    if (invert3(ref_K, Ainv) != MDF_OK || invert4(ref_E, Erinv) != MDF_OK) { free(mats); return MDF_EINVAL; }
    for (int s = 0; s < S; ++s) {
        REAL *m = mats + 64 * s, Esinv[16];
        if (invert4(src_E + 16 * s, Esinv) != MDF_OK || invert3(src_K + 9 * s, m + 32) != MDF_OK) { free(mats); return MDF_EINVAL; }
        matmul4(src_E + 16 * s, Erinv, m);            /* ref camera -> src camera   (:200) */
        matmul4(ref_E, Esinv, m + 16);                /* src camera -> ref camera   (:221) */
    }
#pragma omp parallel for schedule(static)
    for (long long p = 0; p < (long long)HW; ++p) {
        const int y = (int)(p / W), x = (int)(p % W);
        const REAL d = ref_depth[p];
        int counts[9] = {0};
        int nvalid = 0;
        REAL dsum = 0;
        for (int s = 0; s < S; ++s) {
            const REAL *m = mats + 64 * s;
            REAL v[3], q[3], k[3];
            mat3v(Ainv, (REAL)x * d, (REAL)y * d, (REAL)1 * d, v);             /* :196-198 */
            mat34v(m, v, q);                                                   /* :200-201 */
            mat3v(src_K + 9 * s, q[0], q[1], q[2], k);                          /* :203 */
            const REAL xs = k[0] / k[2], ys = k[1] / k[2];                      /* :204 */
            const REAL ds = sample_depth_ac(src_depths[s], H, W, xs, ys);      /* :212 */
            mat3v(m + 32, xs * ds, ys * ds, (REAL)1 * ds, v);                   /* :216-217 */
            mat34v(m + 16, v, q);                                              /* :219-220 */
            const REAL drep = q[2];                                            /* :222 */
            mat3v(ref_K, q[0], q[1], q[2], k);                                 /* :223 */
            const REAL xr = k[0] / k[2], yr = k[1] / k[2];                      /* :224 */
            const REAL dx = xr - (REAL)x, dy = yr - (REAL)y;
            const REAL dist = R_SQRT(dx * dx + dy * dy);                        /* :170 */
            const REAL rel = (REAL)fabs((double)(drep - d)) / d;                /* :173-174 */
            unsigned b = 0;
            for (int i = 2; i < 11; ++i)
                if (dist < (REAL)i / thre1 && rel < (REAL)i / thre2) { b |= 1u << (i - 2); counts[i - 2]++; }
            const int last = (b >> 8) & 1;
            if (bits) bits[(size_t)s * HW + p] = (uint16_t)b;
            if (depth_reprojected) depth_reprojected[(size_t)s * HW + p] = last ? drep : 0;
            if (last) { nvalid++; dsum = dsum + drep; }
        }
        int geo = 0;
        for (int i = 2; i < 11; ++i) geo += counts[i - 2] >= i;
        const int g = S > 0 && geo >= nconditions;
        const int ph = confidence ? confidence[p] > photo_threshold : 1;
        if (depth_averaged) depth_averaged[p] = (dsum + d) / (REAL)(nvalid + 1);
        if (geo_mask) geo_mask[p] = (uint8_t)g;
        if (photo_mask) photo_mask[p] = (uint8_t)ph;
        if (final_mask) final_mask[p] = (uint8_t)(g && ph);
    }
    free(mats);
    return MDF_OK;
}

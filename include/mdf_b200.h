/*
 * mdf_b200.h -- C ABI of the B200-native plane-sweep cost-volume path of MDF-Net.
 *
 * One shared library (mdf_net_b200/libmdf_b200.so, sm_100a only) exports these entry points.
 * They are what a binding for the reference's hot path would call; the reference itself is pure
 * Python/PyTorch, so each entry replaces a *Python call site* (file:line into the reference):
 *
 *   mdf_homo_warp_fwd          net/unit/base.py:85-126           homo_warping(...)
 *   mdf_cost_volume_fwd        net/unit/homoaggregate.py:25-46   VectorAggregate.forward (eval-mode BN)
 *                              (+ :16-20 depth_weight, net/unit/base.py:50-68 ConvBNReLU3D)
 *   mdf_variance_volume_fwd    net/unit/homoaggregate.py:49-69   homo_aggregate_by_variance(...)
 *   mdf_softmax_regress_fwd    net/unit/regular.py:67-69,130-133 F.softmax(x, dim=1)
 *                              + net/unit/regress.py:5-7         depth_regression(...)
 *                              + net/unit/regress.py:9-25        confidence_regress(...)   (optional)
 *   mdf_depth_regression_fwd   net/unit/regress.py:5-7
 *   mdf_confidence_fwd         net/unit/regress.py:9-25 (+ nearest upsample net/core.py:75-77)
 *
 * Conventions:
 *   - all tensors float32, contiguous, NCHW / NCDHW (W fastest), DEVICE pointers unless noted;
 *   - the library allocates nothing and keeps no state between calls (the one exception is errno-like: the
 *     cudaError_t behind the calling thread's last MDF_ERR_CUDA, mdf_last_cuda_error): the caller owns inputs,
 *     outputs and the scratch `workspace` (size from the *_workspace_bytes query, 256-byte aligned);
 *   - test / benchmark / tuning entry points live in mdf_b200_debug.h, not here;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and runs
 *     on the device that owns the output pointer; nothing synchronises the device;
 *   - return value: MDF_OK or a negative status; no exceptions, no exit().  There is NO CPU
 *     fallback: host pointers are rejected with MDF_ERR_NOT_DEVICE.
 *   - depth hypotheses are (B,D,1,1) (`hypos_per_pixel` = 0) or (B,D,H,W) (`hypos_per_pixel` = 1);
 *   - projections are the 4x4 matrices scale_cam returns (net/unit/scale.py:4-20), (B,4,4) each.
 */
#ifndef MDF_B200_H_
#define MDF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDF_ABI_VERSION 1

#define MDF_OK 0
#define MDF_ERR_INVALID_SHAPE (-1)   /* negative / inconsistent sizes, C % G != 0, N < 2 ... */
#define MDF_ERR_UNSUPPORTED (-2)     /* valid but not implemented (e.g. more than MDF_MAX_VIEWS views) */
#define MDF_ERR_NULL_POINTER (-3)
#define MDF_ERR_WORKSPACE (-4)       /* workspace missing, misaligned or too small */
#define MDF_ERR_NOT_DEVICE (-5)      /* a data pointer is not device memory */
#define MDF_ERR_CUDA (-6)            /* launch / driver error; see mdf_last_cuda_error() */

#define MDF_MAX_VIEWS 32             /* reference + 31 sources (reference default: 5 / 11) */

typedef void *mdf_stream_t;          /* cudaStream_t */

#if defined(__GNUC__)
#define MDF_API __attribute__((visibility("default")))
#else
#define MDF_API
#endif

MDF_API int mdf_abi_version(void);
MDF_API const char *mdf_status_string(int status);
/* cudaError_t of the most recent MDF_ERR_CUDA on the calling thread (0 if none). */
MDF_API int mdf_last_cuda_error(void);

/* ---- homo_warping: one source view -> (B,C,D,H,W) ------------------------------------------ */
MDF_API size_t mdf_homo_warp_workspace_bytes(int B);

MDF_API int mdf_homo_warp_fwd(const float *src_fea,        /* (B,C,H,W) */
                      const float *src_proj,       /* (B,4,4)   */
                      const float *ref_proj,       /* (B,4,4)   */
                      const float *depth_hypos, int hypos_per_pixel,
                      int B, int C, int D, int H, int W,
                      float *warped,               /* (B,C,D,H,W) */
                      void *workspace, size_t workspace_bytes,
                      mdf_stream_t stream);

/* ---- VectorAggregate.forward, eval mode -> cost volume (B,G,D,H,W) -------------------------- */
MDF_API size_t mdf_cost_volume_workspace_bytes(int B, int N, int C, int G, int D, int H, int W);

MDF_API int mdf_cost_volume_fwd(const float *const *features,   /* HOST array of N device ptrs, each (B,C,H,W); [0] = reference view */
                        int N,
                        const float *ref_proj,          /* (B,4,4) */
                        const float *const *src_projs,  /* HOST array of N-1 device ptrs, each (B,4,4) */
                        const float *depth_hypos, int hypos_per_pixel,
                        /* depth_weight parameters, device pointers (state-dict names in comments) */
                        const float *conv_weight,       /* depth_weight.0.conv.weight  (1,G,1,1,1) */
                        const float *bn_weight,         /* depth_weight.0.bn.weight        (1,) */
                        const float *bn_bias,           /* depth_weight.0.bn.bias          (1,) */
                        const float *bn_mean,           /* depth_weight.0.bn.running_mean  (1,) */
                        const float *bn_var,            /* depth_weight.0.bn.running_var   (1,) */
                        float bn_eps,
                        const float *fc_weight,         /* depth_weight.1.weight (1,1,1,1,1) */
                        const float *fc_bias,           /* depth_weight.1.bias   (1,) */
                        int B, int C, int G, int D, int H, int W,
                        float *cost_volume,             /* (B,G,D,H,W) */
                        void *workspace, size_t workspace_bytes,
                        mdf_stream_t stream);

/* ---- optional fast entry: the FPN hand-off ---------------------------------------
 * The reference's backbone ends every scale with a bias-free 1x1 convolution (net/unit/backbone.py:43-45, 59-63):
 * features[k] = out_k(x_k).  mdf_fpn_out_prepped_fwd IS that convolution for one view, writing the hot kernel's own input
 * layout instead of NCHW features: per float4 plane j (groups 4j..4j+3) and pixel
 *     source view (s4 != NULL, q4 = cq4 = NULL):   s4 = (y[2g+1] - y[2g]) * log2(e)              (B, G/4, H, W, 4)
 *     reference view (s4 == NULL):                 q4 = 2*sigmoid(y[2g] - y[2g+1]) - 1,  cq4 = depth_weight_conv[g] * q4
 * with y = out_weight (2G, Cin) applied to x (B, Cin, H, W).  mdf_cost_volume_fwd_prepped then computes the same cost
 * volume as mdf_cost_volume_fwd from those maps (s4: the N-1 source views back to back, (N-1, B, G/4, H, W, 4)) without
 * the layout pass.  C == 2G with G in {8, 16, 32}; Cin a multiple of 4.  The NCHW entry point above stays the drop-in. */
MDF_API int mdf_fpn_out_prepped_fwd(const float *x, const float *out_weight, int B, int Cin, int G, int H, int W,
                                    const float *depth_weight_conv, float *s4, float *q4, float *cq4, mdf_stream_t stream);

MDF_API size_t mdf_cost_volume_prepped_workspace_bytes(int B, int N);

MDF_API int mdf_cost_volume_fwd_prepped(const float *s4, const float *q4, const float *cq4, int N, const float *ref_proj,
                                        const float *const *src_projs, const float *depth_hypos, int hypos_per_pixel,
                                        const float *conv_weight, const float *bn_weight, const float *bn_bias,
                                        const float *bn_mean, const float *bn_var, float bn_eps,
                                        const float *fc_weight, const float *fc_bias,
                                        int B, int G, int D, int H, int W, float *cost_volume,
                                        void *workspace, size_t workspace_bytes, mdf_stream_t stream);

/* ---- VectorAggregate under autograd / in train mode (C == 2*G, G in {8,16,32}) ---------------
 * Reference: net/unit/homoaggregate.py:25-46 driven by torch autograd (train.py:33-50).
 * training != 0: BatchNorm3d uses the batch statistics of each source view's z over (B,D,H,W) (the module is
 * applied once per view, homoaggregate.py:40); `batch_stats` (device, (N-1,2) floats, may be NULL) receives
 * (mean, unbiased variance) per view so that the caller can apply the momentum update of the running
 * statistics in view order.  training == 0: running statistics (the gradient of an eval-mode forward).
 * Gradients: `grad_features` is a HOST array of N device pointers (each (B,C,H,W), overwritten; an entry may
 * be NULL to skip that view); `grad_params` (device, 4+G floats) = d bn.weight, d bn.bias, d fc.weight,
 * d fc.bias, d conv.weight[G].  Projections and hypotheses get no gradient (base.py:97 is under no_grad). */
MDF_API size_t mdf_cost_volume_train_workspace_bytes(int B, int N, int C, int G, int D, int H, int W);

MDF_API int mdf_cost_volume_train_fwd(const float *const *features, int N, const float *ref_proj,
                                      const float *const *src_projs, const float *depth_hypos, int hypos_per_pixel,
                                      const float *conv_weight, const float *bn_weight, const float *bn_bias,
                                      const float *bn_mean, const float *bn_var, float bn_eps,
                                      const float *fc_weight, const float *fc_bias, int training,
                                      int B, int C, int G, int D, int H, int W, float *cost_volume,
                                      float *batch_stats, void *workspace, size_t workspace_bytes, mdf_stream_t stream);

MDF_API int mdf_cost_volume_bwd(const float *const *features, int N, const float *ref_proj,
                                const float *const *src_projs, const float *depth_hypos, int hypos_per_pixel,
                                const float *conv_weight, const float *bn_weight, const float *bn_bias,
                                const float *bn_mean, const float *bn_var, float bn_eps,
                                const float *fc_weight, const float *fc_bias, int training,
                                int B, int C, int G, int D, int H, int W,
                                const float *cost_volume /* saved forward output */, const float *grad_out,
                                const float *batch_stats /* what the training forward returned, or NULL: recomputed */,
                                float *const *grad_features, float *grad_params,
                                void *workspace, size_t workspace_bytes, mdf_stream_t stream);

/* The momentum update of BatchNorm3d's running statistics after a train-mode forward (base.py:50-68 inside
 * homoaggregate.py:40: one application per source view => V sequential updates in view order): `batch_stats` is what
 * mdf_cost_volume_train_fwd returned, `momentum` < 0 means momentum=None (cumulative moving average over
 * num_batches_tracked); running_mean / running_var (1 float each) and num_batches_tracked (1 int64, may be NULL) are
 * updated in place.  One launch. */
MDF_API int mdf_bn_running_update(const float *batch_stats, int V, float momentum, float *running_mean, float *running_var,
                                  long long *num_batches_tracked, mdf_stream_t stream);

/* ---- homo_aggregate_by_variance -> (B,C,D,H,W) ---------------------------------------------- */
MDF_API size_t mdf_variance_volume_workspace_bytes(int B, int N, int C, int D, int H, int W);

MDF_API int mdf_variance_volume_fwd(const float *const *features, int N, const float *ref_proj,
                            const float *const *src_projs, const float *depth_hypos, int hypos_per_pixel,
                            int B, int C, int D, int H, int W, float *cost_volume,
                            void *workspace, size_t workspace_bytes, mdf_stream_t stream);

/* ---- head ---------------------------------------------------------------------------------- */
/* Fused softmax over D of the regulariser logits + expectation; optionally also the photometric
 * confidence of the same probability volume.  prob and confidence may be NULL (not produced). */
MDF_API int mdf_softmax_regress_fwd(const float *logits,        /* (B,D,H,W) */
                            const float *depth_hypos, int hypos_per_pixel,
                            int B, int D, int H, int W,
                            float *prob,                /* (B,D,H,W) or NULL */
                            float *depth,               /* (B,H,W) */
                            float *confidence,          /* (B,H*up,W*up) or NULL */
                            int conf_n, int conf_pad_front, int conf_pad_back, int conf_upsample,
                            mdf_stream_t stream);

/* The same pass with HyposByFit's per-pixel curve fit (net/unit/depthhypos.py:78-125, :169-215; core.py:55 hands the
 * probability volume and the regressed depth of this stage to the next stage's Depth_hypos) done on the column while
 * it is on chip: `s` (B,H,W) receives the fitted scale that mdf_hypos_generate_fwd consumes, so the probability
 * volume is never re-read and need not be written (prob == NULL) unless the caller wants it.  curve: 0 = no fit
 * (s must be NULL; identical to mdf_softmax_regress_fwd), 1 = "gauss1", 2 = "laplace".  The fit uses the depth this
 * call regresses. */
MDF_API int mdf_softmax_regress_fit_fwd(const float *logits, const float *depth_hypos, int hypos_per_pixel,
                                int B, int D, int H, int W, float *prob, float *depth, float *confidence,
                                int conf_n, int conf_pad_front, int conf_pad_back, int conf_upsample,
                                int curve, float *s, mdf_stream_t stream);

/* The regulariser's last layer folded in as well: x (B,C,D,H,W) is the feature volume that
 * net/unit/regular.py:67 / :130 feeds to `self.prob = nn.Conv3d(c0, 1, 3, stride=1, padding=1, bias=False)`
 * (regular.py:43,110), prob_weight its (1,C,3,3,3) weight.  One launch: convolution -> softmax over D -> depth
 * expectation (-> confidence) (-> curve fit); the logits only exist in registers unless `logits` (B,D,H,W) is
 * given.  Any of logits / prob / depth / confidence may be NULL; s is required iff curve != 0.  D must be 8, 24 or 48
 * (config.py:199), C <= 64. */
MDF_API int mdf_prob_head_fwd(const float *x, const float *prob_weight, const float *depth_hypos, int hypos_per_pixel,
                              int B, int C, int D, int H, int W, float *logits, float *prob, float *depth,
                              float *confidence, int conf_n, int conf_pad_front, int conf_pad_back, int conf_upsample,
                              int curve, float *s, mdf_stream_t stream);
MDF_API int mdf_depth_regression_fwd(const float *prob, const float *depth_hypos, int hypos_per_pixel,
                             int B, int D, int H, int W, float *depth, mdf_stream_t stream);

/* confidence = (n * avg_pool_D(pad_D(prob, pad_front, pad_back), n))[trunc(sum_d prob_d * d)],
 * replicated `upsample` x `upsample` (1 = regress.py as is, 2 = with core.py:76-77). */
MDF_API int mdf_confidence_fwd(const float *prob, int B, int D, int H, int W,
                       int n, int pad_front, int pad_back, int upsample,
                       float *confidence, mdf_stream_t stream);

/* ---- next-stage depth hypotheses: HyposByFit (net/unit/depthhypos.py:27-76) -----------------
 * curve: 1 = "gauss1" (:169-215), 2 = "laplace" (:78-125), 3 = "gauss0" (:127-167; these two entries only -- the fused
 * tails above take the curves config.py wires, 1 and 2).  mdf_hypos_fit_fwd gives the fitted scale s (B,H,W) of
 * every pixel's probability column; mdf_hypos_generate_fwd upsamples s and depth x2 (bilinear, align_corners
 * = False) when `upsample`, derives the search range from prob_thresh, applies the reference's clamps and writes
 * `ndepths` hypotheses (B,ndepths,2H|H,2W|W).  depth_range: (B,2) device floats (min, max).
 * Stage 0's uniform hypotheses (:31-38) are B*ndepths numbers and stay in torch. */
MDF_API int mdf_hypos_fit_fwd(const float *prob, const float *depth_hypos, int hypos_per_pixel, const float *depth,
                              int curve, int B, int D, int H, int W, float *s, mdf_stream_t stream);
MDF_API int mdf_hypos_generate_fwd(const float *depth, const float *s, const float *depth_range, int curve,
                                   float prob_thresh, int upsample, int B, int H, int W, int ndepths,
                                   float *depth_hypos, mdf_stream_t stream);

/* ---- geometric-consistency filter of the post-processing --------------------
 * Reference: tools/filter/dynamic_filter_gpu.py -- reproject_with_depth (:184-237), check_geometric_consistency
 * (:161-182) and the per-view aggregation of filter() (:57-100), one launch per reference view instead of ~60 ATen
 * kernels + 3 cuSOLVER / cuBLAS calls per (reference, source) pair.  All maps are (H,W) float32 at full resolution;
 * intrinsics 3x3, extrinsics 4x4 (device memory, row major; src_* hold S of them back to back); src_depths is a HOST
 * array of S device pointers.  Outputs (each may be NULL): src_bits (S,H,W) uint16 -- bit i-2 = the dynamic mask of
 * threshold i = 2..10 (dist < i/thre1 px and relative depth difference < i/thre2), i.e. the `masks` list of
 * check_geometric_consistency; depth_reprojected (S,H,W), zero where the loosest mask fails; depth_averaged (H,W);
 * geo_mask / photo_mask / final_mask (H,W) uint8 as filter() computes them (photo_mask = confidence > photo_threshold,
 * all ones when confidence is NULL).  thre1 and thre2 must be positive and finite (MDF_ERR_UNSUPPORTED otherwise). */
#define MDF_MAX_FILTER_VIEWS 32
MDF_API size_t mdf_geo_filter_workspace_bytes(int S);
MDF_API int mdf_geo_filter_fwd(const float *ref_depth, const float *ref_intrinsics, const float *ref_extrinsics,
                               const float *const *src_depths, const float *src_intrinsics, const float *src_extrinsics,
                               int S, int H, int W, const float *confidence, float photo_threshold, int nconditions,
                               float thre1, float thre2, uint16_t *src_bits, float *depth_reprojected,
                               float *depth_averaged, uint8_t *geo_mask, uint8_t *photo_mask, uint8_t *final_mask,
                               void *workspace, size_t workspace_bytes, mdf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MDF_B200_H_ */

/* mdf_b200_debug.h -- test / benchmark / tuning entry points of libmdf_b200.so.
 *
 * NOT part of the drop-in boundary (include/mdf_b200.h): nothing a reference maintainer binds lives here.  These
 * exist so that the parity tests can pin internals (the coordinate chain), bench.py can time the dominant kernel
 * alone, and tools/ can select tuning variants.  Like the product entry points they allocate nothing and keep no
 * state between calls. */
#ifndef MDF_B200_DEBUG_H_
#define MDF_B200_DEBUG_H_

#include "mdf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* mdf_cost_volume_fwd with algorithm selection and an optional timing hook.
 * algo: 0 = auto, 1 = staged (TMA box) kernel, 2 = direct kernel (any C/G, taps straight from the NCHW features).
 * Tuning builds (-DMDF_TUNING, `python -m mdf_net_b200.build --tuning`) add 16 + k = staged kernel, variant k, and
 * 32 + k (+ 256 * rounds per item) = the experimental pre-planned pipeline kernel; the product build answers
 * MDF_ERR_UNSUPPORTED to those.
 * hot_start_event / hot_stop_event: two cudaEvent_t (or NULL, NULL) recorded on `stream` around the hot kernel
 * (cost_volume_staged_kernel) alone -- not around the layout pass.  bench.py's live roofline numbers use them. */
MDF_API int mdf_cost_volume_fwd_ex(const float *const *features, int N, const float *ref_proj,
                           const float *const *src_projs, const float *depth_hypos, int hypos_per_pixel,
                           const float *conv_weight, const float *bn_weight, const float *bn_bias,
                           const float *bn_mean, const float *bn_var, float bn_eps,
                           const float *fc_weight, const float *fc_bias,
                           int B, int C, int G, int D, int H, int W, float *cost_volume,
                           void *workspace, size_t workspace_bytes, int algo,
                           void *hot_start_event, void *hot_stop_event, mdf_stream_t stream);

/* mdf_prob_head_fwd with algo: 0 = default, 1.. = alternative tile / depth-slab shapes of the same kernel
 * (tools/time_prob_head.py). */
MDF_API int mdf_prob_head_fwd_ex(const float *x, const float *prob_weight, const float *depth_hypos, int hypos_per_pixel,
                                 int B, int C, int D, int H, int W, float *logits, float *prob, float *depth,
                                 float *confidence, int conf_n, int conf_pad_front, int conf_pad_back, int conf_upsample,
                                 int curve, float *s, int algo, mdf_stream_t stream);

/* Sample positions (pixel units of the source map, as grid_sample uses them: base.py:102-119 +
 * ATen unnormalize) of every (d, y, x) for one precomposed projection `rot_trans` (12 floats:
 * rot row-major, then trans), computed with the hot kernel's division-free coordinate chain.
 * Exists so that the parity tests can pin that chain bit for bit; no product path calls it. */
MDF_API int mdf_debug_sample_positions(const float *rot_trans, const float *depth_hypos, int hypos_per_pixel,
                                       int D, int H, int W, float *ix /* (D,H,W) */, float *iy /* (D,H,W) */,
                                       mdf_stream_t stream);

#ifdef MDF_TUNING
/* Phase timestamps (SM clock) written by the tracing variant of the hot kernel (algo 31): copies up to max_slots
 * records of 32 words to the host, resets the device-side buffer, returns the number of records. */
MDF_API int mdf_debug_read_trace(long long *host_dst, int max_slots);
#endif

#ifdef __cplusplus
}
#endif
#endif /* MDF_B200_DEBUG_H_ */

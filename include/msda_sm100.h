/*
 * msda_sm100.h -- C ABI of libmsda_sm100.so: multi-scale deformable attention (MSDeformAttn)
 * forward and backward for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's native component models/ops/src/**.  Each entry
 * point replaces one reference launch wrapper and keeps its argument list (plain device pointers and
 * sizes, the caller's stream); no torch/ATen type appears here:
 *
 *   msda_forward_*   <->  ms_deformable_im2col_cuda(stream, value, shapes, start, loc, attn,
 *                           batch, S, M, D, L, Lq, P, out)          ms_deform_im2col_cuda.cuh:923-937
 *                         as driven by ms_deform_attn_cuda_forward  ms_deform_attn_cuda.cu:20-80
 *   msda_backward_*  <->  ms_deformable_col2im_cuda(stream, grad_col, value, shapes, start, loc,
 *                           attn, batch, S, M, D, L, Lq, P, grad_value, grad_loc, grad_attn)
 *                                                                   ms_deform_im2col_cuda.cuh:956-973
 *                         as driven by ms_deform_attn_cuda_backward ms_deform_attn_cuda.cu:83-153
 *
 * The Python-visible functions the reference exports through pybind11 (vision.cpp:13-16,
 * ms_deform_attn.h:36-77) -- ms_deform_attn_forward / ms_deform_attn_backward -- are rebuilt on top
 * of this ABI with ctypes in ocpg_b200/MultiScaleDeformableAttention.py; INTEGRATION.md shows the
 * binding.
 *
 * Layouts (row-major, contiguous, device memory):
 *   value            [batch][spatial_size][num_heads][channels]
 *   spatial_shapes   [num_levels][2]   int64, (H_l, W_l)        -- read on the device, like the reference
 *   level_start_index[num_levels]      int64
 *   sampling_loc     [batch][num_query][num_heads][num_levels][num_point][2]   (x, y) in [0,1]
 *   attn_weight      [batch][num_query][num_heads][num_levels][num_point]
 *   output, grad_output [batch][num_query][num_heads*channels]
 *
 * Differences from the reference wrappers (all strict supersets):
 *   - one launch for any batch: there is no im2col_step chunk loop, hence no batch % im2col_step rule;
 *   - the batch term of every offset is 64-bit (the reference's int overflows beyond 2^31 elements);
 *   - msda_backward_* zero-fills grad_value itself (the reference's caller does, cu:121) and writes
 *     every element of grad_sampling_loc / grad_attn_weight (no zero-fill needed);
 *   - launch errors are returned, not printf'd (cuh:948-952).
 *
 * Every function is stateless, re-entrant and stream-ordered: it enqueues work on `stream` and
 * returns without synchronising; it allocates nothing and keeps no device state.
 * Return value: 0 on success, a cudaError_t value (> 0) if a CUDA call failed, or a negative
 * MSDA_ERR_* code for an argument error; msda_last_error() gives the message for this thread.
 */
#ifndef MSDA_SM100_H_
#define MSDA_SM100_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_ABI_VERSION 4

#define MSDA_OK 0
#define MSDA_ERR_INVALID_ARGUMENT (-1) /* null pointer, non-positive dimension, misaligned pointer */
#define MSDA_ERR_UNSUPPORTED (-2)      /* shape outside what the kernels cover (see msda_kernel_plan) */

/* Opaque cudaStream_t (CUstream); 0 / NULL is the legacy default stream. */
typedef void *msda_stream_t;

int msda_abi_version(void);
const char *msda_last_error(void);

/* ---- fp32 (the reference's production dtype; deformable_transformer.py:250 disables autocast) ---- */
int msda_forward_f32(const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                     const float *sampling_loc, const float *attn_weight,
                     int batch, int spatial_size, int num_heads, int channels, int num_levels,
                     int num_query, int num_point, float *output, msda_stream_t stream);

int msda_backward_f32(const float *grad_output, const float *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index, const float *sampling_loc, const float *attn_weight,
                      int batch, int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point,
                      float *grad_value, float *grad_sampling_loc, float *grad_attn_weight,
                      msda_stream_t stream);

/* ---- fp64 (the reference dispatches float and double, cu:64; its tests run in double) ---- */
int msda_forward_f64(const double *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                     const double *sampling_loc, const double *attn_weight,
                     int batch, int spatial_size, int num_heads, int channels, int num_levels,
                     int num_query, int num_point, double *output, msda_stream_t stream);

int msda_backward_f64(const double *grad_output, const double *value, const int64_t *spatial_shapes,
                      const int64_t *level_start_index, const double *sampling_loc, const double *attn_weight,
                      int batch, int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point,
                      double *grad_value, double *grad_sampling_loc, double *grad_attn_weight,
                      msda_stream_t stream);

/* ---- bf16 value (new; value / output / grad_output are bf16 bit patterns, loc and attn stay fp32,
 *      accumulation is fp32).  grad_value_f32 == NULL with grad_value_bf16 != NULL selects direct accumulation in bf16
 *      (packed bf16 reds, row-major kernel, num_levels <= 4: ~4-6 % of max error); otherwise
 *      grad_value is accumulated in the fp32 buffer grad_value_f32
 *      (same shape as value, zero-filled by the call); if grad_value_bf16 is non-NULL the rounded
 *      result is also written there. ---- */
int msda_forward_bf16(const uint16_t *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                      const float *sampling_loc, const float *attn_weight,
                      int batch, int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point, uint16_t *output, msda_stream_t stream);

int msda_backward_bf16(const uint16_t *grad_output, const uint16_t *value, const int64_t *spatial_shapes,
                       const int64_t *level_start_index, const float *sampling_loc, const float *attn_weight,
                       int batch, int spatial_size, int num_heads, int channels, int num_levels,
                       int num_query, int num_point,
                       float *grad_value_f32, uint16_t *grad_value_bf16,
                       float *grad_sampling_loc, float *grad_attn_weight, msda_stream_t stream);

/* ---- fused module path (new).  Folds the elementwise work that
 *      MSDeformAttn.forward does around the operator into the kernels (models/ops/modules/ms_deform_attn.py):
 *        attention_weights  = softmax(logits) over the num_levels*num_point entries of a (query, head)   :101-102
 *        sampling_locations = ref + offsets / (W_l, H_l)                     (ref_dim == 2)             :104-107
 *                           = ref_xy + offsets / num_point * ref_wh * 0.5    (ref_dim == 4)             :108-110
 *      offsets   [batch][num_query][num_heads][num_levels][num_point][2]   raw output of the sampling_offsets Linear
 *      logits    [batch][num_query][num_heads][num_levels*num_point]       raw output of the attention_weights Linear
 *      ref       [batch][num_query][num_levels][ref_dim]                   reference points (2) or boxes (4)
 *      Forward: sampling_loc_out / attn_weight_out are optional (NULL = not materialised; the encoder discards
 *      them, deformable_transformer.py:251; the decoder reads them, :365-375).
 *      Backward: writes grad_value (zero-filled by the call), grad_offsets, grad_logits (softmax gradient
 *      applied) and, if grad_sampling_loc_out is non-NULL, d/d sampling_locations, which the caller reduces over
 *      heads and points into grad_reference_points when those need a gradient.
 *      Shapes as msda_kernel_plan() == 1 only (channels == 32 ...): MSDA_ERR_UNSUPPORTED otherwise -- the caller
 *      then uses the unfused operator above. ---- */
int msda_fused_forward_f32(const float *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                           const float *offsets, const float *logits, const float *ref, int ref_dim,
                           int batch, int spatial_size, int num_heads, int channels, int num_levels,
                           int num_query, int num_point, float *output,
                           float *sampling_loc_out, float *attn_weight_out, msda_stream_t stream);

int msda_fused_backward_f32(const float *grad_output, const float *value, const int64_t *spatial_shapes,
                            const int64_t *level_start_index, const float *offsets, const float *logits,
                            const float *ref, int ref_dim,
                            int batch, int spatial_size, int num_heads, int channels, int num_levels,
                            int num_query, int num_point,
                            float *grad_value, float *grad_offsets, float *grad_logits,
                            float *grad_sampling_loc_out, msda_stream_t stream);

int msda_fused_forward_bf16(const uint16_t *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                            const float *offsets, const float *logits, const float *ref, int ref_dim,
                            int batch, int spatial_size, int num_heads, int channels, int num_levels,
                            int num_query, int num_point, uint16_t *output,
                            float *sampling_loc_out, float *attn_weight_out, msda_stream_t stream);

int msda_fused_backward_bf16(const uint16_t *grad_output, const uint16_t *value, const int64_t *spatial_shapes,
                             const int64_t *level_start_index, const float *offsets, const float *logits,
                             const float *ref, int ref_dim,
                             int batch, int spatial_size, int num_heads, int channels, int num_levels,
                             int num_query, int num_point,
                             float *grad_value_f32, uint16_t *grad_value_bf16, float *grad_offsets, float *grad_logits,
                             float *grad_sampling_loc_out, msda_stream_t stream);

/* ---- encoder layer epilogue (new).  The elementwise / reduction work of
 *      DeformableTransformerEncoderLayer around the attention and the FFN (models/deformable_transformer.py:243-260)
 *      as HBM-streaming kernels:
 *        forward   z = (x + bias) + residual ;  y = LayerNorm(z) * gamma + beta        (:253-254, :246-247)
 *                  x, residual, z, y: [rows][channels]; bias (may be NULL), gamma, beta: [channels];
 *                  mean, rstd: [rows] (saved for the backward); channels in {128, 256, 512, 1024}
 *        backward  dz = d x = d residual, grad_gamma, grad_beta, grad_bias (NULL = not wanted) -- the three
 *                  [channels] outputs are zero-filled by the call and accumulated with fp32 reds
 *        msda_column_sum_f32                out[c] = sum_r x[r][c]: the bias gradient of a Linear (channels % 4 == 0)
 *        msda_relu_backward_column_sum_f32  dpre = dh * (h > 0) and dbias = column sum of dpre, h = the ReLU output ---- */
int msda_epilogue_ln_forward_f32(const float *x, const float *bias, const float *residual, const float *gamma,
                                 const float *beta, float eps, int64_t rows, int channels,
                                 float *z, float *y, float *mean, float *rstd, msda_stream_t stream);

int msda_epilogue_ln_backward_f32(const float *dy, const float *z, const float *mean, const float *rstd,
                                  const float *gamma, int64_t rows, int channels,
                                  float *dz, float *grad_gamma, float *grad_beta, float *grad_bias, msda_stream_t stream);

int msda_column_sum_f32(const float *x, int64_t rows, int channels, float *out, msda_stream_t stream);

int msda_relu_backward_column_sum_f32(const float *dh, const float *h, int64_t rows, int channels,
                                      float *dpre, float *dbias, msda_stream_t stream);

/* ---- the same epilogue with dropout ACTIVE (training: dropout1 / dropout2 / dropout3 of
 *      DeformableTransformerEncoderLayer, models/deformable_transformer.py:226-235, p = 0.1 in opts.py).
 *      The keep mask is never stored: it is Philox4x32-7 of (rng[0], rng[1], salt, index of the element's 4-float
 *      chunk), `rng` = two 64-bit words in DEVICE memory (read by the kernel: safe to capture in a CUDA graph and
 *      refresh between replays), `salt` = the call site.  keep <=> 32-bit word >= floor(p * 2^32), 0 <= p < 1;
 *      kept elements are scaled by 1 / (1 - p) like torch.nn.functional.dropout.
 *        msda_epilogue_ln_dropout_forward_f32    z = dropout(x + bias) + residual ; y = LayerNorm(z)
 *        msda_epilogue_ln_dropout_backward_f32   dz = d residual, dx = dz * keep / (1 - p) = d x, grad_bias = column sum of dx
 *        msda_dropout_inplace_f32                h *= keep / (1 - p) in place, n = element count (multiple of 4)
 *        msda_relu_dropout_backward_column_sum_f32   for h_dropped = dropout(relu(pre)): dpre = dh * (h_dropped > 0) / (1 - p)
 *                                                (h_dropped > 0 <=> pre > 0 and kept: no mask needed) and its column sum
 *        msda_dropout_mask_u8                    the keep mask itself, one byte per element (tests only) ---- */
int msda_epilogue_ln_dropout_forward_f32(const float *x, const float *bias, const float *residual, const float *gamma,
                                         const float *beta, float eps, int64_t rows, int channels,
                                         const void *rng, uint32_t salt, float p,
                                         float *z, float *y, float *mean, float *rstd, msda_stream_t stream);

int msda_epilogue_ln_dropout_backward_f32(const float *dy, const float *z, const float *mean, const float *rstd,
                                          const float *gamma, int64_t rows, int channels,
                                          const void *rng, uint32_t salt, float p,
                                          float *dz, float *dx, float *grad_gamma, float *grad_beta, float *grad_bias,
                                          msda_stream_t stream);

int msda_dropout_inplace_f32(float *h, int64_t n, const void *rng, uint32_t salt, float p, msda_stream_t stream);

int msda_relu_dropout_backward_column_sum_f32(const float *dh, const float *h_dropped, float p, int64_t rows, int channels,
                                              float *dpre, float *dbias, msda_stream_t stream);

int msda_dropout_mask_u8(const void *rng, uint32_t salt, float p, int64_t n, uint8_t *keep, msda_stream_t stream);

/* ---- decoder-side consumers of the cross-attention's outputs (new; SURVEY.md section 8f rank 3;
 *      DeformableTransformerDecoder.forward, models/deformable_transformer.py:353-375).
 *        msda_decoder_select_samples_f32   per (frame, query): the `top` largest of the num_heads*num_levels*num_point
 *            attention weights in descending order (ties: ascending index), and the sampling locations of those points
 *            divided by their level's valid ratio -- `samples_keep` of :368-375 in one launch.
 *            sampling_loc [batch][num_query][heads][levels][points][2], attn_weight [batch][num_query][heads][levels][points],
 *            valid_ratios [batch][levels][2] (w, h); out: samples_keep [batch][num_query][top][2], top_weights
 *            [batch][num_query][top] (may be NULL), top_idx int64 [batch][num_query][top] (may be NULL).
 *            Needs top <= 32 and top <= heads*levels*points <= 256 (the reference: 30 of 128).
 *        msda_decoder_reference_points_f32 reference_points_input [batch][num_query][levels][ref_dim] =
 *            reference_points [batch][num_query][ref_dim] * valid_ratios (repeated twice for ref_dim == 4)   (:358-363) ---- */
int msda_decoder_select_samples_f32(const float *sampling_loc, const float *attn_weight, const float *valid_ratios,
                                    int batch, int num_query, int num_heads, int num_levels, int num_point, int top,
                                    float *samples_keep, float *top_weights, int64_t *top_idx, msda_stream_t stream);

int msda_decoder_reference_points_f32(const float *reference_points, const float *valid_ratios, int batch, int num_query,
                                      int num_levels, int ref_dim, float *reference_points_input, msda_stream_t stream);

/* ---- the layout traffic either side of the encoder (new; SURVEY.md section 8f rank 4;
 *      DeformableTransformer.forward, models/deformable_transformer.py:149-169 and :205-212).
 *        msda_flatten_levels_f32    src_flatten [batch][S][channels] = cat_l transpose(src_l [batch][channels][H_l*W_l]) and,
 *            when pos_levels != NULL, pos_flatten = the same of pos_l + level_embed[l] (level_embed [levels][channels], may be
 *            NULL); S = sum_l H_l*W_l.  `src_levels` / `pos_levels` / `heights` / `widths` are HOST arrays of num_levels
 *            entries (device pointers, ints) -- the reference takes (h, w) from the maps' shapes on the host too (:151-153).
 *        msda_unflatten_levels_f32  maps[l] [batch][channels][H_l*W_l] = rows [start_l, start_l + H_l*W_l) of
 *            flat [batch][spatial_size][channels], for the FIRST num_levels levels (the reference converts all but the
 *            last, :207); spatial_size >= sum of those levels' pixels (0 = exactly that sum).
 *      Each is the other's backward.  num_levels <= 8; any channel count. ---- */
int msda_flatten_levels_f32(int num_levels, const float *const *src_levels, const float *const *pos_levels,
                            const float *level_embed, const int *heights, const int *widths, int batch, int channels,
                            float *src_flatten, float *pos_flatten, msda_stream_t stream);

int msda_unflatten_levels_f32(int num_levels, const float *flat, const int *heights, const int *widths, int batch,
                              int channels, int spatial_size, float *const *maps, msda_stream_t stream);

/* Which kernel a call with these dimensions runs: 1 = the sm_100a tiled kernel (channels == 32,
 * num_levels <= 16, num_levels*num_point <= 32), 0 = the generic kernel (any shape).  For tests and
 * benchmarks; `elem_bytes` is 2, 4 or 8. */
int msda_kernel_plan(int elem_bytes, int num_heads, int channels, int num_levels, int num_point);

/* Number of kernel launches (memsets excluded) the library has enqueued from this process so far;
 * used by bench.py to report `gpu_launches`. */
uint64_t msda_launch_count(void);

/* Tuning knobs (not needed for normal use; every setting returns correct results).  Known keys:
 *   "fwd_warps", "bwd_warps"             8 or 16 warps per CTA of the query-major kernels (0 = default)
 *   "fwd_ctas_per_sm", "bwd_ctas_per_sm" resident CTAs per SM the persistent grid is sized for (0 = default)
 *   "frame_chunk"                        frames whose passes are interleaved by the task walk (0 = automatic)
 *   "force_generic"                      use the any-shape kernels even when the tiled ones apply
 *   "force_linear_walk"                  walk queries linearly instead of as spatial tiles
 *   "bwd_algo"                           0 = automatic: the row-major backward msda_bwd_sorted for encoder shapes
 *                                        (num_query == spatial_size, num_levels <= 4), the query-major msda_bwd_tiled
 *                                        otherwise; 1 = always query-major; 2 = row-major for any filled launch
 *   "bwd_deep"                           under-filled-launch backward variant (a round's 32 row loads issued before their
 *                                        first use): 0 = automatic (fewer passes than 2 x SMs), 1 = always, -1 = never
 * The measurement-only switches "bwd_mode" / "debug_skip_scatter" (they omit the grad_value scatter) exist only in
 * -DMSDA_EXPERIMENTS builds of the library; the product build answers MSDA_ERR_UNSUPPORTED.
 * Returns 0, or MSDA_ERR_INVALID_ARGUMENT for an unknown key. */
int msda_set_option(const char *key, int value);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_SM100_H_ */

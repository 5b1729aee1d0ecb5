/*
 * contextgs_b200 -- C ABI of the B200-native (sm_100a) ContextGS hot path.
 *
 * Boundary contract:
 *   - plain C, raw DEVICE pointers + sizes + a CUDA stream handle (void* == cudaStream_t);
 *   - no allocation, no host synchronisation and no exceptions inside any entry point:
 *     the caller owns every buffer (the Python shims allocate them with torch) and every call
 *     only enqueues kernels/memsets on `stream`;
 *   - return value 0 = success, negative = error (see cgs_last_error());
 *   - re-entrant per stream; the library keeps no global state besides the last-error string.
 *
 * Each entry point names the reference interface it replaces.  The rasterizer itself is not in
 * the reference tree (git-ignored submodule); its interface is taken from the reference's call
 * sites in gaussian_renderer/__init__.py.
 */
#ifndef CONTEXTGS_B200_H
#define CONTEXTGS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CGS_API __attribute__((visibility("default")))
#else
#define CGS_API
#endif

#define CGS_ABI_VERSION 1
#define CGS_TILE 16            /* tile edge in pixels */
#define CGS_GEOM_STRIDE 12     /* floats per packed per-Gaussian record (48 B) */
#define CGS_SUPER_X 8          /* binning super-tile: 8 x 4 tiles (one 32-bit footprint mask per Gaussian) */
#define CGS_SUPER_Y 4

/* Field-for-field mirror of `GaussianRasterizationSettings`
 * (reference: gaussian_renderer/__init__.py:179-192 and :250-263).
 * viewmatrix / projmatrix are passed exactly as the reference passes them:
 * `world_view_transform` / `full_proj_transform`, i.e. the transposed 4x4 stored row-major. */
typedef struct cgs_raster_settings {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    float bg[3];
    float scale_modifier;
    float viewmatrix[16];
    float projmatrix[16];
    int32_t sh_degree;   /* accepted, unused: ContextGS always passes colors_precomp (line 200-201) */
    float campos[3];
    int32_t prefiltered;
    int32_t debug;
} cgs_raster_settings;

/* Indices into the device-side int32 status block written by cgs_rasterize_forward. */
enum {
    CGS_STATUS_NUM_RENDERED = 0, /* R = number of (Gaussian, tile) instances (low 31 bits) */
    CGS_STATUS_OVERFLOW = 1,     /* 1 if R > R_cap: the image is incomplete, re-run with a larger R_cap */
    CGS_STATUS_NUM_SORTED = 2,   /* min(R, R_cap) */
    CGS_STATUS_NUM_GAUSSIANS = 3,      /* cgs_rasterize_forward_dev: *P_dev as read on the device */
    CGS_STATUS_GAUSSIAN_OVERFLOW = 4,  /* 1 if *P_dev exceeded the capacity P the buffers were sized for */
    CGS_STATUS_NUM_PAIRS = 5,          /* (super-tile, Gaussian) pairs the binning sorted (<= num_rendered) */
    CGS_STATUS_WORDS = 8
};

CGS_API int cgs_abi_version(void);
CGS_API const char *cgs_last_error(void);

/* Measurement hooks (no reference counterpart; the reference only brackets whole iterations with
 * torch.cuda.Event, train.py:116-117,142,213).  Launch accounting is always on; per-stage CUDA
 * events are recorded on the caller's stream only between cgs_stage_timing_enable(1) and (0).
 * cgs_stage_timing_read synchronises on the recorded events and returns, per stage, the summed
 * milliseconds and the number of timed scopes since enable.  cgs_launch_counts returns the number
 * of kernels launched per stage (reset != 0 clears the counters). */
CGS_API int cgs_stage_count(void);
CGS_API const char *cgs_stage_name(int stage);
CGS_API int cgs_stage_timing_enable(int on);
CGS_API int cgs_stage_timing_read(double *ms_sum, int64_t *scopes);
CGS_API int cgs_launch_counts(int64_t *launches, int reset);

/* Diagnostic: one 128-row tile GEMM D[128,N] = A[128,K] * W[N,K]^T on the tcgen05 tensor cores with
 * the building blocks the fused MLP kernels use (A in TMEM, W in shared memory, 3xTF32 when mode = 1,
 * plain TF32 when mode = 0).  *err (device) is set to 1 if the completion barrier timed out. */
CGS_API int cgs_umma_selftest(const float *A, const float *W, int N, int K, int mode, float *D, int32_t *err,
                              void *stream);
/* Probe of the SS form (both operands in shared memory) with the ROW index as the contraction: D[M,N] = P^T Q,
 * P[128,M], Q[128,N], 3xTF32 -- the shape of the backward kernels' weight-gradient GEMMs.
 * skew: extra 16-byte units in the leading-dimension byte offset.  scripts/umma_ss_probe.py runs it. */
CGS_API int cgs_umma_selftest_ss(const float *P, const float *Q, int M, int N, int skew, float *D, int32_t *err,
                                 void *stream);
/* The same contraction with both operands in the MN-major no-swizzle layout ([feature / 4][row][4 floats]: one
 * float4 store per four features of a row).  variant 0: SBO = stride between 4-feature groups, LBO = stride
 * between 8-row groups; variant 1: swapped.  N <= 48. */
/* Micro-benchmark (diagnostic): cycles for a chain of `iters` tcgen05.mma.kind::tf32 (M = 128, K = 8, given N) issued by
 * one thread; form 0 = A in TMEM, 1 = A in shared memory; n_acc independent accumulators round-robin.
 * out_cycles[0] = issue time, [1] = issue + completion (SM clock cycles).  scripts/umma_rate_probe.py. */
CGS_API int cgs_umma_mma_rate(int form, int N, int n_acc, int iters, long long *out_cycles, void *stream);
CGS_API int cgs_umma_selftest_ss_mn(const float *P, const float *Q, int M, int N, int variant, float *D, int32_t *err,
                                    void *stream);

/* ------------------------------------------------------------------ rasterizer (P1, R0-R7) */

/* Replaces `GaussianRasterizer.visible_filter(means3D, scales, rotations)`
 * (reference call: gaussian_renderer/__init__.py:280-285; upstream kernel filter_preprocessCUDA).
 * radii[N] int32: screen radius, 0 when culled. */
CGS_API int cgs_visible_filter(const cgs_raster_settings *s, int N, const float *means3D, const float *scales,
                       const float *rotations, int32_t *radii, void *stream);

/* `prefilter_voxel` in one pass (gaussian_renderer/__init__.py:232-287): the radius test of
 * `visible_filter` over all anchors with scales = get_scaling[:, :3] (scales has a row stride of
 * scale_stride floats, 6 for the reference's [N,6] tensor) and the normalised rotation row
 * rotation_row[4] applied to every anchor (the reference passes rotations[[0], :].repeat(N, 1)),
 * fused with the ordered compaction of the visible anchors:
 *   visible[N] uint8 (the bool mask the reference returns), vis_idx[N] int32 (ascending indices of the
 *   visible anchors), *count_dev = their number.  workspace: cgs_prefilter_workspace_bytes(N). */
CGS_API size_t cgs_prefilter_workspace_bytes(int N);
CGS_API int cgs_prefilter_anchors(const cgs_raster_settings *s, int N, const float *anchor, const float *scales,
                                  int scale_stride, const float *rotation_row, uint8_t *visible, int32_t *vis_idx,
                                  int32_t *count_dev, void *workspace, size_t workspace_bytes, void *stream);

/* Replaces `GaussianRasterizer.markVisible(positions)` (upstream checkFrustum). visible[N] uint8. */
CGS_API int cgs_mark_visible(const cgs_raster_settings *s, int N, const float *means3D, uint8_t *visible, void *stream);

/* Scratch bytes needed by cgs_rasterize_forward for P Gaussians and an instance capacity R_cap. */
CGS_API size_t cgs_raster_workspace_bytes(int P, int64_t R_cap, int W, int H);

/* Replaces `_C.rasterize_gaussians` as driven by `GaussianRasterizer.forward`
 * (reference call: gaussian_renderer/__init__.py:197-205) with shs=None, cov3D_precomp=None.
 *   in : means3D[P,3] colors[P,3] opacities[P] scales[P,3] rotations[P,4]     (fp32, contiguous)
 *   out: out_color[3,H,W], radii[P] int32
 *   kept for backward (caller-allocated):
 *        geom[P,12]   packed record {x, y, conic_a, conic_b, conic_c, opacity, r, g, b, depth,
 *                     radius(int bits), tiles_touched(uint bits)}
 *        point_list[R_cap] uint32  Gaussian ids sorted by (tile, depth, id)
 *        ranges[tiles,2]   uint32  [first, last+1) per tile
 *        final_T[H,W], n_contrib[H,W] uint32
 *        status[CGS_STATUS_WORDS] int32 (device)
 * Pipeline: preprocess -> 4-pass radix sort of Gaussians by depth -> look-back scan in depth order
 * emitting one (super-tile, id) pair per covered 8x4-tile super-tile -> radix sort of the pairs by
 * super-tile -> ballot-ranked expansion into per-tile lists + ranges -> blend (csrc/raster_binning.cu).
 * The resulting point_list/ranges are identical to the classical 64-bit (tile<<32|depth) sort.
 * R_cap must be below 2^30. */
CGS_API int cgs_rasterize_forward(const cgs_raster_settings *s, int P, const float *means3D, const float *colors,
                          const float *opacities, const float *scales, const float *rotations, int64_t R_cap,
                          float *out_color, int32_t *radii, float *geom, uint32_t *point_list, uint32_t *ranges,
                          float *final_T, uint32_t *n_contrib, int32_t *status, void *workspace,
                          size_t workspace_bytes, void *stream);

/* The same forward with the Gaussian count on the DEVICE: P is only the capacity of the per-Gaussian
 * buffers, the kernels read the actual count from *P_dev (e.g. the count_dev written by
 * cgs_neural_gaussians_*_forward), so the host never has to read it back between the two stages
 * (the reference synchronises twice there: boolean indexing at gaussian_renderer/__init__.py:119,136).
 * status[CGS_STATUS_NUM_GAUSSIANS / _GAUSSIAN_OVERFLOW] report the count and a capacity overflow.
 * P_dev == NULL is cgs_rasterize_forward. */
CGS_API int cgs_rasterize_forward_dev(const cgs_raster_settings *s, int P, const int32_t *P_dev, const float *means3D,
                                      const float *colors, const float *opacities, const float *scales,
                                      const float *rotations, int64_t R_cap, float *out_color, int32_t *radii,
                                      float *geom, uint32_t *point_list, uint32_t *ranges, float *final_T,
                                      uint32_t *n_contrib, int32_t *status, void *workspace, size_t workspace_bytes,
                                      void *stream);

/* Scratch bytes needed by cgs_rasterize_backward. */
CGS_API size_t cgs_raster_backward_workspace_bytes(int P);

/* Replaces `_C.rasterize_gaussians_backward` (autograd backward of the call above).
 * dL_dmeans2D[P,3] follows the upstream convention read by training_statis
 * (scene/gaussian_model.py:710): x,y = dL/d(pixel mean) * 0.5*{W,H}, z = 0.
 * All outputs are overwritten (not accumulated). */
CGS_API int cgs_rasterize_backward(const cgs_raster_settings *s, int P, const float *means3D, const float *scales,
                           const float *rotations, const int32_t *radii, const float *geom,
                           const uint32_t *point_list, const uint32_t *ranges, const float *final_T,
                           const uint32_t *n_contrib, const float *dL_dpix, float *dL_dmeans3D,
                           float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity, float *dL_dscales,
                           float *dL_drots, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ anchor -> Gaussians (G1) */

/* Number of floats of the packed decoder-MLP weight block consumed by cgs_neural_gaussians_forward:
 *   W1[54][152] (k-major; columns 0-49 opacity / 50-99 color / 100-149 cov hidden units, 2 zero pads),
 *   b1[152], W2_opacity[50][12] b[12], W2_color[50][32] b[32], W2_cov[50][72] b[72] (h-major, zero padded).
 * Source layers: scene/gaussian_model.py:153-174. */
CGS_API int cgs_neural_gaussians_packed_floats(void);
CGS_API size_t cgs_neural_gaussians_workspace_bytes(int Nv);

/* Replaces gaussian_renderer/__init__.py:106-145 (everything `generate_neural_gaussians` does after
 * the per-anchor attributes have been chosen): view dir/dist, 3 decoder MLPs, `opacity*mask > 0`
 * selection, order-preserving compaction, scale/rotation/position post-processing.
 *   vis_idx[Nv] : indices of the visible anchors (NULL = all Nv anchors in order)
 *   anchor[N,3] feat[N,50] offsets[N,10,3] scaling[N,6] mask[N,10]   (fp32, N = full anchor count)
 *   campos_host : 3 HOST floats (camera centre)
 *   outputs have CAPACITY Nv*10 rows; *count_dev (device int32) receives P, the number emitted:
 *     o_xyz[P,3] o_color[P,3] o_opacity[P] o_scaling[P,3] o_rot[P,4]
 *     o_neural_opacity[Nv*10], o_mask[Nv*10] (uint8) -- the reference's `neural_opacity`, `mask` */
CGS_API int cgs_neural_gaussians_forward(const float *packed_weights, const int32_t *vis_idx, int Nv,
                                         const float *anchor, const float *feat, const float *offsets,
                                         const float *scaling, const float *mask, const float *campos_host,
                                         float *o_xyz, float *o_color, float *o_opacity, float *o_scaling,
                                         float *o_rot, float *o_neural_opacity, uint8_t *o_mask, int32_t *count_dev,
                                         void *workspace, size_t workspace_bytes, void *stream);

/* The same operation with the two MLP layers on the tcgen05 tensor cores (3xTF32, activations resident
 * in tensor memory; csrc/neural_gaussians_umma.cu).  Identical arguments and outputs; the packed
 * weight block differs (TF32 hi/lo split, K-major core-matrix layout: contextgs_b200/neural_gaussians.py
 * pack_decoder_weights_umma).  *count_dev = -1 reports a tensor-core completion time-out. */
CGS_API int cgs_neural_gaussians_umma_packed_floats(void);
CGS_API size_t cgs_neural_gaussians_umma_workspace_bytes(int Nv);
CGS_API int cgs_neural_gaussians_umma_forward(const float *packed_weights, const int32_t *vis_idx, int Nv,
                                              const float *anchor, const float *feat, const float *offsets,
                                              const float *scaling, const float *mask, const float *campos_host,
                                              float *o_xyz, float *o_color, float *o_opacity, float *o_scaling,
                                              float *o_rot, float *o_neural_opacity, uint8_t *o_mask,
                                              int32_t *count_dev, void *workspace, size_t workspace_bytes,
                                              void *stream);

/* Training-mode variant of cgs_neural_gaussians_umma_forward: identical outputs, and additionally leaves behind what
 * the tcgen05 backward consumes, per visible row r: save_h[Nv,176] hidden activations (head h at column 56h),
 * save_hmask[Nv,6] their sign bits, save_pre2[Nv,144] layer-2 pre-activations (opacity 16 | colour 48 | covariance 80),
 * save_rowpos[Nv,2] tile-local rank of the first kept Gaussian of (row, half), save_tilebase[ceil(Nv/128)] rank of
 * each tile's first Gaussian (cgs_neural_gaussians_save_floats(0..4) returns 176, 6, 144, 2, 128). */
CGS_API int cgs_neural_gaussians_save_floats(int what);
CGS_API int cgs_neural_gaussians_umma_forward_train(const float *packed_weights, const int32_t *vis_idx, int Nv,
                                                    const float *anchor, const float *feat, const float *offsets,
                                                    const float *scaling, const float *mask, const float *campos_host,
                                                    float *o_xyz, float *o_color, float *o_opacity, float *o_scaling,
                                                    float *o_rot, float *o_neural_opacity, uint8_t *o_mask,
                                                    int32_t *count_dev, float *save_h, uint32_t *save_hmask,
                                                    float *save_pre2, uint32_t *save_rowpos, uint32_t *save_tilebase,
                                                    void *workspace, size_t workspace_bytes, void *stream);

/* Backward on the tcgen05 tensor cores (csrc/neural_gaussians_bwd_umma.cu), same gradients and conventions as
 * cgs_neural_gaussians_backward below (reference: autograd through gaussian_renderer/__init__.py:106-145 and
 * scene/gaussian_model.py:153-174), fed by the activations cgs_neural_gaussians_umma_forward_train saved: a
 * data-gradient kernel (dOut -> dH -> dX, 3xTF32, TMEM resident) and a weight-gradient kernel (SS-form MMAs with the
 * row index as the contraction, accumulators resident in TMEM across a persistent CTA).  packed_bwd:
 * cgs_neural_gaussians_bwd_umma_packed_floats() floats (contextgs_b200/neural_gaussians.py
 * pack_decoder_weights_bwd_umma).  scratch_dout[Nv,144], scratch_dpre[Nv,176]: hand-over between the two kernels.
 * *err (device, caller-zeroed) is set to 1 if a tensor-core completion barrier timed out. */
CGS_API int cgs_neural_gaussians_bwd_umma_packed_floats(void);
/* Diagnostic switches for timing experiments (results are WRONG while one is set; scripts/wgrad_probe.py).  key 1: the G1
 * weight-gradient kernel skips its tcgen05.mma instructions (bit 0), the converters' operand stores (bit 1), the bulk
 * copies (bit 2).  Returns -1 for an unknown key. */
CGS_API int cgs_debug_set(int key, int value);
CGS_API int cgs_neural_gaussians_backward_umma(const float *packed_bwd, const int32_t *vis_idx, int Nv,
                                               const float *anchor, const float *feat, const float *offsets,
                                               const float *scaling, const float *mask, const float *campos_host,
                                               const uint8_t *keep_mask, const float *save_h, const uint32_t *save_hmask,
                                               const float *save_pre2, const uint32_t *save_rowpos,
                                               const uint32_t *save_tilebase, const float *g_xyz, const float *g_color,
                                               const float *g_opacity, const float *g_scaling, const float *g_rot,
                                               float *d_anchor, float *d_feat, float *d_offsets, float *d_scaling,
                                               float *d_mask, float *d_packed_fwd, float *scratch_dout,
                                               float *scratch_dpre, int32_t *err, void *stream);

/* Backward of the two entry points above = what autograd does for gaussian_renderer/__init__.py:106-145
 * plus scene/gaussian_model.py:153-174 in the reference (SURVEY 8a row T1 lists the gradients train.py
 * consumes).  g_* are the gradients of the emitted Gaussians in emission order ([P,3] [P,3] [P] [P,3]
 * [P,4]); keep_mask is the forward's o_mask.  Rows of visible anchors of d_anchor[N,3] d_feat[N,50]
 * d_offsets[N,30] d_scaling[N,6] d_mask[N,10] are OVERWRITTEN (the caller zero-fills the arrays);
 * d_packed_fwd (cgs_neural_gaussians_packed_floats() floats, forward layout) is ACCUMULATED into.
 * packed_bwd: transposed weights, cgs_neural_gaussians_backward_packed_floats() floats
 * (contextgs_b200/neural_gaussians.py pack_decoder_weights_transposed). */
CGS_API int cgs_neural_gaussians_backward_packed_floats(void);
CGS_API size_t cgs_neural_gaussians_backward_workspace_bytes(int Nv);
CGS_API int cgs_neural_gaussians_backward(const float *packed_fwd, const float *packed_bwd, const int32_t *vis_idx,
                                          int Nv, const float *anchor, const float *feat, const float *offsets,
                                          const float *scaling, const float *mask, const float *campos_host,
                                          const uint8_t *keep_mask, const float *g_xyz, const float *g_color,
                                          const float *g_opacity, const float *g_scaling, const float *g_rot,
                                          float *d_anchor, float *d_feat, float *d_offsets, float *d_scaling,
                                          float *d_mask, float *d_packed_fwd, void *workspace, size_t workspace_bytes,
                                          void *stream);

/* Order-preserving compaction of a byte mask into an index list (the device-side replacement of
 * the reference's `tensor[bool_mask]` / torch.nonzero host-synchronising idiom, e.g.
 * gaussian_renderer/__init__.py:44-50).  mask must be 8-byte aligned. *count_dev = popcount. */
CGS_API size_t cgs_compact_workspace_bytes(int N);
CGS_API int cgs_compact_indices(const uint8_t *mask, int N, int32_t *out_idx, int32_t *count_dev, void *workspace,
                                size_t workspace_bytes, void *stream);

/* The same for `values[i] > 0` on int32 input (the radii of cgs_visible_filter): replaces
 * `visible_mask = radii_pure > 0` + boolean indexing (gaussian_renderer/__init__.py:287, :44-50).
 * values must be 16-byte aligned. */
CGS_API int cgs_compact_positive_i32(const int32_t *values, int N, int32_t *out_idx, int32_t *count_dev,
                                     void *workspace, size_t workspace_bytes, void *stream);

/* cgs_neural_gaussians_umma_forward with the visible-anchor count on the DEVICE: Nv is the capacity of
 * vis_idx, the kernel reads the count from *nv_dev (NULL = Nv), writes at most out_cap Gaussians (count_dev
 * always receives the true total, so the caller can detect an overflow), and skips the training-side
 * outputs when o_neural_opacity / o_mask are NULL. */
CGS_API int cgs_neural_gaussians_umma_forward_dev(const float *packed_weights, const int32_t *vis_idx, int Nv,
                                                  const int32_t *nv_dev, int64_t out_cap, const float *anchor,
                                                  const float *feat, const float *offsets, const float *scaling,
                                                  const float *mask, const float *campos_host, float *o_xyz,
                                                  float *o_color, float *o_opacity, float *o_scaling, float *o_rot,
                                                  float *o_neural_opacity, uint8_t *o_mask, int32_t *count_dev,
                                                  void *workspace, size_t workspace_bytes, void *stream);

/* One inference frame after prefilter_voxel = `render` of gaussian_renderer/__init__.py:155-229 on a decoded
 * model: cgs_neural_gaussians_umma_forward_dev followed by cgs_rasterize_forward_dev on its outputs, enqueued
 * by ONE call (no host work between the stages).  g_* are the capacity-sized (P_cap rows) Gaussian attribute
 * buffers written by the first stage and read by the second; g_count (device int32) is the Gaussian count.
 * All other arguments as in the two entry points. */
CGS_API int cgs_render_anchors_forward(const cgs_raster_settings *s, const float *packed_weights,
                                       const int32_t *vis_idx, int Nv_cap, const int32_t *nv_dev, int P_cap,
                                       const float *anchor, const float *feat, const float *offsets,
                                       const float *scaling, const float *mask, float *g_xyz, float *g_color,
                                       float *g_opacity, float *g_scaling, float *g_rot, int32_t *g_count,
                                       void *g1_workspace, size_t g1_workspace_bytes, int64_t R_cap, float *out_color,
                                       int32_t *radii, float *geom, uint32_t *point_list, uint32_t *ranges,
                                       float *final_T, uint32_t *n_contrib, int32_t *status, void *workspace,
                                       size_t workspace_bytes, void *stream);


/* ------------------------------------------------------------------ context / entropy model (E4-E7, G2) */

/* Floats per channel of the packed EntropyBottleneck parameters (softplus(matrices), biases,
 * tanh(factors), median) -- see oracle/entropy_ref.py EntropyBottleneckRef.packed(). */
CGS_API int cgs_eb_param_floats(void);

/* Replaces `pc.latent_codec(hyper, training=...)` = compressai EntropyBottleneck.forward
 * (reference call: scene/gaussian_model.py:1556).  hyper[N,C]; noise[N,C] (training) or NULL (eval:
 * round about the median); outputs hyper_q[N,C], likelihood[N,C] (floored at 1e-9).  If bit_sum is
 * given, sum(-log2 likelihood) over the anchors selected by choose[N] (NULL = all) is ADDED to it. */
CGS_API int cgs_eb_forward(const float *packed_params, int C, const float *hyper, const float *noise, int N,
                           float *hyper_q, float *likelihood, const uint8_t *choose, double *bit_sum, void *stream);

/* Packed context-MLP weights for one level: W1[in_dim][100] | b1[100] | W2[100][176] | b2[176]
 * (k-major; scene/gaussian_model.py:177-188).  in_dim is 71 (context 59 + hyper 12) or 15 (xyz + hyper). */
CGS_API int cgs_context_level_packed_floats(int in_dim);

/* One level of the coarse-to-fine autoregression, fused (replaces the loop body
 * scene/gaussian_model.py:1562-1652 plus the Entropy_gaussian calls at :1667-1670 and the sums at
 * :1685-1693 for the rows of this level):
 *   rows r < n_rows: orig_idx[r] = anchor coded by the row; ctx_src[r] = anchor whose already
 *   quantised (anchor, feat_q, scaling_q) form the row's context (in_dim 71), or level_anchor[r,3]
 *   (in_dim 15).  noise[n_rows,86] != NULL selects training-mode `x + U*Q`, NULL selects
 *   STE_multistep rounding (utils/encodings.py:203-213).  Quantised values are scattered into
 *   feat_q/scaling_q/offsets_q at orig_idx; bits of rows with choose[orig] != 0 are added to
 *   bit_sums[0..2] (feat, scaling, masked offsets) and their count to bit_sums[3] (fp64). */
CGS_API int cgs_context_level_forward(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                      const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                      const float *anchor, const float *hyper_q, const float *feat,
                                      const float *scaling, const float *offsets, const float *mask,
                                      const uint8_t *choose, const float *noise, float feat_mean, float scaling_mean,
                                      float offset_mean, float *feat_q, float *scaling_q, float *offsets_q,
                                      float *bits_out, double *bit_sums, void *stream);

/* The same level on the tcgen05 tensor cores (3xTF32, activations resident in tensor memory;
 * csrc/context_model_umma.cu).  Identical arguments and results; the packed block is the TF32 hi/lo split
 * in the K-major core-matrix layout (contextgs_b200/context_model.py pack_grid_weights_umma).
 * *err_flag (device int32) is set to 1 if a tensor-core completion barrier timed out. */
CGS_API int cgs_context_level_umma_packed_floats(int in_dim);
CGS_API int cgs_context_level_umma_forward(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                           const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                           const float *anchor, const float *hyper_q, const float *feat,
                                           const float *scaling, const float *offsets, const float *mask,
                                           const uint8_t *choose, const float *noise, float feat_mean,
                                           float scaling_mean, float offset_mean, float *feat_q, float *scaling_q,
                                           float *offsets_q, float *bits_out, double *bit_sums, int32_t *err_flag,
                                           void *stream);

/* The same level kernel for the BITSTREAM CODEC (replaces the per-level prediction inside the loops of
 * `conduct_encoding` / `conduct_decoding`, scene/gaussian_model.py:1112-1232,1380-1477):
 *   params_out[n_rows][176] (optional) receives what the entropy coder needs for every level row:
 *     mean[86] | scale[86] (raw MLP outputs, feat 50 | scaling 6 | offsets 30) | Q_feat Q_scaling Q_offsets | 0;
 *   predict_only != 0 : ONLY params_out is produced -- the decoder calls this before the level's attributes exist
 *     (feat / scaling / offsets / mask / offsets_q may be NULL; feat_q / scaling_q are read as context only);
 *   symbol_minmax (optional, device int32[6], ignored when predict_only): min / max of the coded symbols rint(value / Q) of
 *     the level's feat, scaling and (unmasked) offsets streams = the alphabets the bitstream codec needs (the same values
 *     cgs_codec_gauss_level_minmax computes in a separate pass); max < min marks an empty stream. */
CGS_API int cgs_context_level_umma_forward_ex(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                              const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                              const float *anchor, const float *hyper_q, const float *feat,
                                              const float *scaling, const float *offsets, const float *mask,
                                              const uint8_t *choose, const float *noise, float feat_mean,
                                              float scaling_mean, float offset_mean, float *feat_q, float *scaling_q,
                                              float *offsets_q, float *bits_out, double *bit_sums, int32_t *err_flag,
                                              float *params_out, int predict_only, int32_t *symbol_minmax,
                                              void *stream);

/* Training-mode variant of cgs_context_level_umma_forward (scene/gaussian_model.py:1596-1652 with training=True): same
 * outputs, and additionally params_out[n_rows,176] (mean[86] | scale[86] | Q_feat Q_scaling Q_offsets | 0, biases applied),
 * save_h[n_rows,112] (hidden activations) and save_hmask[n_rows,4] (their sign bits) for the tcgen05 backward. */
CGS_API int cgs_context_level_umma_forward_train(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                                 const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                                 const float *anchor, const float *hyper_q, const float *feat,
                                                 const float *scaling, const float *offsets, const float *mask,
                                                 const uint8_t *choose, const float *noise, float feat_mean,
                                                 float scaling_mean, float offset_mean, float *feat_q, float *scaling_q,
                                                 float *offsets_q, float *bits_out, double *bit_sums, int32_t *err_flag,
                                                 float *params_out, float *save_h, uint32_t *save_hmask, void *stream);

/* Backward of one level on the tcgen05 tensor cores (csrc/context_model_bwd_umma.cu): same gradients and in / out
 * conventions as cgs_context_level_backward below (reference: autograd through scene/gaussian_model.py:1596-1652,
 * 1666-1670 and utils/entropy_models.py:30-50,141-156), over ALL rows of the level, fed by what
 * cgs_context_level_umma_forward_train saved.  packed_bwd: cgs_context_level_bwd_umma_packed_floats(in_dim) floats
 * (contextgs_b200/context_model.py pack_grid_weights_bwd_umma: W2^T and W1^T as K-major B operands, TF32 hi / lo);
 * d_packed_w: the gradient in the layout of cgs_context_level_backward (accumulated).  scratch_dout[n_rows,176],
 * scratch_dpre[n_rows,112]: hand-over between the three kernels.  *err (device, caller-zeroed): completion time-out.
 * n_full: the level rows [0, n_full) are the ones chosen for the bit-rate term (the training forward orders them first);
 * rows [n_full, n_rows) take the reduced path (output gradient = the three step columns).  n_full = n_rows is always valid. */
CGS_API int cgs_context_level_bwd_umma_packed_floats(int in_dim);
CGS_API int cgs_context_level_backward_umma(int in_dim, const float *packed_bwd, const int32_t *orig_idx,
                                            const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                            const float *anchor, const float *hyper_q, const float *feat_q,
                                            const float *scaling_q, const float *offsets_q, const float *mask,
                                            const uint8_t *choose, const float *noise, float feat_mean,
                                            float scaling_mean, float offset_mean, const float *g_bits_dev,
                                            float bits_factor, const float *params, const float *save_h,
                                            const uint32_t *save_hmask, float *G_feat, float *G_scaling, float *G_offsets,
                                            float *d_mask, float *d_hyper_q, float *d_anchor, float *d_packed_w,
                                            float *scratch_dout, float *scratch_dpre, int32_t *err, int n_full,
                                            void *stream);

/* Backward of one level in TRAINING mode (noise != NULL in the forward): what autograd does in the
 * reference for the loop body scene/gaussian_model.py:1562-1652 plus the Entropy_gaussian terms of
 * bit_per_param (:1666-1693).  Launch fine -> coarse.  G_feat/G_scaling/G_offsets [N,*] hold the gradient
 * arriving on the quantised attributes and receive (a) in place, the gradient of the unquantised
 * attributes for the rows of this level, (b) by atomic add, the gradient flowing to the quantised
 * attributes of the coarser-level context sources.  d L / d bit_per_param is read from the device
 * scalar g_bits_dev (NULL = no rate term); bits_factor = mask_anchor_rate / (n_chosen * 86).
 * packed_w / d_packed_w use the backward layout W1[in][101] | b1[100] | W2[100][177] | b2[176]
 * (cgs_context_level_backward_packed_floats); d_packed_w, d_mask, d_anchor are accumulated into,
 * d_hyper_q rows of this level are overwritten.  ticket_dev: one device uint32 of scratch. */
CGS_API int cgs_context_level_backward_packed_floats(int in_dim);
CGS_API int cgs_context_level_backward(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                       const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                       const float *anchor, const float *hyper_q, const float *feat_q,
                                       const float *scaling_q, const float *offsets_q, const float *mask,
                                       const uint8_t *choose, const float *noise, float feat_mean, float scaling_mean,
                                       float offset_mean, const float *g_bits_dev, float bits_factor, float *G_feat,
                                       float *G_scaling, float *G_offsets, float *d_mask, float *d_hyper_q,
                                       float *d_anchor, float *d_packed_w, uint32_t *ticket_dev, void *stream);

/* The same backward restricted to the level rows row_list[0..n_rows) (NULL: rows 0..n_rows-1).  lite != 0 promises
 * that none of these rows is chosen for the bit-rate term (choose[orig] == 0): such rows reach the context MLP only
 * through the three adaptive quantisation steps, so the kernel back-propagates 3 of the 175 outputs (the host splits
 * every level into its chosen rows -- full kernel -- and the other ~85 % -- lite kernel). */
CGS_API int cgs_context_level_backward_rows(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                            const int32_t *ctx_src, const float *level_anchor, const int32_t *row_list,
                                            int n_rows, int lite, const float *anchor, const float *hyper_q,
                                            const float *feat_q, const float *scaling_q, const float *offsets_q,
                                            const float *mask, const uint8_t *choose, const float *noise, float feat_mean,
                                            float scaling_mean, float offset_mean, const float *g_bits_dev,
                                            float bits_factor, float *G_feat, float *G_scaling, float *G_offsets,
                                            float *d_mask, float *d_hyper_q, float *d_anchor, float *d_packed_w,
                                            uint32_t *ticket_dev, void *stream);

/* Backward of cgs_eb_forward's bit term: d_hyper[N,C] += w * d(-log2 likelihood)/d hyper_q for the chosen
 * anchors, d_packed_params[C,59] += the same w.r.t. the packed parameters (w = *g_bits_dev * bits_factor).
 * CompressAI's LowerBound gradient rule is followed. */
CGS_API int cgs_eb_backward(const float *packed_params, int C, const float *hyper_q, int N, const uint8_t *choose,
                            const float *g_bits_dev, float bits_factor, float *d_hyper, float *d_packed_params,
                            void *stream);

/* Replaces `Entropy_gaussian.forward` (utils/entropy_models.py:34-50) and its autograd backward
 * incl. `Low_bound` (:141-156).  x/mean/scale/bits are [n,D]; Q is [n] (q_per_elem=0) or [n,D]. */
CGS_API int cgs_gaussian_bits_forward(const float *x, const float *mean, const float *scale, const float *Q,
                                      int q_per_elem, float x_mean, int64_t n, int D, float *bits, void *stream);
CGS_API int cgs_gaussian_bits_backward(const float *x, const float *mean, const float *scale, const float *Q,
                                       int q_per_elem, float x_mean, int64_t n, int D, const float *grad_bits,
                                       float *dx, float *dmean, float *dscale, float *dQ, void *stream);

/* Replaces `STE_multistep.forward` (utils/encodings.py:203-213): x[n,D], Q[n] -> out[n,D]. */
CGS_API int cgs_ste_multistep(const float *x, const float *Q, int64_t n, int D, float *out, void *stream);

/* Replaces `Quantize_anchor.forward` (utils/encodings.py:219-227): anchors[n,3], HOST min/max[3]
 * -> anchors_q[n,3], quantized_v[n,3] (0..65535 as float). */
CGS_API int cgs_quantize_anchor(const float *anchors, const float *min_host, const float *max_host, int64_t n,
                                float *anchors_q, float *quantized_v, void *stream);

/* Level division of the anchor set: rows = round(points / voxel_size / level_scale)
 * (scene/gaussian_model.py:1760; points of anchors with keep[i] == 0 are first multiplied by 0, :1758-1759),
 * then what `torch_unique_with_indices` returns (utils/multi_level.py:3-31) for those rows:
 *   inverse[n]  : index of each point's row among the lexicographically SORTED unique rows,
 *   first[count]: minimum source index of every unique row   (capacity n),
 *   status_dev[0] = count, status_dev[1] = 1 if a rounded coordinate left the supported +-2^20 range.
 * keep may be NULL.  No host synchronisation; the library's own radix sort does the ordering. */
CGS_API size_t cgs_unique_voxels_workspace_bytes(int n);
CGS_API int cgs_unique_voxels(const float *points, const uint8_t *keep, int n, float voxel_size, float level_scale,
                              int32_t *inverse, int32_t *first, int32_t *status_dev, void *workspace,
                              size_t workspace_bytes, void *stream);

/* Stand-alone access to the library's own stable LSD radix sort of (uint32 key, uint32 value)
 * pairs on key bits [begin_bit, end_bit) -- exported for tests and for the level-division path.
 * n lives on the device (n_dev) and is bounded by n_cap; vals_in may be NULL (= 0..n-1).
 * The sorted result is written to keys_out/vals_out; keys_tmp/vals_tmp are ping-pong scratch. */
CGS_API size_t cgs_sort_workspace_bytes(int64_t n_cap, int begin_bit, int end_bit);
CGS_API int cgs_sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                       uint32_t *keys_tmp, uint32_t *vals_tmp, const uint32_t *n_dev, int64_t n_cap, int begin_bit,
                       int end_bit, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ bitstream codec
 * GPU replacement of `encoder_gaussian` / `decoder_gaussian` + torchac (utils/encodings.py:83-144; chunk loops
 * scene/gaussian_model.py:1192-1232,1422-1477), of `latent_codec.compress/decompress` (:1088,1331) and of the
 * binary-mask `encoder` / `decoder` (utils/encodings.py:147-183).  One GPU thread carries the (low, range) state of
 * one chunk of one stream through a byte-wise 32-bit range coder; Gaussian CDFs are evaluated in closed form from params
 * (cgs_context_level_umma_forward_ex), never tabulated.  Container format: this library's own
 * (csrc/entropy_codec.cu), torchac's is not pinnable.
 *
 * Gaussian streams, per level: attr 0 feat[.,50] / 1 scaling[.,6] / 2 offsets[.,30]; offsets whose mask[anchor][k/3]
 * is 0 are not coded and decode to 0.  The three streams of a level are coded by the same launches; chunk ids run over
 * them back to back (n_chunks[a] = ceil(n_rows / chunk_rows[a]) chunks of stream a; chunk c of stream a = level rows
 * [c*chunk_rows[a], (c+1)*chunk_rows[a])).  chunk_rows and n_chunks are HOST arrays of three ints.
 *   symbol = rint(value / Q); alphabet of stream a = minmax[2a], minmax[2a+1] (device int32 x6) = min / max symbol of the
 *   whole stream, computed by cgs_codec_gauss_level_minmax (parallel pass; cgs_context_level_umma_forward_ex returns the
 *   same bounds for free when it quantises the level) and stored with the stream;
 *   cgs_codec_gauss_level_chunks: chunk counts of the three streams; returns the 32-bit words of `scratch` an encode
 *   needs (chunk c of stream a sits at a fixed stride of cgs_codec_gauss_stream_capacity(a, chunk_rows[a]) bytes);
 *   cgs_codec_gauss_level_encode: pass 1 writes the 16-bit coding interval of each of the n_rows*86 values to `intervals`
 *   (uint32 each), pass 2 codes every chunk with ONE thread into its scratch slot and writes its byte count to
 *   stream_len[chunk id]; *err != 0 reports an uncodable symbol (1), an alphabet over 32768 (2), a capacity overflow (3);
 *   cgs_codec_gauss_level_pack: scratch slots -> packed + stream_off[chunk id] (stream_off = exclusive prefix sum of
 *   stream_len over the level, so the three streams are three consecutive slices of `packed`);
 *   cgs_codec_gauss_level_decode: reads chunk c of stream a at {feat,scaling,offsets}_bytes + stream_off[c] -
 *   stream_off[first chunk of a] and writes value = symbol * Q at {feat,scaling,offsets}_q[orig_idx[row]][k] --
 *   bit-identical to the encoder's input. */
CGS_API int64_t cgs_codec_gauss_stream_capacity(int attr, int chunk_rows);
/* The coder's normal CDF: Phi(z) = T[j] + f * (T[j+1] - T[j]) with t = clamp((z - z0) * inv_h, 0, n - 1), j = min((int)t,
 * n - 2), f = t - j, every operation one correctly rounded fp32 operation in this order; T = fp32(0.5 erfc(-z_j / sqrt 2))
 * at z_j = z0 + j / inv_h, T[0] = 0, T[n-1] = 1.  Copies T (n = 4097 entries) to HOST memory, and z0 / inv_h if not NULL.
 * Cumulative frequency of symbol boundary s: min(rn(Phi((((s - 0.5) * Q) - mean) * (1 / max(scale, 1e-9))) * M), M)
 * + (s - smin), M = 65536 - (smax - smin + 1). */
CGS_API int cgs_codec_phi_table(float *table, int n, float *z0, float *inv_h);
CGS_API int64_t cgs_codec_gauss_level_chunks(int n_rows, const int *chunk_rows, int32_t *n_chunks);
CGS_API int cgs_codec_gauss_level_minmax(const int32_t *orig_idx, int n_rows, const float *params, const float *mask,
                                         const float *feat_q, const float *scaling_q, const float *offsets_q,
                                         int32_t *minmax, void *stream);
CGS_API int cgs_codec_gauss_level_encode(const int32_t *orig_idx, int n_rows, const int *chunk_rows, const float *params,
                                         const float *mask, const float *feat_q, const float *scaling_q,
                                         const float *offsets_q, const int32_t *minmax, uint32_t *intervals,
                                         uint32_t *scratch, int32_t *stream_len, int32_t *err, void *stream);
CGS_API int cgs_codec_gauss_level_pack(int n_rows, const int *chunk_rows, const uint32_t *scratch, const int32_t *stream_len,
                                       const int64_t *stream_off, uint8_t *packed, void *stream);
CGS_API int cgs_codec_gauss_level_decode(const int32_t *orig_idx, int n_rows, const int *chunk_rows, const float *params,
                                         const float *mask, const uint8_t *feat_bytes, const uint8_t *scaling_bytes,
                                         const uint8_t *offsets_bytes, const int64_t *stream_off, const int32_t *stream_len,
                                         const int32_t *minmax, float *feat_q, float *scaling_q, float *offsets_q,
                                         void *stream);
/* Static-table streams: symbols[n_rows][C] int16 (index into the table of channel c, tables[c % T][0..table_ld),
 * cumulative 16-bit frequencies, tables[.][len] = 65536); chunking and outputs as above. */
CGS_API int cgs_codec_table_encode(const int16_t *symbols, int n_rows, int C, int chunk_rows, const uint32_t *tables,
                                   int T, int table_ld, uint32_t *scratch, int64_t cap_bytes, int32_t *stream_len,
                                   int32_t *err, void *stream);
CGS_API int cgs_codec_table_decode(const uint8_t *bytes, const int64_t *stream_off, const int32_t *stream_len,
                                   int n_rows, int C, int chunk_rows, const uint32_t *tables, const int32_t *table_len,
                                   int T, int table_ld, int16_t *symbols, void *stream);
/* Gathers the fixed-stride chunks of an encode call into one byte string: chunk c -> packed + stream_off[c]. */
CGS_API int cgs_codec_pack_streams(const uint32_t *scratch, int64_t cap_bytes, const int32_t *stream_len,
                                   const int64_t *stream_off, int n_streams, uint8_t *packed, void *stream);

/* ------------------------------------------------------------------ photometric loss
 * Fused replacement of `l1_loss` and `ssim` (utils/loss_utils.py:17-18,33-64; consumer train.py:200-204) for
 * [3,H,W] fp32 images: 11x11 Gaussian window (sigma 1.5, zero padding), C1 = 0.01^2, C2 = 0.03^2.
 *   forward : sums[0] = sum |img - gt|, sums[1] = sum of the SSIM map (device fp64; divide by 3*H*W for the means);
 *             dm / dp / dq [3,H,W] (optional, all or none) receive d ssim / d (G*x), d (G*x^2), d (G*xy) for the backward.
 *   backward: d_img = g_l1 * sign(img - gt) / n + g_ssim * d mean(ssim) / d img, n = 3*H*W; g_l1 / g_ssim are DEVICE
 *             scalars (the gradients arriving on the two means; NULL = term absent).  gt receives no gradient. */
CGS_API int cgs_l1_ssim_forward(const float *img, const float *gt, int H, int W, float *dm, float *dp, float *dq,
                                double *sums, void *stream);
CGS_API int cgs_l1_ssim_backward(const float *img, const float *gt, int H, int W, const float *dm, const float *dp,
                                 const float *dq, const float *g_l1, const float *g_ssim, float *d_img, void *stream);

/* ------------------------------------------------------------------ densification statistics
 * Replaces `GaussianModel.training_statis` (scene/gaussian_model.py:696-713; called every iteration in
 * 1500 < it < update_until, train.py:243) without its five boolean-index synchronisations:
 *   anchor_visible[N] uint8, offset_selection[n_slots = visible anchors * K] uint8 (the `mask` G1 returns),
 *   neural_opacity[n_slots], viewspace_grad[P,3] (dL/d means2D of the P emitted Gaussians), update_filter[P] uint8;
 *   opacity_accum[N] += sum_k max(opacity, 0), anchor_demon[N] += 1 for visible anchors;
 *   offset_gradient_accum[N*K] += |grad_xy|, offset_denom[N*K] += 1 for drawn Gaussians.
 * Masks must be 8-byte aligned. */
CGS_API size_t cgs_training_statis_workspace_bytes(int N, int K);
CGS_API int cgs_training_statis(int N, int K, const uint8_t *anchor_visible, const uint8_t *offset_selection, int n_slots,
                                const float *neural_opacity, const float *viewspace_grad, const uint8_t *update_filter,
                                int P, float *opacity_accum, float *anchor_demon, float *offset_gradient_accum,
                                float *offset_denom, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ anchor growing
 * One depth of `GaussianModel.anchor_growing` (scene/gaussian_model.py:778-816; called from adjust_anchor :856-863 every
 * update_interval iterations, train.py:246-247), replacing torch.unique(dim=0), the chunked O(|unique| x N)
 * occupancy comparison (:790-802) and torch_scatter.scatter_max (:812-816):
 *   anchor_q[N,3] = get_anchor, offset[N,K,3], scaling = get_scaling (row stride scaling_stride, columns 0..2 used),
 *   feat[N,feat_dim], hyper[N,hyper_dim], candidate[N*K] uint8 (8-byte aligned): the slots that pass the gradient
 *   threshold, the offset mask and the random thinning of :766-771 (computed by the caller, torch's own generator);
 *   cell = round((anchor_q + offset * scaling) / cur_size).int(); the sorted unique cells that no existing anchor's
 *   cell round(anchor_q / cur_size) occupies become new anchors, in sorted order:
 *   new_anchor[n_new,3] = cell * cur_size, new_feat / new_hyper = channel-wise maximum over the cell's candidates of
 *   the SOURCE anchor's rows.  cand_cap >= number of set candidate flags, new_cap >= n_new (<= cand_cap).
 *   status_dev[5]: [0] n_new, [1] 1 if a cell coordinate left +-2^20, [2] candidates, [3] unique cells,
 *                  [4] bit 0: n_new > new_cap, bit 1: more candidates than cand_cap (result truncated; an error).
 * No host synchronisation, no allocation. */
CGS_API size_t cgs_anchor_growing_workspace_bytes(int n_anchors, int n_offsets, int cand_cap);
CGS_API int cgs_anchor_growing(const float *anchor_q, const float *offset, const float *scaling, int scaling_stride,
                               const float *feat, int feat_dim, const float *hyper, int hyper_dim,
                               const uint8_t *candidate, int n_anchors, int n_offsets, float cur_size, int cand_cap,
                               float *new_anchor, float *new_feat, float *new_hyper, int new_cap, int32_t *status_dev,
                               void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------ initial anchor scales (create_from_pcd)
 * What `simple_knn._C.distCUDA2(points)` returns (third-party, not in the reference tree; call sites
 * scene/gaussian_model.py:389,407): mean_dist2[i] = mean of the three smallest squared distances from point i to the
 * OTHER points (FLT_MAX entries when n < 4, as simple-knn).  Exact neighbours through a uniform grid of edge `cell`
 * anchored at bbox_min_host[3] (host floats; (max - min) / cell must stay below 2^21), shells of cells around each
 * point, and an all-points scan for isolated points.
 *   status_dev[2]: [0] points finished by the all-points scan, [1] 1 if there were too many of them
 *   (> max(4096, n/32)): nothing was written for those, call again with a larger cell. */
CGS_API size_t cgs_knn3_workspace_bytes(int n);
CGS_API int cgs_knn3_mean_dist2(const float *points, int n, const float *bbox_min_host, float cell, float *mean_dist2,
                                int32_t *status_dev, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CONTEXTGS_B200_H */

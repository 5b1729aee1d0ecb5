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
    CGS_STATUS_WORDS = 8
};

CGS_API int cgs_abi_version(void);
CGS_API const char *cgs_last_error(void);

/* ------------------------------------------------------------------ rasterizer (P1, R0-R7) */

/* Replaces `GaussianRasterizer.visible_filter(means3D, scales, rotations)`
 * (reference call: gaussian_renderer/__init__.py:280-285; upstream kernel filter_preprocessCUDA).
 * radii[N] int32: screen radius, 0 when culled. */
CGS_API int cgs_visible_filter(const cgs_raster_settings *s, int N, const float *means3D, const float *scales,
                       const float *rotations, int32_t *radii, void *stream);

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
 * Pipeline: preprocess -> 4-pass radix sort of Gaussians by depth -> look-back scan of
 * tiles_touched in depth order -> emit (tile, id) -> radix sort by tile id -> ranges -> blend.
 * The resulting point_list/ranges are identical to the classical 64-bit (tile<<32|depth) sort. */
CGS_API int cgs_rasterize_forward(const cgs_raster_settings *s, int P, const float *means3D, const float *colors,
                          const float *opacities, const float *scales, const float *rotations, int64_t R_cap,
                          float *out_color, int32_t *radii, float *geom, uint32_t *point_list, uint32_t *ranges,
                          float *final_T, uint32_t *n_contrib, int32_t *status, void *workspace,
                          size_t workspace_bytes, void *stream);

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

/* Stand-alone access to the library's own stable LSD radix sort of (uint32 key, uint32 value)
 * pairs on key bits [begin_bit, end_bit) -- exported for tests and for the level-division path.
 * n lives on the device (n_dev) and is bounded by n_cap; vals_in may be NULL (= 0..n-1).
 * The sorted result is written to keys_out/vals_out; keys_tmp/vals_tmp are ping-pong scratch. */
CGS_API size_t cgs_sort_workspace_bytes(int64_t n_cap, int begin_bit, int end_bit);
CGS_API int cgs_sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                       uint32_t *keys_tmp, uint32_t *vals_tmp, const uint32_t *n_dev, int64_t n_cap, int begin_bit,
                       int end_bit, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CONTEXTGS_B200_H */

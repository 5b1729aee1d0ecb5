/*
 * oracle/raster_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement (OpenMP over Gaussians in preprocess and over pixel rows in the
 * forward blend -- every pixel/Gaussian is computed independently, results do not depend on the
 * thread count; the sort and the backward are single-thread) of the tile-based differentiable
 * Gaussian rasterizer that ContextGS calls through `diff_gaussian_rasterization`
 * (reference call sites: gaussian_renderer/__init__.py:179-205 `rasterizer(...)`
 * and :250-285 `rasterizer.visible_filter(...)`).
 *
 * PARITY UNPINNED: the rasterizer's source is NOT in /root/reference
 * (submodules/ is git-ignored there) and the reference
 * holds no tests or golden vectors for it.  This file restates the PUBLISHED
 * algorithm of graphdeco-inria/diff-gaussian-rasterization (+ the Scaffold-GS
 * `filter_preprocess` entry) from its documented constants:
 *   near cull z_view <= 0.2, 1/(w+1e-7), fov clamp 1.3x, +0.3 low-pass on the
 *   2D covariance diagonal, radius = ceil(3*sqrt(lambda_max)) with the
 *   max(0.1, .) guard, ndc2Pix = ((v+1)*S-1)/2, 16x16 tiles, rect =
 *   [(p-r)/16, (p+r+15)/16] clamped to the grid, key = tile<<32 | depth bits,
 *   stable LSD sort, alpha = min(0.99, o*exp(power)), skip alpha < 1/255,
 *   stop when T*(1-alpha) < 1e-4, out = C + T*bg, back-to-front backward that
 *   rebuilds T by division.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call this.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 * All arithmetic is IEEE fp32, one rounding per operation (no FMA contraction),
 * so that integer outputs (radii, tile rects, sort order, ranges) are exactly
 * reproducible by a device kernel compiled with -fmad=false that performs the
 * same operation sequence.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define NEAR_Z 0.2f
#define LOWPASS 0.3f
#define FOV_CLAMP 1.3f
#define ALPHA_MAX 0.99f
#define ALPHA_MIN (1.0f / 255.0f)
#define T_EPS 0.0001f

typedef struct {
    int W, H;
    float tanfovx, tanfovy;
    float bg[3];
    float scale_modifier;
    float view[16]; /* as the reference passes it: world_view_transform, i.e. the
                       transposed matrix stored row-major == column-major W2V */
    float proj[16]; /* full_proj_transform, same convention */
} ref_settings;

static inline float ndc2pix(float v, int S) { return ((v + 1.0f) * (float)S - 1.0f) * 0.5f; }

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

static void get_rect(float px, float py, int radius, int gx, int gy, int *x0, int *y0, int *x1, int *y1)
{
    float r = (float)radius;
    *x0 = imin(gx, imax(0, (int)((px - r) / (float)TILE)));
    *y0 = imin(gy, imax(0, (int)((py - r) / (float)TILE)));
    *x1 = imin(gx, imax(0, (int)((px + r + (float)(TILE - 1)) / (float)TILE)));
    *y1 = imin(gy, imax(0, (int)((py + r + (float)(TILE - 1)) / (float)TILE)));
}

/* R(q) for an UN-normalised quaternion (r,x,y,z); row-major standard rotation. */
static void quat_to_R(const float *q, float R[9])
{
    float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1.0f - 2.0f * (y * y + z * z);
    R[1] = 2.0f * (x * y - r * z);
    R[2] = 2.0f * (x * z + r * y);
    R[3] = 2.0f * (x * y + r * z);
    R[4] = 1.0f - 2.0f * (x * x + z * z);
    R[5] = 2.0f * (y * z - r * x);
    R[6] = 2.0f * (x * z - r * y);
    R[7] = 2.0f * (y * z + r * x);
    R[8] = 1.0f - 2.0f * (x * x + y * y);
}

/* Sigma = R diag(s^2) R^T, six unique entries (00,01,02,11,12,22).
 * Operation order: N = R*diag(s) first, then Sigma_ab = N_a0*N_b0 + N_a1*N_b1 + N_a2*N_b2. */
static void cov3d(const float *scale, float mod, const float *q, float S6[6], float N[9])
{
    float R[9];
    quat_to_R(q, R);
    float s0 = mod * scale[0], s1 = mod * scale[1], s2 = mod * scale[2];
    for (int i = 0; i < 3; ++i) {
        N[3 * i + 0] = R[3 * i + 0] * s0;
        N[3 * i + 1] = R[3 * i + 1] * s1;
        N[3 * i + 2] = R[3 * i + 2] * s2;
    }
    S6[0] = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
    S6[1] = N[0] * N[3] + N[1] * N[4] + N[2] * N[5];
    S6[2] = N[0] * N[6] + N[1] * N[7] + N[2] * N[8];
    S6[3] = N[3] * N[3] + N[4] * N[4] + N[5] * N[5];
    S6[4] = N[3] * N[6] + N[4] * N[7] + N[5] * N[8];
    S6[5] = N[6] * N[6] + N[7] * N[7] + N[8] * N[8];
}

/* Per-Gaussian projection state shared by forward and backward. */
typedef struct {
    float t[3];       /* view-space mean, x/y AFTER the fov clamp */
    float txtz, tytz; /* unclamped ratios */
    float A[6];       /* 2x3, A = J * Rv */
    float J00, J02, J11, J12;
    float cov[3]; /* cx, cy, cz (with low-pass) */
} proj_state;

static void cov2d(const float *p, const float *S6, const ref_settings *st, float fx, float fy, proj_state *ps)
{
    const float *vm = st->view;
    float tx = vm[0] * p[0] + vm[4] * p[1] + vm[8] * p[2] + vm[12];
    float ty = vm[1] * p[0] + vm[5] * p[1] + vm[9] * p[2] + vm[13];
    float tz = vm[2] * p[0] + vm[6] * p[1] + vm[10] * p[2] + vm[14];
    float limx = FOV_CLAMP * st->tanfovx, limy = FOV_CLAMP * st->tanfovy;
    ps->txtz = tx / tz;
    ps->tytz = ty / tz;
    tx = fminf(limx, fmaxf(-limx, ps->txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, ps->tytz)) * tz;
    ps->t[0] = tx; ps->t[1] = ty; ps->t[2] = tz;
    ps->J00 = fx / tz;
    ps->J02 = -(fx * tx) / (tz * tz);
    ps->J11 = fy / tz;
    ps->J12 = -(fy * ty) / (tz * tz);
    /* Rv(i,j) = vm[i + 4j] */
    for (int c = 0; c < 3; ++c) {
        ps->A[c] = ps->J00 * vm[0 + 4 * c] + ps->J02 * vm[2 + 4 * c];
        ps->A[3 + c] = ps->J11 * vm[1 + 4 * c] + ps->J12 * vm[2 + 4 * c];
    }
    const float *A = ps->A;
    /* B = A * Sigma (2x3) */
    float B00 = A[0] * S6[0] + A[1] * S6[1] + A[2] * S6[2];
    float B01 = A[0] * S6[1] + A[1] * S6[3] + A[2] * S6[4];
    float B02 = A[0] * S6[2] + A[1] * S6[4] + A[2] * S6[5];
    float B10 = A[3] * S6[0] + A[4] * S6[1] + A[5] * S6[2];
    float B11 = A[3] * S6[1] + A[4] * S6[3] + A[5] * S6[4];
    float B12 = A[3] * S6[2] + A[4] * S6[4] + A[5] * S6[5];
    ps->cov[0] = (B00 * A[0] + B01 * A[1] + B02 * A[2]) + LOWPASS;
    ps->cov[1] = B00 * A[3] + B01 * A[4] + B02 * A[5];
    ps->cov[2] = (B10 * A[3] + B11 * A[4] + B12 * A[5]) + LOWPASS;
}

/*
 * Forward preprocess.  filter_only != 0 is the Scaffold-GS `visible_filter`
 * entry (only radii are produced; reference call: gaussian_renderer/__init__.py:280).
 * Outputs (all [P] unless noted): radii i32, xy [P,2], depths, cov3D [P,6],
 * conic_opacity [P,4], tiles_touched u32.
 */
void ref_preprocess(const ref_settings *st, int P, const float *means, const float *scales, const float *rots,
                    const float *opac, int filter_only, int32_t *radii, float *xy, float *depths, float *cov3D,
                    float *conic_opacity, uint32_t *tiles_touched)
{
    int gx = (st->W + TILE - 1) / TILE, gy = (st->H + TILE - 1) / TILE;
    float fx = (float)st->W / (2.0f * st->tanfovx);
    float fy = (float)st->H / (2.0f * st->tanfovy);
    const float *vm = st->view, *pm = st->proj;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        radii[i] = 0;
        if (!filter_only) {
            tiles_touched[i] = 0;
            xy[2 * i] = xy[2 * i + 1] = 0.0f;
            depths[i] = 0.0f;
            for (int k = 0; k < 6; ++k) cov3D[6 * i + k] = 0.0f;
            for (int k = 0; k < 4; ++k) conic_opacity[4 * i + k] = 0.0f;
        }
        const float *p = means + 3 * i;
        float vz = vm[2] * p[0] + vm[6] * p[1] + vm[10] * p[2] + vm[14];
        if (vz <= NEAR_Z) continue;
        float hx = pm[0] * p[0] + pm[4] * p[1] + pm[8] * p[2] + pm[12];
        float hy = pm[1] * p[0] + pm[5] * p[1] + pm[9] * p[2] + pm[13];
        float hw = pm[3] * p[0] + pm[7] * p[1] + pm[11] * p[2] + pm[15];
        float pw = 1.0f / (hw + 0.0000001f);
        float S6[6], N[9];
        cov3d(scales + 3 * i, st->scale_modifier, rots + 4 * i, S6, N);
        proj_state ps;
        cov2d(p, S6, st, fx, fy, &ps);
        float cx = ps.cov[0], cy = ps.cov[1], cz = ps.cov[2];
        float det = cx * cz - cy * cy;
        if (det == 0.0f) continue;
        float det_inv = 1.0f / det;
        float mid = 0.5f * (cx + cz);
        float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
        float lambda1 = mid + disc, lambda2 = mid - disc;
        int my_radius = (int)ceilf(3.0f * sqrtf(fmaxf(lambda1, lambda2)));
        float px = ndc2pix(hx * pw, st->W), py = ndc2pix(hy * pw, st->H);
        int x0, y0, x1, y1;
        get_rect(px, py, my_radius, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        radii[i] = my_radius;
        if (filter_only) continue;
        depths[i] = vz;
        xy[2 * i] = px; xy[2 * i + 1] = py;
        for (int k = 0; k < 6; ++k) cov3D[6 * i + k] = S6[k];
        conic_opacity[4 * i + 0] = cz * det_inv;
        conic_opacity[4 * i + 1] = -cy * det_inv;
        conic_opacity[4 * i + 2] = cx * det_inv;
        conic_opacity[4 * i + 3] = opac[i];
        tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
    }
}

/* Stable LSD radix sort of (key64, val32) on bits [0, nbits). */
static void radix_sort_pairs(uint64_t *keys, uint32_t *vals, int64_t n, int nbits)
{
    uint64_t *k2 = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
    uint32_t *v2 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    for (int shift = 0; shift < nbits; shift += 8) {
        int64_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < n; ++i) cnt[((keys[i] >> shift) & 0xFF) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < n; ++i) {
            int64_t dst = cnt[(keys[i] >> shift) & 0xFF]++;
            k2[dst] = keys[i];
            v2[dst] = vals[i];
        }
        memcpy(keys, k2, sizeof(uint64_t) * (size_t)n);
        memcpy(vals, v2, sizeof(uint32_t) * (size_t)n);
    }
    free(k2);
    free(v2);
}

/* Number of (Gaussian, tile) instances = sum of tiles_touched. */
int64_t ref_count_instances(int P, const uint32_t *tiles_touched)
{
    int64_t R = 0;
    for (int i = 0; i < P; ++i) R += tiles_touched[i];
    return R;
}

/*
 * duplicateWithKeys + stable sort + identifyTileRanges.
 * keys_sorted [R] u64, point_list [R] u32, ranges [tiles,2] u32 (zero for empty tiles).
 */
void ref_bin(const ref_settings *st, int P, const int32_t *radii, const float *xy, const float *depths, int64_t R,
             uint64_t *keys_sorted, uint32_t *point_list, uint32_t *ranges)
{
    int gx = (st->W + TILE - 1) / TILE, gy = (st->H + TILE - 1) / TILE;
    int64_t off = 0;
    for (int i = 0; i < P; ++i) {
        if (radii[i] <= 0) continue;
        int x0, y0, x1, y1;
        get_rect(xy[2 * i], xy[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        uint32_t dbits;
        memcpy(&dbits, depths + i, 4);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                uint64_t key = (uint64_t)(y * gx + x);
                keys_sorted[off] = (key << 32) | dbits;
                point_list[off] = (uint32_t)i;
                ++off;
            }
    }
    int tiles = gx * gy, bit = 0;
    while ((1 << bit) < tiles + 1 && bit < 31) ++bit; /* enough bits for all tile ids */
    radix_sort_pairs(keys_sorted, point_list, R, 32 + bit);
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)tiles);
    for (int64_t i = 0; i < R; ++i) {
        uint32_t t = (uint32_t)(keys_sorted[i] >> 32);
        if (i == 0 || t != (uint32_t)(keys_sorted[i - 1] >> 32)) {
            ranges[2 * t] = (uint32_t)i;
            if (i > 0) ranges[2 * (keys_sorted[i - 1] >> 32) + 1] = (uint32_t)i;
        }
        if (i == R - 1) ranges[2 * t + 1] = (uint32_t)R;
    }
}

/* Per-tile front-to-back blend. out_color [3,H,W], final_T [H,W], n_contrib [H,W] u32. */
void ref_render_forward(const ref_settings *st, const uint32_t *ranges, const uint32_t *point_list, const float *xy,
                        const float *colors, const float *conic_opacity, float *out_color, float *final_T,
                        uint32_t *n_contrib)
{
    int W = st->W, H = st->H;
    int gx = (W + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            int tile = (py / TILE) * gx + (px / TILE);
            uint32_t b = ranges[2 * tile], e = ranges[2 * tile + 1];
            float T = 1.0f, C[3] = {0, 0, 0};
            uint32_t contributor = 0, last = 0;
            float fxp = (float)px, fyp = (float)py;
            for (uint32_t k = b; k < e; ++k) {
                uint32_t g = point_list[k];
                ++contributor;
                float dx = xy[2 * g] - fxp, dy = xy[2 * g + 1] - fyp;
                const float *co = conic_opacity + 4 * g;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                float alpha = fminf(ALPHA_MAX, co[3] * expf(power));
                if (alpha < ALPHA_MIN) continue;
                float test_T = T * (1.0f - alpha);
                if (test_T < T_EPS) break;
                for (int ch = 0; ch < 3; ++ch) C[ch] += colors[3 * g + ch] * alpha * T;
                T = test_T;
                last = contributor;
            }
            int pid = py * W + px;
            final_T[pid] = T;
            n_contrib[pid] = last;
            for (int ch = 0; ch < 3; ++ch) out_color[ch * H * W + pid] = C[ch] + T * st->bg[ch];
        }
}

/*
 * Per-tile back-to-front backward.  Accumulates (+=) into:
 *   dL_dxy [P,2]   true derivative wrt the pixel-space mean
 *   dL_dconic [P,3] true partial derivatives wrt (a, b, c) of
 *                  power = -0.5(a dx^2 + c dy^2) - b dx dy
 *   dL_dopacity [P], dL_dcolors [P,3].
 * Accumulation is in double so that the oracle is order-independent.
 */
void ref_render_backward(const ref_settings *st, const uint32_t *ranges, const uint32_t *point_list, const float *xy,
                         const float *colors, const float *conic_opacity, const float *final_T,
                         const uint32_t *n_contrib, const float *dL_dpix, double *dL_dxy, double *dL_dconic,
                         double *dL_dopacity, double *dL_dcolors)
{
    int W = st->W, H = st->H;
    int gx = (W + TILE - 1) / TILE;
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            int tile = (py / TILE) * gx + (px / TILE);
            uint32_t b = ranges[2 * tile], e = ranges[2 * tile + 1];
            int pid = py * W + px;
            float T_final = final_T[pid], T = T_final;
            uint32_t last_contributor = n_contrib[pid];
            float accum[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.0f;
            float dpix[3];
            for (int ch = 0; ch < 3; ++ch) dpix[ch] = dL_dpix[ch * H * W + pid];
            float bg_dot = st->bg[0] * dpix[0] + st->bg[1] * dpix[1] + st->bg[2] * dpix[2];
            float fxp = (float)px, fyp = (float)py;
            for (uint32_t k = b + last_contributor; k-- > b;) {
                uint32_t g = point_list[k];
                float dx = xy[2 * g] - fxp, dy = xy[2 * g + 1] - fyp;
                const float *co = conic_opacity + 4 * g;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                float G = expf(power);
                float alpha = fminf(ALPHA_MAX, co[3] * G);
                if (alpha < ALPHA_MIN) continue;
                T = T / (1.0f - alpha);
                float dchannel_dcolor = alpha * T;
                float dL_dalpha = 0.0f;
                for (int ch = 0; ch < 3; ++ch) {
                    float c = colors[3 * g + ch];
                    accum[ch] = last_alpha * last_color[ch] + (1.0f - last_alpha) * accum[ch];
                    last_color[ch] = c;
                    dL_dalpha += (c - accum[ch]) * dpix[ch];
                    dL_dcolors[3 * g + ch] += (double)(dchannel_dcolor * dpix[ch]);
                }
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                float dL_dG = co[3] * dL_dalpha;
                float gdx = G * dx, gdy = G * dy;
                float dG_ddelx = -gdx * co[0] - gdy * co[1];
                float dG_ddely = -gdy * co[2] - gdx * co[1];
                dL_dxy[2 * g + 0] += (double)(dL_dG * dG_ddelx);
                dL_dxy[2 * g + 1] += (double)(dL_dG * dG_ddely);
                dL_dconic[3 * g + 0] += (double)(-0.5f * gdx * dx * dL_dG);
                dL_dconic[3 * g + 1] += (double)(-gdx * dy * dL_dG);
                dL_dconic[3 * g + 2] += (double)(-0.5f * gdy * dy * dL_dG);
                dL_dopacity[g] += (double)(G * dL_dalpha);
            }
        }
}

/*
 * Preprocess backward: (dL_dxy, dL_dconic) -> dL_dmeans3D [P,3], dL_dscales [P,3],
 * dL_drots [P,4]; also dL_dmeans2D [P,3] in the upstream convention
 * (x: dL/dxy.x * 0.5*W, y: dL/dxy.y * 0.5*H, z: 0) which is what
 * `viewspace_points.grad` exposes to training_statis (scene/gaussian_model.py:710).
 * Follows the upstream chain: the fov-clamp mask multiplies dL/dt.x, dL/dt.y only;
 * dL/dt.z uses the clamped t (no extra clamp term); no gradient flows through depth.
 */
void ref_preprocess_backward(const ref_settings *st, int P, const float *means, const float *scales, const float *rots,
                             const int32_t *radii, const double *dL_dxy_d, const double *dL_dconic_d,
                             float *dL_dmeans2D, float *dL_dmeans, float *dL_dscales, float *dL_drots)
{
    float fx = (float)st->W / (2.0f * st->tanfovx);
    float fy = (float)st->H / (2.0f * st->tanfovy);
    const float *vm = st->view, *pm = st->proj;
    float mod = st->scale_modifier;
    for (int i = 0; i < P; ++i) {
        for (int k = 0; k < 3; ++k) dL_dmeans[3 * i + k] = dL_dscales[3 * i + k] = dL_dmeans2D[3 * i + k] = 0.0f;
        for (int k = 0; k < 4; ++k) dL_drots[4 * i + k] = 0.0f;
        if (!(radii[i] > 0)) continue;
        const float *p = means + 3 * i;
        float gxy0 = (float)dL_dxy_d[2 * i], gxy1 = (float)dL_dxy_d[2 * i + 1];
        float ga = (float)dL_dconic_d[3 * i], gb = (float)dL_dconic_d[3 * i + 1], gc = (float)dL_dconic_d[3 * i + 2];
        float S6[6], N[9];
        cov3d(scales + 3 * i, mod, rots + 4 * i, S6, N);
        proj_state ps;
        cov2d(p, S6, st, fx, fy, &ps);
        float cx = ps.cov[0], cy = ps.cov[1], cz = ps.cov[2];
        float det = cx * cz - cy * cy;
        float g_cx = 0, g_cy = 0, g_cz = 0;
        if (det != 0.0f) {
            float inv2 = 1.0f / (det * det);
            g_cx = inv2 * (-cz * cz * ga + cy * cz * gb - cy * cy * gc);
            g_cz = inv2 * (-cy * cy * ga + cx * cy * gb - cx * cx * gc);
            g_cy = inv2 * (2.0f * cy * cz * ga - (cx * cz + cy * cy) * gb + 2.0f * cx * cy * gc);
        }
        /* symmetric 2x2 gradient G2 = [[g_cx, g_cy/2],[g_cy/2, g_cz]] */
        float h = 0.5f * g_cy;
        const float *A = ps.A;
        /* GS = A^T G2 A (3x3 symmetric): dL/dSigma as a full matrix */
        float GA0[3], GA1[3]; /* G2 * A rows */
        for (int c = 0; c < 3; ++c) {
            GA0[c] = g_cx * A[c] + h * A[3 + c];
            GA1[c] = h * A[c] + g_cz * A[3 + c];
        }
        float GS[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) GS[3 * r + c] = A[r] * GA0[c] + A[3 + r] * GA1[c];
        /* Sigma = N N^T  =>  dL/dN = 2 GS N ; N_ij = R_ij * s_j */
        float R[9];
        quat_to_R(rots + 4 * i, R);
        float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
        float dN[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                dN[3 * r + c] = 2.0f * (GS[3 * r] * N[c] + GS[3 * r + 1] * N[3 + c] + GS[3 * r + 2] * N[6 + c]);
        float dR[9];
        for (int c = 0; c < 3; ++c) {
            dL_dscales[3 * i + c] = mod * (dN[c] * R[c] + dN[3 + c] * R[3 + c] + dN[6 + c] * R[6 + c]);
            for (int r = 0; r < 3; ++r) dR[3 * r + c] = dN[3 * r + c] * s[c];
        }
        {
            float r = rots[4 * i], x = rots[4 * i + 1], y = rots[4 * i + 2], z = rots[4 * i + 3];
            dL_drots[4 * i + 0] = 2.0f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
            dL_drots[4 * i + 1] = 2.0f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.0f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2.0f * x * dR[8]);
            dL_drots[4 * i + 2] = 2.0f * (-2.0f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2.0f * y * dR[8]);
            dL_drots[4 * i + 3] = 2.0f * (-2.0f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.0f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
        }
        /* cov = A Sigma A^T => dL/dA = 2 G2 A Sigma (2x3) */
        float Sfull[9] = {S6[0], S6[1], S6[2], S6[1], S6[3], S6[4], S6[2], S6[4], S6[5]};
        float dA[6];
        for (int c = 0; c < 3; ++c) {
            dA[c] = 2.0f * (GA0[0] * Sfull[c] + GA0[1] * Sfull[3 + c] + GA0[2] * Sfull[6 + c]);
            dA[3 + c] = 2.0f * (GA1[0] * Sfull[c] + GA1[1] * Sfull[3 + c] + GA1[2] * Sfull[6 + c]);
        }
        /* A = J Rv, Rv(i,j) = vm[i+4j]:  dL/dJ_rk = sum_c dA_rc * Rv(k,c) */
        float dJ00 = dA[0] * vm[0] + dA[1] * vm[4] + dA[2] * vm[8];
        float dJ02 = dA[0] * vm[2] + dA[1] * vm[6] + dA[2] * vm[10];
        float dJ11 = dA[3] * vm[1] + dA[4] * vm[5] + dA[5] * vm[9];
        float dJ12 = dA[3] * vm[2] + dA[4] * vm[6] + dA[5] * vm[10];
        float limx = FOV_CLAMP * st->tanfovx, limy = FOV_CLAMP * st->tanfovy;
        float xm = (ps.txtz < -limx || ps.txtz > limx) ? 0.0f : 1.0f;
        float ym = (ps.tytz < -limy || ps.tytz > limy) ? 0.0f : 1.0f;
        float tz = 1.0f / ps.t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        float dtx = xm * -fx * tz2 * dJ02;
        float dty = ym * -fy * tz2 * dJ12;
        float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.0f * fx * ps.t[0]) * tz3 * dJ02 +
                    (2.0f * fy * ps.t[1]) * tz3 * dJ12;
        float dm[3];
        for (int j = 0; j < 3; ++j) dm[j] = vm[0 + 4 * j] * dtx + vm[1 + 4 * j] * dty + vm[2 + 4 * j] * dtz;
        /* mean2D path */
        float hx = pm[0] * p[0] + pm[4] * p[1] + pm[8] * p[2] + pm[12];
        float hy = pm[1] * p[0] + pm[5] * p[1] + pm[9] * p[2] + pm[13];
        float hw = pm[3] * p[0] + pm[7] * p[1] + pm[11] * p[2] + pm[15];
        float mw = 1.0f / (hw + 0.0000001f);
        float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
        float g2x = gxy0 * (0.5f * (float)st->W), g2y = gxy1 * (0.5f * (float)st->H);
        dL_dmeans2D[3 * i] = g2x;
        dL_dmeans2D[3 * i + 1] = g2y;
        for (int j = 0; j < 3; ++j)
            dm[j] += (pm[4 * j] * mw - pm[4 * j + 3] * mul1) * g2x + (pm[4 * j + 1] * mw - pm[4 * j + 3] * mul2) * g2y;
        for (int j = 0; j < 3; ++j) dL_dmeans[3 * i + j] = dm[j];
    }
}

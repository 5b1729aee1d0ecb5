// Per-Gaussian projection kernels: preprocess (forward), visible_filter, mark_visible and the
// preprocess backward.  THIS FILE IS COMPILED WITH -fmad=false: every fp32 operation is rounded
// once, in source order, so that the integer outputs that drive binning (radius, tile rectangle,
// depth bits) are bit-reproducible on the CPU oracle (oracle/raster_ref.c, gcc -ffp-contract=off).
// The kernels are HBM-bound (116 B/Gaussian forward), so losing FMA contraction costs nothing.
//
// Replaces upstream preprocessCUDA / filter_preprocessCUDA / checkFrustum / computeCov2DCUDA-bwd
// (not in the reference tree; call sites gaussian_renderer/__init__.py:197-205,280-285).
#include "common.cuh"

namespace cgs {

constexpr float kNearZ = 0.2f;
constexpr float kLowPass = 0.3f;
constexpr float kFovClamp = 1.3f;

struct Proj2D {
    float tx, ty, tz;      // view-space mean, x/y after the fov clamp
    float txtz, tytz;      // unclamped ratios
    float A[6];            // J * Rv (2x3)
    float cx, cy, cz;      // 2D covariance incl. low-pass
};

__device__ __forceinline__ void quat_to_R(const float4 q, float R[9])
{
    const float r = q.x, x = q.y, y = q.z, z = q.w;
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

// Sigma = (R diag(s)) (R diag(s))^T ; N = R diag(s)
__device__ __forceinline__ void cov3d(const float3 scale, float mod, const float4 q, float S6[6], float N[9])
{
    float R[9];
    quat_to_R(q, R);
    const float s0 = mod * scale.x, s1 = mod * scale.y, s2 = mod * scale.z;
#pragma unroll
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

__device__ __forceinline__ void cov2d(const float3 p, const float S6[6], const CamParams &cam, Proj2D &o,
                                      float &J00, float &J02, float &J11, float &J12)
{
    const float *vm = cam.view;
    float tx = vm[0] * p.x + vm[4] * p.y + vm[8] * p.z + vm[12];
    float ty = vm[1] * p.x + vm[5] * p.y + vm[9] * p.z + vm[13];
    const float tz = vm[2] * p.x + vm[6] * p.y + vm[10] * p.z + vm[14];
    const float limx = kFovClamp * cam.tanfovx, limy = kFovClamp * cam.tanfovy;
    o.txtz = tx / tz;
    o.tytz = ty / tz;
    tx = fminf(limx, fmaxf(-limx, o.txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, o.tytz)) * tz;
    o.tx = tx; o.ty = ty; o.tz = tz;
    J00 = cam.focal_x / tz;
    J02 = -(cam.focal_x * tx) / (tz * tz);
    J11 = cam.focal_y / tz;
    J12 = -(cam.focal_y * ty) / (tz * tz);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        o.A[c] = J00 * vm[0 + 4 * c] + J02 * vm[2 + 4 * c];
        o.A[3 + c] = J11 * vm[1 + 4 * c] + J12 * vm[2 + 4 * c];
    }
    const float *A = o.A;
    const float B00 = A[0] * S6[0] + A[1] * S6[1] + A[2] * S6[2];
    const float B01 = A[0] * S6[1] + A[1] * S6[3] + A[2] * S6[4];
    const float B02 = A[0] * S6[2] + A[1] * S6[4] + A[2] * S6[5];
    const float B10 = A[3] * S6[0] + A[4] * S6[1] + A[5] * S6[2];
    const float B11 = A[3] * S6[1] + A[4] * S6[3] + A[5] * S6[4];
    const float B12 = A[3] * S6[2] + A[4] * S6[4] + A[5] * S6[5];
    o.cx = (B00 * A[0] + B01 * A[1] + B02 * A[2]) + kLowPass;
    o.cy = B00 * A[3] + B01 * A[4] + B02 * A[5];
    o.cz = (B10 * A[3] + B11 * A[4] + B12 * A[5]) + kLowPass;
}

__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0f) * (float)S - 1.0f) * 0.5f; }

__device__ __forceinline__ float3 load3(const float *p, int i)
{
    return make_float3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
}

// Shared front half of preprocess / filter: returns radius (0 = culled) and fills the outputs.
__device__ __forceinline__ int project_one(const CamParams &cam, const float3 p, const float3 scale, const float4 q,
                                           float &px, float &py, float &depth, float3 &conic, int &tiles,
                                           uint2 *rect = nullptr)
{
    const float *vm = cam.view, *pm = cam.proj;
    const float vz = vm[2] * p.x + vm[6] * p.y + vm[10] * p.z + vm[14];
    if (vz <= kNearZ) return 0;
    const float hx = pm[0] * p.x + pm[4] * p.y + pm[8] * p.z + pm[12];
    const float hy = pm[1] * p.x + pm[5] * p.y + pm[9] * p.z + pm[13];
    const float hw = pm[3] * p.x + pm[7] * p.y + pm[11] * p.z + pm[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    float S6[6], N[9];
    cov3d(scale, cam.scale_modifier, q, S6, N);
    Proj2D pr;
    float J00, J02, J11, J12;
    cov2d(p, S6, cam, pr, J00, J02, J11, J12);
    const float cx = pr.cx, cy = pr.cy, cz = pr.cz;
    const float det = cx * cz - cy * cy;
    if (det == 0.0f) return 0;
    const float det_inv = 1.0f / det;
    const float mid = 0.5f * (cx + cz);
    const float disc = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda1 = mid + disc, lambda2 = mid - disc;
    const int radius = (int)ceilf(3.0f * sqrtf(fmaxf(lambda1, lambda2)));
    px = ndc2pix(hx * pw, cam.W);
    py = ndc2pix(hy * pw, cam.H);
    int x0, y0, x1, y1;
    get_rect(px, py, radius, cam.grid_x, cam.grid_y, x0, y0, x1, y1);
    tiles = (x1 - x0) * (y1 - y0);
    if (tiles == 0) return 0;
    if (rect) *rect = make_uint2((uint32_t)x0 | ((uint32_t)x1 << 16), (uint32_t)y0 | ((uint32_t)y1 << 16));
    depth = vz;
    conic = make_float3(cz * det_inv, -cy * det_inv, cx * det_inv);
    return radius;
}

__global__ void __launch_bounds__(256)
preprocess_kernel(CamParams cam, const uint32_t *__restrict__ p_dev, const float *__restrict__ means, const float *__restrict__ colors,
                  const float *__restrict__ opac, const float *__restrict__ scales, const float *__restrict__ rots,
                  int32_t *__restrict__ radii, float *__restrict__ geom, uint32_t *__restrict__ depth_keys,
                  uint2 *__restrict__ rects)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int)*p_dev) return;  // P lives on the device: the host only knows a capacity
    const float3 p = load3(means, i);
    const float3 sc = load3(scales, i);
    const float4 q = reinterpret_cast<const float4 *>(rots)[i];
    float px = 0.f, py = 0.f, depth = 0.f;
    float3 conic = make_float3(0.f, 0.f, 0.f);
    int tiles = 0;
    uint2 rect = make_uint2(0u, 0u);   // tile rectangle {x0 | x1 << 16, y0 | y1 << 16}: all the binning stages need
    const int radius = project_one(cam, p, sc, q, px, py, depth, conic, tiles, &rect);
    rects[i] = radius > 0 ? rect : make_uint2(0u, 0u);
    float4 *g = reinterpret_cast<float4 *>(geom + (size_t)i * kGeomStride);
    if (radius > 0) {
        const float3 c = load3(colors, i);
        g[0] = make_float4(px, py, conic.x, conic.y);
        g[1] = make_float4(conic.z, opac[i], c.x, c.y);
        g[2] = make_float4(c.z, depth, __int_as_float(radius), __uint_as_float((uint32_t)tiles));
        depth_keys[i] = __float_as_uint(depth);
    } else {
        g[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        g[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        g[2] = make_float4(0.f, 0.f, 0.f, 0.f);
        depth_keys[i] = 0xFFFFFFFFu;  // culled Gaussians sort to the end and emit nothing
    }
    radii[i] = radius;
}

__global__ void __launch_bounds__(256)
filter_kernel(CamParams cam, int N, const float *__restrict__ means, const float *__restrict__ scales,
              const float *__restrict__ rots, int32_t *__restrict__ radii)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float px, py, depth;
    float3 conic;
    int tiles;
    radii[i] = project_one(cam, load3(means, i), load3(scales, i), reinterpret_cast<const float4 *>(rots)[i], px, py,
                           depth, conic, tiles);
}

// prefilter_voxel in ONE pass (gaussian_renderer/__init__.py:232-287): radius test of every anchor with
// scales = get_scaling[:, :3] (row stride `scale_stride` floats) and the rotation of anchor 0 for all
// anchors (the reference passes rotations[[0], :].repeat(N, 1)), fused with the ordered compaction
// of the visible anchors (decoupled look-back across 1024-anchor tiles).  Writes the boolean mask the
// reference returns AND the index list + device-side count the G1 kernel consumes, so neither
// `radii > 0` nor a separate compaction pass (nor the N x 4 rotation copy) ever touches HBM.
constexpr int kPrefilterItems = 4, kPrefilterTile = 256 * kPrefilterItems;

__global__ void __launch_bounds__(256)
prefilter_anchors_kernel(CamParams cam, int N, const float *__restrict__ means, const float *__restrict__ scales,
                         int scale_stride, const float *__restrict__ rot_row, uint8_t *__restrict__ visible,
                         int *__restrict__ out_idx, unsigned long long *scan_state, uint32_t *ticket,
                         int32_t *__restrict__ count_out)
{
    // item k of thread t is anchor tile*1024 + k*256 + t: coalesced loads, and the index order is (k, warp, lane)
    __shared__ uint32_t s_tile, s_cnt[kPrefilterItems * 8], s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 q = make_float4(rot_row[0], rot_row[1], rot_row[2], rot_row[3]);
    const float nrm = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);  // F.normalize
    q = make_float4(q.x / nrm, q.y / nrm, q.z / nrm, q.w / nrm);
    uint32_t ball[kPrefilterItems];
    bool vis[kPrefilterItems];
#pragma unroll
    for (int k = 0; k < kPrefilterItems; ++k) {
        const int i = tile * kPrefilterTile + k * 256 + threadIdx.x;
        vis[k] = false;
        if (i < N) {
            const float *sp = scales + (size_t)i * scale_stride;
            float px, py, depth;
            float3 conic;
            int tiles;
            vis[k] = project_one(cam, load3(means, i), make_float3(sp[0], sp[1], sp[2]), q, px, py, depth, conic, tiles) > 0;
            visible[i] = vis[k] ? 1 : 0;
        }
        ball[k] = __ballot_sync(0xffffffffu, vis[k]);
        if (lane == 0) s_cnt[k * 8 + warp] = __popc(ball[k]);
    }
    __syncthreads();
    if (warp == 0) {
        // exclusive scan of the 32 (item, warp) counts, then the chained scan across tiles
        const uint32_t c = s_cnt[lane];
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        s_cnt[lane] = incl - c;
        const uint64_t excl = lookback_exclusive(scan_state, tile, total);
        if (lane == 0) {
            s_base = (uint32_t)excl;
            if (tile == (N - 1) / kPrefilterTile) *count_out = (int32_t)(excl + total);
        }
    }
    __syncthreads();
    const uint32_t base = s_base, lt = (1u << lane) - 1;
#pragma unroll
    for (int k = 0; k < kPrefilterItems; ++k)
        if (vis[k]) out_idx[base + s_cnt[k * 8 + warp] + __popc(ball[k] & lt)] = tile * kPrefilterTile + k * 256 + threadIdx.x;
}

__global__ void __launch_bounds__(256)
mark_visible_kernel(CamParams cam, int N, const float *__restrict__ means, uint8_t *__restrict__ visible)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float3 p = load3(means, i);
    const float *vm = cam.view;
    const float vz = vm[2] * p.x + vm[6] * p.y + vm[10] * p.z + vm[14];
    visible[i] = vz > kNearZ ? 1 : 0;
}

// Backward of the projection.  acc[P,9] holds the render-backward accumulators
// {dL/dx, dL/dy, dL/da, dL/db, dL/dc, dL/dopacity, dL/dr, dL/dg, dL/db_col} (true partials).
__global__ void __launch_bounds__(256)
preprocess_backward_kernel(CamParams cam, int P, const float *__restrict__ means, const float *__restrict__ scales,
                           const float *__restrict__ rots, const int32_t *__restrict__ radii,
                           const float *__restrict__ acc, float *__restrict__ dL_dmeans,
                           float *__restrict__ dL_dmeans2D, float *__restrict__ dL_dcolors,
                           float *__restrict__ dL_dopacity, float *__restrict__ dL_dscales,
                           float *__restrict__ dL_drots)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float dm[3] = {0.f, 0.f, 0.f}, ds[3] = {0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
    float g2x = 0.f, g2y = 0.f, dop = 0.f, dcol[3] = {0.f, 0.f, 0.f};
    if (radii[i] > 0) {
        const float *a = acc + (size_t)i * 9;
        const float gxy0 = a[0], gxy1 = a[1], ga = a[2], gb = a[3], gc = a[4];
        dop = a[5];
        dcol[0] = a[6]; dcol[1] = a[7]; dcol[2] = a[8];
        const float3 p = load3(means, i);
        const float3 sc = load3(scales, i);
        const float4 q = reinterpret_cast<const float4 *>(rots)[i];
        const float mod = cam.scale_modifier;
        float S6[6], N[9];
        cov3d(sc, mod, q, S6, N);
        Proj2D pr;
        float J00, J02, J11, J12;
        cov2d(p, S6, cam, pr, J00, J02, J11, J12);
        const float cx = pr.cx, cy = pr.cy, cz = pr.cz;
        const float det = cx * cz - cy * cy;
        float g_cx = 0.f, g_cy = 0.f, g_cz = 0.f;
        if (det != 0.0f) {
            const float inv2 = 1.0f / (det * det);
            g_cx = inv2 * (-cz * cz * ga + cy * cz * gb - cy * cy * gc);
            g_cz = inv2 * (-cy * cy * ga + cx * cy * gb - cx * cx * gc);
            g_cy = inv2 * (2.0f * cy * cz * ga - (cx * cz + cy * cy) * gb + 2.0f * cx * cy * gc);
        }
        const float h = 0.5f * g_cy;
        const float *A = pr.A;
        float GA0[3], GA1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            GA0[c] = g_cx * A[c] + h * A[3 + c];
            GA1[c] = h * A[c] + g_cz * A[3 + c];
        }
        float GS[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) GS[3 * r + c] = A[r] * GA0[c] + A[3 + r] * GA1[c];
        float R[9];
        quat_to_R(q, R);
        const float s[3] = {mod * sc.x, mod * sc.y, mod * sc.z};
        float dN[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                dN[3 * r + c] = 2.0f * (GS[3 * r] * N[c] + GS[3 * r + 1] * N[3 + c] + GS[3 * r + 2] * N[6 + c]);
        float dR[9];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ds[c] = mod * (dN[c] * R[c] + dN[3 + c] * R[3 + c] + dN[6 + c] * R[6 + c]);
#pragma unroll
            for (int r = 0; r < 3; ++r) dR[3 * r + c] = dN[3 * r + c] * s[c];
        }
        {
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            dq[0] = 2.0f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
            dq[1] = 2.0f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.0f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] -
                            2.0f * x * dR[8]);
            dq[2] = 2.0f * (-2.0f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] -
                            2.0f * y * dR[8]);
            dq[3] = 2.0f * (-2.0f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.0f * z * dR[4] + y * dR[5] +
                            x * dR[6] + y * dR[7]);
        }
        const float Sf[9] = {S6[0], S6[1], S6[2], S6[1], S6[3], S6[4], S6[2], S6[4], S6[5]};
        float dA[6];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            dA[c] = 2.0f * (GA0[0] * Sf[c] + GA0[1] * Sf[3 + c] + GA0[2] * Sf[6 + c]);
            dA[3 + c] = 2.0f * (GA1[0] * Sf[c] + GA1[1] * Sf[3 + c] + GA1[2] * Sf[6 + c]);
        }
        const float *vm = cam.view, *pm = cam.proj;
        const float dJ00 = dA[0] * vm[0] + dA[1] * vm[4] + dA[2] * vm[8];
        const float dJ02 = dA[0] * vm[2] + dA[1] * vm[6] + dA[2] * vm[10];
        const float dJ11 = dA[3] * vm[1] + dA[4] * vm[5] + dA[5] * vm[9];
        const float dJ12 = dA[3] * vm[2] + dA[4] * vm[6] + dA[5] * vm[10];
        const float limx = kFovClamp * cam.tanfovx, limy = kFovClamp * cam.tanfovy;
        const float xm = (pr.txtz < -limx || pr.txtz > limx) ? 0.0f : 1.0f;
        const float ym = (pr.tytz < -limy || pr.tytz > limy) ? 0.0f : 1.0f;
        const float fx = cam.focal_x, fy = cam.focal_y;
        const float tz = 1.0f / pr.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = xm * -fx * tz2 * dJ02;
        const float dty = ym * -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.0f * fx * pr.tx) * tz3 * dJ02 +
                          (2.0f * fy * pr.ty) * tz3 * dJ12;
#pragma unroll
        for (int j = 0; j < 3; ++j) dm[j] = vm[0 + 4 * j] * dtx + vm[1 + 4 * j] * dty + vm[2 + 4 * j] * dtz;
        const float hx = pm[0] * p.x + pm[4] * p.y + pm[8] * p.z + pm[12];
        const float hy = pm[1] * p.x + pm[5] * p.y + pm[9] * p.z + pm[13];
        const float hw = pm[3] * p.x + pm[7] * p.y + pm[11] * p.z + pm[15];
        const float mw = 1.0f / (hw + 0.0000001f);
        const float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
        g2x = gxy0 * (0.5f * (float)cam.W);
        g2y = gxy1 * (0.5f * (float)cam.H);
#pragma unroll
        for (int j = 0; j < 3; ++j)
            dm[j] += (pm[4 * j] * mw - pm[4 * j + 3] * mul1) * g2x + (pm[4 * j + 1] * mw - pm[4 * j + 3] * mul2) * g2y;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        dL_dmeans[3 * i + j] = dm[j];
        dL_dscales[3 * i + j] = ds[j];
        dL_dcolors[3 * i + j] = dcol[j];
    }
    dL_dmeans2D[3 * i] = g2x;
    dL_dmeans2D[3 * i + 1] = g2y;
    dL_dmeans2D[3 * i + 2] = 0.f;
    dL_dopacity[i] = dop;
    reinterpret_cast<float4 *>(dL_drots)[i] = make_float4(dq[0], dq[1], dq[2], dq[3]);
}

// ---- launchers used by raster_api.cu ---------------------------------------------------
void launch_preprocess(const CamParams &cam, int P_cap, const uint32_t *p_dev, const float *means, const float *colors,
                       const float *opac, const float *scales, const float *rots, int32_t *radii, float *geom,
                       uint32_t *depth_keys, uint2 *rects, cudaStream_t st)
{
    if (P_cap <= 0) return;
    preprocess_kernel<<<(P_cap + 255) / 256, 256, 0, st>>>(cam, p_dev, means, colors, opac, scales, rots, radii, geom,
                                                           depth_keys, rects);
}

void launch_filter(const CamParams &cam, int N, const float *means, const float *scales, const float *rots,
                   int32_t *radii, cudaStream_t st)
{
    if (N <= 0) return;
    filter_kernel<<<(N + 255) / 256, 256, 0, st>>>(cam, N, means, scales, rots, radii);
}

void launch_prefilter_anchors(const CamParams &cam, int N, const float *means, const float *scales, int scale_stride,
                              const float *rot_row, uint8_t *visible, int *out_idx, unsigned long long *scan_state,
                              uint32_t *ticket, int32_t *count_out, cudaStream_t st)
{
    if (N <= 0) return;
    prefilter_anchors_kernel<<<(N + kPrefilterTile - 1) / kPrefilterTile, 256, 0, st>>>(cam, N, means, scales, scale_stride, rot_row, visible,
                                                              out_idx, scan_state, ticket, count_out);
}

void launch_mark_visible(const CamParams &cam, int N, const float *means, uint8_t *visible, cudaStream_t st)
{
    if (N <= 0) return;
    mark_visible_kernel<<<(N + 255) / 256, 256, 0, st>>>(cam, N, means, visible);
}

void launch_preprocess_backward(const CamParams &cam, int P, const float *means, const float *scales,
                                const float *rots, const int32_t *radii, const float *acc, float *dL_dmeans,
                                float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity, float *dL_dscales,
                                float *dL_drots, cudaStream_t st)
{
    if (P <= 0) return;
    preprocess_backward_kernel<<<(P + 255) / 256, 256, 0, st>>>(cam, P, means, scales, rots, radii, acc, dL_dmeans,
                                                                dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dscales,
                                                                dL_drots);
}

}  // namespace cgs

// Per-tile front-to-back alpha blending (forward) and back-to-front gradient pass (backward).
// One CTA (256 threads = 16x16 pixels) per tile; the tile's depth-ordered 48-byte Gaussian records
// are gathered into shared memory in 256-record batches with cp.async (LDGSTS, 3 x 16 B per
// thread), double-buffered so the gather of batch k+1 overlaps the blend of batch k; the blend
// loop reads the staged records as warp-broadcast LDS.128.
//
// Replaces upstream renderCUDA<3> forward/backward (forward.cu / backward.cu, not in the
// reference tree; call site gaussian_renderer/__init__.py:197-205).
#include "common.cuh"

namespace cgs {

constexpr int kBatch = 256;
constexpr float kAlphaMax = 0.99f;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kTEps = 0.0001f;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Stage one batch: thread t gathers the record of instance (first + t).
__device__ __forceinline__ void stage_batch(float4 *dst, const float *__restrict__ geom, uint32_t gid, bool valid)
{
    if (valid) {
        const float4 *src = reinterpret_cast<const float4 *>(geom + (size_t)gid * kGeomStride);
        cp_async16(dst + 0, src + 0);
        cp_async16(dst + 1, src + 1);
        cp_async16(dst + 2, src + 2);
    }
}

__global__ void __launch_bounds__(kTilePixels)
render_forward_kernel(const uint32_t *__restrict__ ranges, const uint32_t *__restrict__ point_list,
                      const float *__restrict__ geom, int W, int H, float bg0, float bg1, float bg2,
                      float *__restrict__ out_color, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib)
{
    __shared__ __align__(16) float4 s_rec[2][kBatch * 3];

    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int tx = threadIdx.x & (kTile - 1), ty = threadIdx.x / kTile;
    const int px = blockIdx.x * kTile + tx, py = blockIdx.y * kTile + ty;
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;

    const uint32_t rb = ranges[2 * tile], re = ranges[2 * tile + 1];
    const int todo = (int)(re - rb);
    const int rounds = (todo + kBatch - 1) / kBatch;

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t contributor = 0, last_contributor = 0;
    bool done = !inside;

    // prologue: ids for batch 0 and 1, records for batch 0
    uint32_t next_id = 0;
    {
        const bool v0 = (int)threadIdx.x < todo;
        const uint32_t id0 = v0 ? point_list[rb + threadIdx.x] : 0;
        stage_batch(&s_rec[0][threadIdx.x * 3], geom, id0, v0);
        cp_async_commit();
        const int i1 = kBatch + threadIdx.x;
        if (i1 < todo) next_id = point_list[rb + i1];
    }

    for (int r = 0; r < rounds; ++r) {
        const int buf = r & 1;
        // issue the gather of batch r+1 (its ids were fetched one round ago), prefetch ids of r+2
        {
            const int i1 = (r + 1) * kBatch + threadIdx.x;
            stage_batch(&s_rec[buf ^ 1][threadIdx.x * 3], geom, next_id, i1 < todo);
            cp_async_commit();
            const int i2 = (r + 2) * kBatch + threadIdx.x;
            if (i2 < todo) next_id = point_list[rb + i2];
        }
        cp_async_wait<1>();
        if (__syncthreads_count(done) == kTilePixels) break;

        const int count = min(kBatch, todo - r * kBatch);
        const float4 *rec = s_rec[buf];
        for (int j = 0; !done && j < count; ++j) {
            ++contributor;
            const float4 a = rec[3 * j + 0];  // x y ca cb
            const float4 b = rec[3 * j + 1];  // cc op r g
            const float dx = a.x - fx, dy = a.y - fy;
            const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(kAlphaMax, b.y * __expf(power));
            if (alpha < kAlphaMin) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < kTEps) {
                done = true;
                continue;
            }
            const float w = alpha * T;
            C0 += b.z * w;
            C1 += b.w * w;
            C2 += rec[3 * j + 2].x * w;
            T = test_T;
            last_contributor = contributor;
        }
        __syncthreads();  // everyone is done with s_rec[buf] before round r+1 overwrites it
    }
    cp_async_wait<0>();

    if (inside) {
        const size_t pid = (size_t)py * W + px;
        const size_t HW = (size_t)H * W;
        final_T[pid] = T;
        n_contrib[pid] = last_contributor;
        out_color[pid] = C0 + T * bg0;
        out_color[HW + pid] = C1 + T * bg1;
        out_color[2 * HW + pid] = C2 + T * bg2;
    }
}

// ---------------------------------------------------------------------------------------
// Backward.  Per pixel the upstream recurrence is followed exactly (T rebuilt by division,
// suffix colour `accum`, bg term).  Per-Gaussian gradients are reduced across the warp with
// shuffles BEFORE touching memory: one lane issues one red.global.add per value per warp
// instead of upstream's one atomic per pixel (32x fewer atomics, and whole warps that do not
// touch a Gaussian skip it with a single ballot).
// acc[P,9] = {dL/dx, dL/dy, dL/da, dL/db, dL/dc, dL/dopacity, dL/dr, dL/dg, dL/dblue}.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kTilePixels)
render_backward_kernel(const uint32_t *__restrict__ ranges, const uint32_t *__restrict__ point_list,
                       const float *__restrict__ geom, int W, int H, float bg0, float bg1, float bg2,
                       const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
                       const float *__restrict__ dL_dpix, float *__restrict__ acc)
{
    __shared__ __align__(16) float4 s_rec[kBatch * 3];
    __shared__ uint32_t s_id[kBatch];

    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int tx = threadIdx.x & (kTile - 1), ty = threadIdx.x / kTile;
    const int px = blockIdx.x * kTile + tx, py = blockIdx.y * kTile + ty;
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const int lane = threadIdx.x & 31;

    const uint32_t rb = ranges[2 * tile], re = ranges[2 * tile + 1];
    const int todo = (int)(re - rb);
    const int rounds = (todo + kBatch - 1) / kBatch;

    const size_t pid = (size_t)py * W + px;
    const size_t HW = (size_t)H * W;
    const float T_final = inside ? final_T[pid] : 0.f;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pid] : 0;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    if (inside) {
        d0 = dL_dpix[pid];
        d1 = dL_dpix[HW + pid];
        d2 = dL_dpix[2 * HW + pid];
    }
    const float bg_dot = bg0 * d0 + bg1 * d1 + bg2 * d2;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;

    // the deepest contributor of any pixel in this tile bounds the work
    __shared__ int s_max_contrib;
    if (threadIdx.x == 0) s_max_contrib = 0;
    __syncthreads();
    {
        int m = last_contributor;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(&s_max_contrib, m);
    }
    __syncthreads();
    const int max_contrib = s_max_contrib;  // positions >= max_contrib contribute to no pixel
    (void)rounds;

    // walk positions max_contrib-1 .. 0 in batches
    for (int hi = max_contrib; hi > 0; hi -= kBatch) {
        const int count = min(kBatch, hi);
        __syncthreads();
        if ((int)threadIdx.x < count) {
            const uint32_t gid = point_list[rb + hi - 1 - threadIdx.x];
            s_id[threadIdx.x] = gid;
            const float4 *src = reinterpret_cast<const float4 *>(geom + (size_t)gid * kGeomStride);
            s_rec[threadIdx.x * 3 + 0] = src[0];
            s_rec[threadIdx.x * 3 + 1] = src[1];
            s_rec[threadIdx.x * 3 + 2] = src[2];
        }
        __syncthreads();
        for (int j = 0; j < count; ++j) {
            const int pos = hi - 1 - j;  // 0-based position in the tile's list
            float g_x = 0.f, g_y = 0.f, g_a = 0.f, g_b = 0.f, g_c = 0.f, g_o = 0.f, g_r = 0.f, g_g = 0.f, g_bl = 0.f;
            bool active = false;
            if (pos < last_contributor) {
                const float4 a = s_rec[3 * j + 0];
                const float4 b = s_rec[3 * j + 1];
                const float dx = a.x - fx, dy = a.y - fy;
                const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
                if (power <= 0.0f) {
                    const float G = __expf(power);
                    const float alpha = fminf(kAlphaMax, b.y * G);
                    if (alpha >= kAlphaMin) {
                        active = true;
                        const float cb = s_rec[3 * j + 2].x;
                        T = T / (1.0f - alpha);
                        const float dch = alpha * T;
                        acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0;
                        acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1;
                        acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2;
                        lc0 = b.z; lc1 = b.w; lc2 = cb;
                        float dL_dalpha = (b.z - acc0) * d0 + (b.w - acc1) * d1 + (cb - acc2) * d2;
                        g_r = dch * d0; g_g = dch * d1; g_bl = dch * d2;
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                        const float dL_dG = b.y * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        g_x = dL_dG * (-gdx * a.z - gdy * a.w);
                        g_y = dL_dG * (-gdy * b.x - gdx * a.w);
                        g_a = -0.5f * gdx * dx * dL_dG;
                        g_b = -gdx * dy * dL_dG;
                        g_c = -0.5f * gdy * dy * dL_dG;
                        g_o = G * dL_dalpha;
                    }
                }
            }
            if (__ballot_sync(0xffffffffu, active) == 0u) continue;
            g_x = warp_sum(g_x); g_y = warp_sum(g_y); g_a = warp_sum(g_a);
            g_b = warp_sum(g_b); g_c = warp_sum(g_c); g_o = warp_sum(g_o);
            g_r = warp_sum(g_r); g_g = warp_sum(g_g); g_bl = warp_sum(g_bl);
            if (lane == 0) {
                float *dst = acc + (size_t)s_id[j] * 9;
                atomicAdd(dst + 0, g_x); atomicAdd(dst + 1, g_y); atomicAdd(dst + 2, g_a);
                atomicAdd(dst + 3, g_b); atomicAdd(dst + 4, g_c); atomicAdd(dst + 5, g_o);
                atomicAdd(dst + 6, g_r); atomicAdd(dst + 7, g_g); atomicAdd(dst + 8, g_bl);
            }
        }
    }
}

void launch_render_forward(const CamParams &cam, const uint32_t *ranges, const uint32_t *point_list,
                           const float *geom, float *out_color, float *final_T, uint32_t *n_contrib, cudaStream_t st)
{
    dim3 grid(cam.grid_x, cam.grid_y);
    render_forward_kernel<<<grid, kTilePixels, 0, st>>>(ranges, point_list, geom, cam.W, cam.H, cam.bg[0], cam.bg[1],
                                                        cam.bg[2], out_color, final_T, n_contrib);
}

void launch_render_backward(const CamParams &cam, const uint32_t *ranges, const uint32_t *point_list,
                            const float *geom, const float *final_T, const uint32_t *n_contrib, const float *dL_dpix,
                            float *acc, cudaStream_t st)
{
    dim3 grid(cam.grid_x, cam.grid_y);
    render_backward_kernel<<<grid, kTilePixels, 0, st>>>(ranges, point_list, geom, cam.W, cam.H, cam.bg[0], cam.bg[1],
                                                         cam.bg[2], final_T, n_contrib, dL_dpix, acc);
}

}  // namespace cgs

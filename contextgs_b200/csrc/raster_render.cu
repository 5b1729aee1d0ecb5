// Per-tile front-to-back alpha blending (forward) and back-to-front gradient pass (backward).
//
// One CTA (256 threads) per 16x16 tile, as in upstream renderCUDA<3> (forward.cu / backward.cu, not
// in the reference tree; call site gaussian_renderer/__init__.py:197-205) -- but the work inside a
// tile is organised for Blackwell's issue-bound regime (the loop is FP32/SFU-issue bound, not HBM
// bound: ncu r01 showed 90 % issue-slot utilisation at 1.3 % DRAM):
//
//   * each WARP owns an 8x4 pixel block, so a Gaussian's footprint maps to few warps;
//   * the tile's depth-ordered records are staged 256 at a time (one per thread, register
//     prefetch of the next batch) into double-buffered shared memory, pre-scaled to the log2
//     domain (one ex2 per evaluation, no extra multiplies);
//   * while staging, each thread runs an exact-bound test of ITS record against the 8 warp
//     blocks (is alpha >= 1/255 reachable anywhere in the block?) and publishes an 8-bit mask;
//     every warp then compacts the 256 masks into its own index list and blends only records
//     that can touch its pixels.  `point_list` / `ranges` remain the reference's bounding-square
//     binning, bit for bit; the masks only drop (record, block) pairs the per-pixel test would
//     reject for every pixel of the block, so image, final_T and n_contrib are unchanged;
//   * one __syncthreads per batch; warps whose 32 pixels have saturated skip their lists.
#include "common.cuh"

namespace cgs {

constexpr int kBatch = 256;
constexpr int kWarps = kTilePixels / 32;   // 8 warps, each an 8x4 pixel block: 2 across, 4 down
constexpr int kBlockW = 8, kBlockH = 4;
constexpr float kAlphaMax = 0.99f;
constexpr int kDirectLanes = 8;   // backward: warps with at most this many contributing pixels skip the warp reduction (measured: 3 -> 2.90 ms, 6..10 -> 2.80-2.82, 16 -> 3.4; reduction always: 2.95)
constexpr float kTEps = 0.0001f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Conservative test: can a Gaussian with centre (gx, gy), log2-domain conic (A2, B2, C2)
//   p2(u, v) = A2 u^2 + B2 u v + C2 v^2   (A2 = -a/2 log2e, B2 = -b log2e, C2 = -c/2 log2e),
// reach p2 >= thr2 (i.e. alpha >= 1/255) at ANY point of the pixel rectangle [x0,x1] x [y0,y1]?
// p2 is a concave quadratic in (u, v) = pixel - centre; over a rectangle that does not contain
// the centre its maximum lies on one of the (at most two) edges facing the centre, and on an edge
// it is a 1-D parabola whose clamped vertex is exact.
__device__ __forceinline__ bool rect_may_contribute(float gx, float gy, float A2, float B2, float C2, float thr2,
                                                    float x0, float y0, float x1, float y1)
{
    const float u0 = x0 - gx, u1 = x1 - gx, v0 = y0 - gy, v1 = y1 - gy;
    const bool in_x = u0 <= 0.0f && u1 >= 0.0f, in_y = v0 <= 0.0f && v1 >= 0.0f;
    if (in_x && in_y) return true;
    float best = -3.0e38f, mag = 0.0f;
    if (!in_x) {
        const float u = u0 > 0.0f ? u0 : u1;
        const float v = fminf(v1, fmaxf(v0, -0.5f * B2 * u / C2));
        const float m = A2 * u * u + C2 * v * v;   // <= 0
        const float q = m + B2 * u * v;
        if (q > best) { best = q; mag = -m; }
    }
    if (!in_y) {
        const float v = v0 > 0.0f ? v0 : v1;
        const float u = fminf(u1, fmaxf(u0, -0.5f * B2 * v / A2));
        const float m = A2 * u * u + C2 * v * v;
        const float q = m + B2 * u * v;
        if (q > best) { best = q; mag = -m; }
    }
    return best >= thr2 - (0.02f + 2.0e-5f * mag);  // margin covers the fp32 rounding of both sides
}

// 2^x for x <= 0 with flush-to-zero: one MUFU.EX2 (exp2f adds a denormal-range fix-up of 3 instructions;
// alphas that small are far below the 1/255 threshold anyway)
__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct StagedRecord {
    float4 a;  // x y A2 B2
    float4 b;  // C2 thr2 opacity r
    float2 c;  // g b
    uint32_t mask;  // bit w: warp block w may receive a contribution
};

// Build the staged form of one geom record and its warp-block mask for the tile at (tx0, ty0).
__device__ __forceinline__ StagedRecord stage_record(const float4 g0, const float4 g1, const float4 g2, bool valid,
                                                     int tx0, int ty0, int W, int H)
{
    StagedRecord s;
    const float ca = g0.z, cb = g0.w, cc = g1.x, op = g1.y;
    const float A2 = (-0.5f * kLog2e) * ca, B2 = -kLog2e * cb, C2 = (-0.5f * kLog2e) * cc;
    const float thr2 = -__log2f(255.0f * op);  // op <= 0 -> NaN/inf: `p2 >= thr2` is then never true
    s.a = make_float4(g0.x, g0.y, A2, B2);
    s.b = make_float4(C2, thr2, op, g1.z);
    s.c = make_float2(g1.w, g2.x);
    uint32_t mask = 0;
    if (valid && op > 0.0f && thr2 <= 0.02f) {
        const bool ellipse = ca > 0.0f && cc > 0.0f && ca * cc > cb * cb;
        if (!ellipse) {
            mask = 0xffu;  // degenerate conic: let the per-pixel test decide
        } else if (rect_may_contribute(g0.x, g0.y, A2, B2, C2, thr2, (float)tx0, (float)ty0,
                                       (float)min(tx0 + kTile - 1, W - 1), (float)min(ty0 + kTile - 1, H - 1))) {
            // axis-aligned box of the ellipse {p2 >= thr2} (inflated): most warp blocks are rejected by
            // four comparisons, the exact edge test only runs for blocks that overlap the box
            const float det = A2 * C2 - 0.25f * B2 * B2;                 // > 0
            const float t = fminf(thr2, 0.0f) - 0.05f;                    // < 0
            const float hx = sqrtf(t * C2 / det) * 1.01f + 0.01f, hy = sqrtf(t * A2 / det) * 1.01f + 0.01f;
            const float x_lo = g0.x - hx, x_hi = g0.x + hx, y_lo = g0.y - hy, y_hi = g0.y + hy;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const int bx = tx0 + (w & 1) * kBlockW, by = ty0 + (w >> 1) * kBlockH;
                const float bx1 = (float)min(bx + kBlockW - 1, W - 1), by1 = (float)min(by + kBlockH - 1, H - 1);
                if (bx < W && by < H && x_hi >= (float)bx && x_lo <= bx1 && y_hi >= (float)by && y_lo <= by1 &&
                    rect_may_contribute(g0.x, g0.y, A2, B2, C2, thr2, (float)bx, (float)by, bx1, by1))
                    mask |= 1u << w;
            }
        }
    } else if (valid && !(op <= 0.0f) && !(op > 0.0f)) {
        mask = 0xffu;  // NaN opacity: keep the reference's per-pixel behaviour
    }
    s.mask = mask;
    return s;
}

// Each warp compacts the batch's masks into its own ascending index list.  Returns the list length.
__device__ __forceinline__ int build_warp_list(const uint8_t *__restrict__ s_mask, uint8_t *__restrict__ list, int warp,
                                               int lane)
{
    int cnt = 0;
    const uint32_t lt = (1u << lane) - 1;
#pragma unroll
    for (int i = 0; i < kBatch / 32; ++i) {
        const uint32_t m = s_mask[i * 32 + lane];
        const bool bit = (m >> warp) & 1u;
        const uint32_t b = __ballot_sync(0xffffffffu, bit);
        if (bit) list[cnt + __popc(b & lt)] = (uint8_t)(i * 32 + lane);
        cnt += __popc(b);
    }
    __syncwarp();
    return cnt;
}

__global__ void __launch_bounds__(kTilePixels)
render_forward_kernel(const uint32_t *__restrict__ ranges, const uint32_t *__restrict__ point_list,
                      const float *__restrict__ geom, int W, int H, float bg0, float bg1, float bg2,
                      float *__restrict__ out_color, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib)
{
    // one 48-byte record per staged Gaussian {a, b, c, pad}: the blend loop forms ONE address per record and reads the three
    // parts at immediate offsets (three separate arrays cost an address computation each)
    struct __align__(16) Rec {
        float4 a, b;
        float2 c;
        float2 pad;
    };
    __shared__ Rec s_rec[2][kBatch];
    __shared__ uint8_t s_mask[2][kBatch];
    __shared__ uint8_t s_list[kWarps][kBatch];

    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * kTile, ty0 = blockIdx.y * kTile;
    const int px = tx0 + (warp & 1) * kBlockW + (lane & (kBlockW - 1));
    const int py = ty0 + (warp >> 1) * kBlockH + (lane >> 3);
    const bool inside = px < W && py < H;
    // A pixel that has saturated (or lies outside the image) is moved infinitely far away: every later record then fails
    // the `p2 >= thr` test on its own (p2 = -inf for any ellipse), so the loop carries no `done` flag.
    constexpr float kFar = -3.0e38f;
    float fx = inside ? (float)px : kFar;
    const float fy = (float)py;

    const uint32_t rb = ranges[2 * tile], re = ranges[2 * tile + 1];
    const int todo = (int)(re - rb);
    const int rounds = (todo + kBatch - 1) / kBatch;

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last_contributor = 0;

    // software pipeline: the record of batch r+1 and the id of batch r+2 are in flight while batch r blends
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, g2 = g0;
    bool gv = (int)threadIdx.x < todo;
    uint32_t next_id = 0;
    if (gv) {
        const uint32_t id0 = point_list[rb + threadIdx.x];
        const float4 *src = reinterpret_cast<const float4 *>(geom + (size_t)id0 * kGeomStride);
        g0 = __ldg(src); g1 = __ldg(src + 1); g2 = __ldg(src + 2);
    }
    if (kBatch + (int)threadIdx.x < todo) next_id = point_list[rb + kBatch + threadIdx.x];

    for (int r = 0; r < rounds; ++r) {
        const int buf = r & 1;
        {
            const StagedRecord s = stage_record(g0, g1, g2, gv, tx0, ty0, W, H);
            s_rec[buf][threadIdx.x].a = s.a;
            s_rec[buf][threadIdx.x].b = s.b;
            s_rec[buf][threadIdx.x].c = s.c;
            s_mask[buf][threadIdx.x] = (uint8_t)s.mask;
            const int i1 = (r + 1) * kBatch + threadIdx.x;
            gv = i1 < todo;
            if (gv) {
                const float4 *src = reinterpret_cast<const float4 *>(geom + (size_t)next_id * kGeomStride);
                g0 = __ldg(src); g1 = __ldg(src + 1); g2 = __ldg(src + 2);
            }
            const int i2 = (r + 2) * kBatch + threadIdx.x;
            if (i2 < todo) next_id = point_list[rb + i2];
        }
        // one barrier per batch: publishes buffer `buf`; buffer buf^1 (batch r-1) is free again
        // because every thread finished blending it before arriving here
        const bool done = fx == kFar;     // saturated / outside pixels carry their state in fx
        if (__syncthreads_count(done) == kTilePixels) break;
        if (__all_sync(0xffffffffu, done)) continue;

        const int n = build_warp_list(s_mask[buf], s_list[warp], warp, lane);
        const Rec *recs = s_rec[buf];
        const uint8_t *list = s_list[warp];
        const uint32_t pos_base = (uint32_t)(r * kBatch + 1);
        for (int j = 0; j < n; ++j) {
            const int idx = list[j];
            const Rec *rec = recs + idx;
            const float4 a = rec->a;   // x y A2 B2
            const float4 b = rec->b;   // C2 thr2 op r
            const float dx = a.x - fx, dy = a.y - fy;
            const float p2 = dx * (a.z * dx + a.w * dy) + b.x * (dy * dy);
            // skip: alpha < 1/255 or the reference's `power > 0` guard (a saturated pixel sits at -3e38: p2 = -inf)
            if (!(p2 >= b.y) || p2 > 0.0f) continue;
            const float alpha = fminf(kAlphaMax, b.z * fast_exp2(p2));
            const float test_T = T * (1.0f - alpha);
            if (test_T < kTEps) {
                fx = kFar;
            } else {
                const float2 c = rec->c;
                const float w = alpha * T;
                C0 += b.w * w;
                C1 += c.x * w;
                C2 += c.y * w;
                T = test_T;
                last_contributor = pos_base + (uint32_t)idx;
            }
        }
    }

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
// suffix colour `accum`, bg term); the skip tests are the forward's, on the same staged values.
// Per-Gaussian gradients are reduced across the warp BEFORE touching memory with a
// transpose-reduce butterfly (8 values in 4+2+1+1+1 = 9 shuffles instead of 40; the 9th value
// with a plain 5-step reduction), then 9 lanes issue one red.global.add each: one atomic per
// value per warp instead of upstream's one per pixel.
// acc[P,9] = {dL/dx, dL/dy, dL/da, dL/db, dL/dc, dL/dopacity, dL/dr, dL/dg, dL/dblue}.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// After the call, lane L with (L & 3) == 0 holds the warp sum of v[id], id = (L >> 2) bit-reversed
// over 3 bits: id = 4*bit4(L) + 2*bit3(L) + bit2(L).
__device__ __forceinline__ float warp_transpose_reduce8(float (&v)[8], int lane)
{
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h16 ? v[i] : v[i + 4];
        const float keep = h16 ? v[i + 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h8 ? v[i] : v[i + 2];
        const float keep = h8 ? v[i + 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const float send = h4 ? v[0] : v[1];
        const float keep = h4 ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

__global__ void __launch_bounds__(kTilePixels)
render_backward_kernel(const uint32_t *__restrict__ ranges, const uint32_t *__restrict__ point_list,
                       const float *__restrict__ geom, int W, int H, float bg0, float bg1, float bg2,
                       const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
                       const float *__restrict__ dL_dpix, float *__restrict__ acc)
{
    struct __align__(16) Rec {      // one 48-byte staged record: one address per record in the loop below
        float4 a, b;
        float2 c;
        uint32_t id, pad;
    };
    __shared__ Rec s_rec[2][kBatch];
    __shared__ uint8_t s_mask[2][kBatch];
    __shared__ uint8_t s_list[kWarps][kBatch];
    __shared__ int s_max_contrib;

    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * kTile, ty0 = blockIdx.y * kTile;
    const int px = tx0 + (warp & 1) * kBlockW + (lane & (kBlockW - 1));
    const int py = ty0 + (warp >> 1) * kBlockH + (lane >> 3);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;

    const uint32_t rb = ranges[2 * tile];

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

    // the deepest contributor of any pixel in this tile / warp block bounds the work
    if (threadIdx.x == 0) s_max_contrib = 0;
    __syncthreads();
    int warp_max = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_max = max(warp_max, __shfl_xor_sync(0xffffffffu, warp_max, o));
    if (lane == 0) atomicMax(&s_max_contrib, warp_max);
    __syncthreads();
    const int max_contrib = s_max_contrib;  // list positions >= max_contrib contribute to no pixel
    const int rounds = (max_contrib + kBatch - 1) / kBatch;

    // batch r covers list positions hi-1 .. hi-count (descending), hi = max_contrib - r*kBatch;
    // thread t stages position hi-1-t
    float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, g2 = g0;
    uint32_t gid = 0, next_id = 0;
    bool gv = (int)threadIdx.x < max_contrib;
    if (gv) {
        gid = point_list[rb + max_contrib - 1 - threadIdx.x];
        const float4 *src = reinterpret_cast<const float4 *>(geom + (size_t)gid * kGeomStride);
        g0 = __ldg(src); g1 = __ldg(src + 1); g2 = __ldg(src + 2);
    }
    if (max_contrib - kBatch - 1 - (int)threadIdx.x >= 0) next_id = point_list[rb + max_contrib - kBatch - 1 - threadIdx.x];

    for (int r = 0; r < rounds; ++r) {
        const int buf = r & 1;
        const int hi = max_contrib - r * kBatch;
        {
            const StagedRecord s = stage_record(g0, g1, g2, gv, tx0, ty0, W, H);
            s_rec[buf][threadIdx.x].a = s.a;
            s_rec[buf][threadIdx.x].b = s.b;
            s_rec[buf][threadIdx.x].c = s.c;
            s_rec[buf][threadIdx.x].id = gid;
            s_mask[buf][threadIdx.x] = (uint8_t)s.mask;
            const int p1 = hi - kBatch - 1 - (int)threadIdx.x;
            gv = p1 >= 0;
            gid = next_id;
            if (gv) {
                const float4 *src = reinterpret_cast<const float4 *>(geom + (size_t)gid * kGeomStride);
                g0 = __ldg(src); g1 = __ldg(src + 1); g2 = __ldg(src + 2);
            }
            const int p2i = hi - 2 * kBatch - 1 - (int)threadIdx.x;
            if (p2i >= 0) next_id = point_list[rb + p2i];
        }
        __syncthreads();
        if (hi - kBatch >= warp_max) continue;  // the whole batch lies behind this warp's deepest contributor

        const int n = build_warp_list(s_mask[buf], s_list[warp], warp, lane);
        const Rec *recs = s_rec[buf];
        const uint8_t *list = s_list[warp];
        for (int j = 0; j < n; ++j) {
            const int idx = list[j];
            const Rec *rec = recs + idx;
            const int pos = hi - 1 - idx;  // 0-based position in the tile's list
            float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            float g_bl = 0.f;
            bool active = false;
            if (pos < last_contributor) {
                const float4 a = rec->a;   // x y A2 B2
                const float4 b = rec->b;   // C2 thr2 op r
                const float dx = a.x - fx, dy = a.y - fy;
                const float p2 = dx * (a.z * dx + a.w * dy) + b.x * (dy * dy);
                if (p2 >= b.y && !(p2 > 0.0f)) {
                    active = true;
                    const float2 c = rec->c;
                    const float G = fast_exp2(p2);
                    const float alpha = fminf(kAlphaMax, b.z * G);
                    const float ca = a.z * (-2.0f * kLn2), cb = a.w * (-kLn2), cc = b.x * (-2.0f * kLn2);
                    T = T / (1.0f - alpha);
                    const float dch = alpha * T;
                    acc0 = last_alpha * lc0 + (1.0f - last_alpha) * acc0;
                    acc1 = last_alpha * lc1 + (1.0f - last_alpha) * acc1;
                    acc2 = last_alpha * lc2 + (1.0f - last_alpha) * acc2;
                    lc0 = b.w; lc1 = c.x; lc2 = c.y;
                    float dL_dalpha = (b.w - acc0) * d0 + (c.x - acc1) * d1 + (c.y - acc2) * d2;
                    g[6] = dch * d0; g[7] = dch * d1; g_bl = dch * d2;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                    const float dL_dG = b.z * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    g[0] = dL_dG * (-gdx * ca - gdy * cb);
                    g[1] = dL_dG * (-gdy * cc - gdx * cb);
                    g[2] = -0.5f * gdx * dx * dL_dG;
                    g[3] = -gdx * dy * dL_dG;
                    g[4] = -0.5f * gdy * dy * dL_dG;
                    g[5] = G * dL_dalpha;
                }
            }
            const uint32_t act = __ballot_sync(0xffffffffu, active);
            if (act == 0u) continue;
            float *dst = acc + (size_t)rec->id * 9;
            if (__popc(act) <= kDirectLanes) {
                // a record that reaches only a few pixels of the block (most small Gaussians): their lanes add straight
                // into the accumulator -- nine predicated atomics instead of the 14-shuffle reduction
                if (active) {
#pragma unroll
                    for (int v = 0; v < 8; ++v) atomicAdd(dst + v, g[v]);
                    atomicAdd(dst + 8, g_bl);
                }
                continue;
            }
            const float s8 = warp_transpose_reduce8(g, lane);
            g_bl = warp_sum(g_bl);
            if ((lane & 3) == 0) {
                const int id = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                atomicAdd(dst + id, s8);
            } else if (lane == 1) {
                atomicAdd(dst + 8, g_bl);
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

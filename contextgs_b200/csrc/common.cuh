// Shared helpers for the contextgs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/contextgs_b200.h"

namespace cgs {

constexpr int kTile = CGS_TILE;
constexpr int kTilePixels = kTile * kTile;
constexpr int kGeomStride = CGS_GEOM_STRIDE;
constexpr int kNumSMs = 148;  // B200

// geom record slots
enum { G_X = 0, G_Y, G_CA, G_CB, G_CC, G_OP, G_R, G_G, G_B, G_DEPTH, G_RADIUS, G_TILES };

void set_error(const char *fmt, ...);
int check_launch(const char *what);

// ---- per-stage device timing + launch accounting (raster_api.cu) ------------------------
// Every launcher wraps its kernels in a StageScope.  The kernel-launch counter is always kept;
// CUDA events are recorded around the stage only after cgs_stage_timing_enable(1) (bench.py's
// separate roofline pass), so the product path pays one predictable branch.
enum Stage {
    ST_FILTER = 0, ST_COMPACT, ST_G1_FWD, ST_G1_BWD, ST_PREPROCESS, ST_DEPTH_SORT, ST_SCAN, ST_EMIT, ST_TILE_SORT,
    ST_RANGES, ST_RENDER_FWD, ST_RENDER_BWD, ST_PRE_BWD, ST_EB, ST_CTX_LEVEL, ST_CTX_LEVEL_BWD, ST_BITS, ST_ELEMWISE,
    ST_LEVEL_DIVIDE, ST_CODEC, ST_LOSS, ST_DENSIFY, ST_COUNT
};
struct StageScope {
    int stage;
    cudaStream_t st;
    int slot;
    StageScope(int stage, cudaStream_t st, int kernels);
    ~StageScope();
};

#define CGS_CHECK_PTR(p)                                   \
    do {                                                   \
        if ((p) == nullptr) {                              \
            cgs::set_error("%s: null pointer " #p, __func__); \
            return -1;                                     \
        }                                                  \
    } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Camera constants passed by value to kernels (lives in the constant bank).
struct CamParams {
    float view[16];
    float proj[16];
    float tanfovx, tanfovy;
    float focal_x, focal_y;
    float scale_modifier;
    int W, H, grid_x, grid_y;
    float bg[3];
};

static inline CamParams make_cam(const cgs_raster_settings *s)
{
    CamParams c;
    for (int i = 0; i < 16; ++i) {
        c.view[i] = s->viewmatrix[i];
        c.proj[i] = s->projmatrix[i];
    }
    c.tanfovx = s->tanfovx;
    c.tanfovy = s->tanfovy;
    c.W = s->image_width;
    c.H = s->image_height;
    c.focal_x = (float)c.W / (2.0f * c.tanfovx);
    c.focal_y = (float)c.H / (2.0f * c.tanfovy);
    c.scale_modifier = s->scale_modifier;
    c.grid_x = (c.W + kTile - 1) / kTile;
    c.grid_y = (c.H + kTile - 1) / kTile;
    for (int i = 0; i < 3; ++i) c.bg[i] = s->bg[i];
    return c;
}

__device__ __forceinline__ void get_rect(float px, float py, int radius, int gx, int gy, int &x0, int &y0, int &x1,
                                         int &y1)
{
    float r = (float)radius;
    x0 = min(gx, max(0, (int)((px - r) / (float)kTile)));
    y0 = min(gy, max(0, (int)((py - r) / (float)kTile)));
    x1 = min(gx, max(0, (int)((px + r + (float)(kTile - 1)) / (float)kTile)));
    y1 = min(gy, max(0, (int)((py + r + (float)(kTile - 1)) / (float)kTile)));
}

// ---- decoupled look-back, warp-wide ------------------------------------------------------------
// Chained scan across CTA tiles: state[t] packs a 2-bit flag (0 empty, 1 aggregate, 2 inclusive
// prefix) with a 62-bit value in ONE 64-bit word, so no fence is needed between value and flag.
// Called by all 32 lanes of one warp; `tile` and `total` are warp-uniform.  Returns the exclusive
// prefix of `tile`.  The warp inspects 32 predecessors per round (one global round trip instead of
// 32 dependent ones), which matters when ~148 persistent CTAs publish at about the same time.
constexpr uint64_t kLbAggregate = 1ull << 62, kLbInclusive = 2ull << 62, kLbMask = (1ull << 62) - 1;

__device__ __forceinline__ uint64_t lookback_exclusive(unsigned long long *state, int tile, uint64_t total)
{
    volatile unsigned long long *st = state;
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) st[0] = kLbInclusive | total;
        return 0;
    }
    if (lane == 0) st[tile] = kLbAggregate | total;
    uint64_t excl = 0;
    for (int base = tile - 1;; base -= 32) {
        const int t = base - lane;
        uint64_t s = t >= 0 ? st[t] : kLbInclusive;  // a virtual inclusive zero precedes tile 0
        while (__any_sync(0xffffffffu, (s >> 62) == 0))
            if ((s >> 62) == 0) s = st[t];
        const uint32_t incl = __ballot_sync(0xffffffffu, (s >> 62) == 2ull);
        const int first = incl ? __ffs(incl) - 1 : 32;
        uint64_t v = lane <= first ? (s & kLbMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (incl) break;
    }
    if (lane == 0) st[tile] = kLbInclusive | (excl + total);
    return excl;
}

// The walk alone, for kernels that publish the aggregate of a tile EARLY (kLbAggregate | total, or kLbInclusive | total
// for tile 0, written by whoever knows the total first) and resolve the prefix LATER, when it is actually needed:
// by then the predecessors have long published, so nobody waits on another CTA's progress.
__device__ __forceinline__ uint64_t lookback_walk(unsigned long long *state, int tile, uint64_t total)
{
    volatile unsigned long long *st = state;
    const int lane = threadIdx.x & 31;
    if (tile == 0) return 0;
    uint64_t excl = 0;
    for (int base = tile - 1;; base -= 32) {
        const int t = base - lane;
        uint64_t s = t >= 0 ? st[t] : kLbInclusive;  // a virtual inclusive zero precedes tile 0
        while (__any_sync(0xffffffffu, (s >> 62) == 0))
            if ((s >> 62) == 0) s = st[t];
        const uint32_t incl = __ballot_sync(0xffffffffu, (s >> 62) == 2ull);
        const int first = incl ? __ffs(incl) - 1 : 32;
        uint64_t v = lane <= first ? (s & kLbMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (incl) break;
    }
    if (lane == 0) st[tile] = kLbInclusive | (excl + total);
    return excl;
}

// ---- sort / scan primitives (radix_sort.cu) -------------------------------------------
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

struct SortPlan {
    int npass;
    int64_t tiles_cap;
    size_t hist_off, lookback_off, ticket_off, zero_bytes, total_bytes;
};
SortPlan make_sort_plan(int64_t n_cap, int begin_bit, int end_bit);

// Enqueue a stable LSD sort.  `ws` must hold plan.total_bytes; its first plan.zero_bytes are
// expected to be ZERO on entry (sort_pairs zeroes them itself when zero_ws is true).
// The final pass writes to (keys_out, vals_out); passes ping-pong through (keys_tmp, vals_tmp).
int sort_pairs(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
               uint32_t *keys_tmp, uint32_t *vals_tmp, const uint32_t *n_dev, int64_t n_cap, int begin_bit,
               int end_bit, void *ws, bool zero_ws, cudaStream_t stream);

}  // namespace cgs

// Level division of the anchor set: for one level,
//     rows = round(anchor / voxel_size / level_scale)                    (scene/gaussian_model.py:1760)
//     sorted unique rows, inverse index, first (= minimum) source index  (utils/multi_level.py:3-31)
// The reference calls torch.unique(dim=0) (a thrust lexicographic row sort) twice per training
// iteration plus a float64 scatter-min.  Here the three rounded coordinates are packed into one
// order-preserving 63-bit key, sorted with the library's own onesweep radix sort (two 32-bit
// halves, stable LSD), and a single chained-scan kernel turns the sorted order into
// (inverse, first, count): no cub / thrust, no host synchronisation.
#include "common.cuh"

namespace cgs {
namespace lvd {
constexpr int kBias = 1 << 20;          // |round(coordinate)| must stay below 2^20
constexpr int kThreads = 256, kItems = 8, kTile = kThreads * kItems;

__global__ void __launch_bounds__(256)
voxel_keys_kernel(const float *__restrict__ pts, const uint8_t *__restrict__ keep, int n, float voxel, float scale,
                  uint32_t *__restrict__ key_lo, uint32_t *__restrict__ key_hi, int32_t *__restrict__ status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = 0;
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = pts[3 * (size_t)i + c];
        if (keep && !keep[i]) v = 0.0f;  // masked anchors collapse onto the origin voxel (:1758-1759)
        const float r = rintf(__fdiv_rn(__fdiv_rn(v, voxel), scale));
        bad |= !(fabsf(r) < (float)kBias);
        const int q = (int)r + kBias;
        key = (key << 21) | (uint64_t)(uint32_t)(q & 0x1fffff);
    }
    if (bad) atomicExch(status + 1, 1);
    key_lo[i] = (uint32_t)key;
    key_hi[i] = (uint32_t)(key >> 32);
}

__global__ void __launch_bounds__(256)
gather_u32_kernel(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, int n, uint32_t *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// order[i] = source index of the i-th smallest key (ties: ascending source index).  Group heads get
// consecutive ids; inverse[source] = id, first[id] = source of the head (= minimum source index).
__global__ void __launch_bounds__(kThreads)
unique_scan_kernel(const uint32_t *__restrict__ key_lo, const uint32_t *__restrict__ key_hi,
                   const uint32_t *__restrict__ order, int n, int32_t *__restrict__ inverse,
                   int32_t *__restrict__ first, unsigned long long *scan_state, uint32_t *ticket,
                   int32_t *__restrict__ status)
{
    __shared__ uint32_t s_tile, s_warp[kThreads / 32], s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base_i = tile * kTile + threadIdx.x * kItems;
    uint32_t src[kItems];
    uint32_t heads = 0;
    uint64_t prev = 0;
    if (base_i > 0 && base_i < n) {
        const uint32_t p = order[base_i - 1];
        prev = ((uint64_t)key_hi[p] << 32) | key_lo[p];
    }
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int i = base_i + k;
        src[k] = 0;
        if (i < n) {
            src[k] = order[i];
            const uint64_t key = ((uint64_t)key_hi[src[k]] << 32) | key_lo[src[k]];
            if (i == 0 || key != prev) heads |= 1u << k;
            prev = key;
        }
    }
    const uint32_t c = __popc(heads);
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wex = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        wex += w < warp ? s_warp[w] : 0u;
        total += s_warp[w];
    }
    if (warp == 0) {
        const uint64_t excl = lookback_exclusive(scan_state, tile, total);
        if (lane == 0) {
            s_base = (uint32_t)excl;
            if (tile == (n - 1) / kTile) status[0] = (int32_t)(excl + total);
        }
    }
    __syncthreads();
    uint32_t id = s_base + wex + incl - c;  // id of the first head of this thread; elements before it continue id - 1
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int i = base_i + k;
        if (i < n) {
            if (heads & (1u << k)) {
                first[id] = (int32_t)src[k];
                ++id;
            }
            inverse[src[k]] = (int32_t)(id - 1);
        }
    }
}

struct Plan {
    SortPlan sort;
    size_t sort_ws, scan_state, ticket, zero_bytes, key_lo, key_hi, hi_perm, keys_out, keys_tmp, vals_a, vals_b, vals_tmp,
        n_dev, total;
};

static Plan make_plan(int n)
{
    Plan p;
    const size_t N = (size_t)(n > 0 ? n : 1);
    p.sort = make_sort_plan((int64_t)N, 0, 32);
    size_t off = 0;
    p.sort_ws = off; off += 2 * p.sort.total_bytes;               // one zeroed workspace per sort
    p.scan_state = off; off += align_up(((N + kTile - 1) / kTile) * 8);
    p.ticket = off; off += align_up(16);
    p.zero_bytes = off;
    auto arr = [&](size_t &slot) { slot = off; off += align_up(N * 4); };
    arr(p.key_lo); arr(p.key_hi); arr(p.hi_perm); arr(p.keys_out); arr(p.keys_tmp); arr(p.vals_a); arr(p.vals_b);
    arr(p.vals_tmp);
    p.n_dev = off; off += align_up(16);
    p.total = off;
    return p;
}

__global__ void set_u32(uint32_t *p, uint32_t v) { *p = v; }
}  // namespace lvd
}  // namespace cgs

using namespace cgs;

extern "C" size_t cgs_unique_voxels_workspace_bytes(int n) { return lvd::make_plan(n).total; }

extern "C" int cgs_unique_voxels(const float *points, const uint8_t *keep, int n, float voxel_size, float level_scale,
                                 int32_t *inverse, int32_t *first, int32_t *status_dev, void *workspace,
                                 size_t workspace_bytes, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CGS_CHECK_PTR(status_dev);
    cudaMemsetAsync(status_dev, 0, 2 * sizeof(int32_t), st);
    if (n <= 0) return check_launch(__func__);
    CGS_CHECK_PTR(points); CGS_CHECK_PTR(inverse); CGS_CHECK_PTR(first); CGS_CHECK_PTR(workspace);
    if (!(voxel_size > 0.f) || !(level_scale > 0.f)) {
        set_error("%s: voxel_size and level_scale must be positive", __func__);
        return -2;
    }
    const lvd::Plan p = lvd::make_plan(n);
    if (workspace_bytes < p.total) {
        set_error("%s: workspace %zu < %zu bytes", __func__, workspace_bytes, p.total);
        return -3;
    }
    char *ws = static_cast<char *>(workspace);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(ws + off); };
    StageScope sc(ST_LEVEL_DIVIDE, st, 14);
    cudaMemsetAsync(ws, 0, p.zero_bytes, st);
    uint32_t *n_dev = u32(p.n_dev);
    lvd::set_u32<<<1, 1, 0, st>>>(n_dev, (uint32_t)n);
    const int grid = (n + 255) / 256;
    lvd::voxel_keys_kernel<<<grid, 256, 0, st>>>(points, keep, n, voxel_size, level_scale, u32(p.key_lo), u32(p.key_hi),
                                                 status_dev);
    // stable LSD over the 63-bit key: low word first (values = source index), then the high word
    if (int e = sort_pairs(u32(p.key_lo), nullptr, u32(p.keys_out), u32(p.vals_a), u32(p.keys_tmp), u32(p.vals_tmp), n_dev,
                           n, 0, 32, ws + p.sort_ws, false, st))
        return e;
    lvd::gather_u32_kernel<<<grid, 256, 0, st>>>(u32(p.key_hi), u32(p.vals_a), n, u32(p.hi_perm));
    if (int e = sort_pairs(u32(p.hi_perm), u32(p.vals_a), u32(p.keys_out), u32(p.vals_b), u32(p.keys_tmp), u32(p.vals_tmp),
                           n_dev, n, 0, 32, ws + p.sort_ws + p.sort.total_bytes, false, st))
        return e;
    lvd::unique_scan_kernel<<<(n + lvd::kTile - 1) / lvd::kTile, lvd::kThreads, 0, st>>>(
        u32(p.key_lo), u32(p.key_hi), u32(p.vals_b), n, inverse, first,
        reinterpret_cast<unsigned long long *>(ws + p.scan_state), u32(p.ticket), status_dev);
    return check_launch(__func__);
}

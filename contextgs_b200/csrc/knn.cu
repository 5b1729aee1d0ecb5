// Mean squared distance to the 3 nearest neighbours of every point: what `simple_knn._C.distCUDA2` returns
// (third-party, NOT in the reference tree; call sites scene/gaussian_model.py:389,407 inside `create_from_pcd`,
// where it sets the initial voxel size and the initial anchor scales).  Restated from the published behaviour of
// simple-knn (for every point: the three smallest squared distances to OTHER points, averaged); parity unpinned.
//
// simple-knn orders points along a Morton curve and prunes boxes.  Here: points are binned into a uniform grid
// (order-preserving 63-bit cell keys, the library's own radix sort), cell heads go into an open-addressing hash
// map (cell -> first sorted position), and one thread per point visits the shells of cells around its own cell
// until the third-best distance is provably inside the visited block.  The few points whose neighbourhood is
// still empty after kMaxRing shells (isolated outliers of an SfM cloud) are finished by a warp-wide scan over all
// points.  Exact (not approximate) nearest neighbours; -fmad=false so that a distance is (dx*dx + dy*dy) + dz*dz.
#include <cfloat>

#include "common.cuh"

namespace cgs {
namespace knn {
constexpr int kMaxRing = 3;
constexpr int kCoordMax = (1 << 21) - 1;
constexpr unsigned long long kEmpty = ~0ull;

struct Grid {
    float min[3];
    float h;
};

__device__ __forceinline__ unsigned long long pack(int x, int y, int z)
{
    return ((unsigned long long)(uint32_t)x << 42) | ((unsigned long long)(uint32_t)y << 21) | (unsigned long long)(uint32_t)z;
}

__device__ __forceinline__ void cell_of(float x, float y, float z, const Grid &g, int (&c)[3])
{
    const float p[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float f = floorf(__fdiv_rn(p[d] - g.min[d], g.h));
        c[d] = (int)fminf(fmaxf(f, 0.0f), (float)kCoordMax);
    }
}

__device__ __forceinline__ uint32_t slot_of(unsigned long long key, uint32_t mask)
{
    key ^= key >> 30; key *= 0xbf58476d1ce4e5b9ull;
    key ^= key >> 27; key *= 0x94d049bb133111ebull;
    key ^= key >> 31;
    return (uint32_t)key & mask;
}

__global__ void __launch_bounds__(256)
cell_keys_kernel(const float *__restrict__ pts, int n, Grid g, uint32_t *__restrict__ key_lo, uint32_t *__restrict__ key_hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c[3];
    cell_of(pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], g, c);
    const unsigned long long key = pack(c[0], c[1], c[2]);
    key_lo[i] = (uint32_t)key;
    key_hi[i] = (uint32_t)(key >> 32);
}

__global__ void __launch_bounds__(256)
gather_u32_kernel(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, int n, uint32_t *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// sorted position i <- point order[i]; the first point of every cell registers (cell -> i) in the hash map
__global__ void __launch_bounds__(256)
bin_points_kernel(const float *__restrict__ pts, const uint32_t *__restrict__ order, const uint32_t *__restrict__ key_lo,
                  const uint32_t *__restrict__ key_hi, int n, float4 *__restrict__ spos, unsigned long long *__restrict__ skey,
                  unsigned long long *table_keys, uint32_t *__restrict__ table_vals, uint32_t mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = order[i];
    const unsigned long long key = ((unsigned long long)key_hi[src] << 32) | key_lo[src];
    spos[i] = make_float4(pts[3 * (size_t)src], pts[3 * (size_t)src + 1], pts[3 * (size_t)src + 2], __uint_as_float(src));
    skey[i] = key;
    bool head = i == 0;
    if (!head) {
        const uint32_t p = order[i - 1];
        head = (((unsigned long long)key_hi[p] << 32) | key_lo[p]) != key;
    }
    if (head) {
        for (uint32_t h = slot_of(key, mask);; h = (h + 1) & mask) {
            if (atomicCAS(table_keys + h, kEmpty, key) == kEmpty) {   // cells are unique among heads
                table_vals[h] = (uint32_t)i;
                break;
            }
        }
    }
}

__device__ __forceinline__ void push_best(float d, float (&best)[3])
{
    if (d < best[2]) {
        if (d < best[1]) {
            best[2] = best[1];
            if (d < best[0]) {
                best[1] = best[0];
                best[0] = d;
            } else {
                best[1] = d;
            }
        } else {
            best[2] = d;
        }
    }
}

__device__ __forceinline__ float dist2(const float4 &a, const float4 &b)
{
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;
}

__global__ void __launch_bounds__(128)
knn3_grid_kernel(const float4 *__restrict__ spos, const unsigned long long *__restrict__ skey, int n, Grid g,
                 const unsigned long long *__restrict__ table_keys, const uint32_t *__restrict__ table_vals, uint32_t mask,
                 float *__restrict__ out, int32_t *__restrict__ outliers, int32_t *__restrict__ n_outliers)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = spos[i];
    const unsigned long long own = skey[i];
    const int cx = (int)(own >> 42), cy = (int)((own >> 21) & kCoordMax), cz = (int)(own & kCoordMax);
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    bool closed = false;
    for (int r = 0; r <= kMaxRing && !closed; ++r) {
        for (int dz = -r; dz <= r; ++dz)
            for (int dy = -r; dy <= r; ++dy)
                for (int dx = -r; dx <= r; ++dx) {
                    if (max(max(abs(dx), abs(dy)), abs(dz)) != r) continue;     // shell r only
                    const int x = cx + dx, y = cy + dy, z = cz + dz;
                    if ((x | y | z) < 0 || x > kCoordMax || y > kCoordMax || z > kCoordMax) continue;
                    const unsigned long long key = pack(x, y, z);
                    uint32_t start = 0xffffffffu;
                    for (uint32_t h = slot_of(key, mask);; h = (h + 1) & mask) {
                        const unsigned long long t = table_keys[h];
                        if (t == key) {
                            start = table_vals[h];
                            break;
                        }
                        if (t == kEmpty) break;
                    }
                    if (start == 0xffffffffu) continue;
                    for (int j = (int)start; j < n && skey[j] == key; ++j)
                        if (j != i) push_best(dist2(p, spos[j]), best);
                }
        // everything outside the (2r+1)^3 block is at least r*h away (0.999: the cell assignment rounds in fp32)
        const float reach = (float)r * g.h;
        closed = best[2] <= reach * reach * 0.999f;
    }
    if (closed) {
        out[__float_as_uint(p.w)] = (best[0] + best[1] + best[2]) / 3.0f;
    } else {
        outliers[atomicAdd(n_outliers, 1)] = i;
    }
}

// One warp per unresolved point: all points, 32 at a time; then the three smallest of the 96 lane-local candidates.
__global__ void __launch_bounds__(256)
knn3_scan_kernel(const float4 *__restrict__ spos, int n, const int32_t *__restrict__ outliers,
                 const int32_t *__restrict__ n_outliers, int max_outliers, float *__restrict__ out,
                 int32_t *__restrict__ status)
{
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int total = *n_outliers;
    if (total > max_outliers) {      // the cell size does not fit the data: O(n) scans for this many points would take
        if (blockIdx.x == 0 && threadIdx.x == 0) status[1] = 1;      // seconds -- report, the caller retries coarser
        return;
    }
    for (int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; o < total; o += warps) {
        const int i = outliers[o];
        const float4 p = spos[i];
        float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
        for (int j = lane; j < n; j += 32)
            if (j != i) push_best(dist2(p, spos[j]), best);
        float top[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float m = best[0];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, s));
            top[k] = m;
            const uint32_t owners = __ballot_sync(0xffffffffu, best[0] == m);
            if (lane == __ffs(owners) - 1) {         // exactly one lane pops its head
                best[0] = best[1];
                best[1] = best[2];
                best[2] = FLT_MAX;
            }
        }
        if (lane == 0) out[__float_as_uint(p.w)] = (top[0] + top[1] + top[2]) / 3.0f;
    }
}

struct Plan {
    SortPlan sort;
    uint32_t table_slots;
    size_t sort_ws, n_dev, counter, zero_bytes, table_keys, table_vals, key_lo, key_hi, hi_perm, keys_out, keys_tmp, vals_a,
        vals_b, vals_tmp, outliers, skey, spos, total;
};

static Plan make_plan(int n)
{
    Plan p;
    const size_t N = (size_t)(n > 0 ? n : 1);
    p.sort = make_sort_plan((int64_t)N, 0, 32);
    p.table_slots = 1024;
    while ((size_t)p.table_slots < 2 * N) p.table_slots <<= 1;
    size_t off = 0;
    p.sort_ws = off; off += 2 * p.sort.total_bytes;
    p.n_dev = off; off += align_up(16);
    p.counter = off; off += align_up(16);
    p.zero_bytes = off;
    p.table_keys = off; off += align_up((size_t)p.table_slots * 8);
    p.table_vals = off; off += align_up((size_t)p.table_slots * 4);
    auto arr = [&](size_t &slot) { slot = off; off += align_up(N * 4); };
    arr(p.key_lo); arr(p.key_hi); arr(p.hi_perm); arr(p.keys_out); arr(p.keys_tmp); arr(p.vals_a); arr(p.vals_b);
    arr(p.vals_tmp); arr(p.outliers);
    p.skey = off; off += align_up(N * 8);
    p.spos = off; off += align_up(N * 16);
    p.total = off;
    return p;
}

__global__ void set_u32(uint32_t *p, uint32_t v) { *p = v; }
}  // namespace knn
}  // namespace cgs

using namespace cgs;

extern "C" size_t cgs_knn3_workspace_bytes(int n) { return knn::make_plan(n).total; }

extern "C" int cgs_knn3_mean_dist2(const float *points, int n, const float *bbox_min_host, float cell, float *mean_dist2,
                                   int32_t *status_dev, void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CGS_CHECK_PTR(status_dev);
    cudaMemsetAsync(status_dev, 0, 2 * sizeof(int32_t), st);
    if (n <= 0) return check_launch(__func__);
    CGS_CHECK_PTR(points); CGS_CHECK_PTR(bbox_min_host); CGS_CHECK_PTR(mean_dist2); CGS_CHECK_PTR(workspace);
    if (!(cell > 0.f)) {
        set_error("%s: cell must be positive", __func__);
        return -2;
    }
    const knn::Plan p = knn::make_plan(n);
    if (workspace_bytes < p.total) {
        set_error("%s: workspace %zu < %zu bytes", __func__, workspace_bytes, p.total);
        return -3;
    }
    knn::Grid g;
    for (int d = 0; d < 3; ++d) g.min[d] = bbox_min_host[d];
    g.h = cell;
    char *ws = static_cast<char *>(workspace);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(ws + off); };
    StageScope sc(ST_DENSIFY, st, 15);
    cudaMemsetAsync(ws, 0, p.zero_bytes, st);
    cudaMemsetAsync(ws + p.table_keys, 0xff, (size_t)p.table_slots * 8, st);
    uint32_t *n_dev = u32(p.n_dev);
    knn::set_u32<<<1, 1, 0, st>>>(n_dev, (uint32_t)n);
    const int grid = (n + 255) / 256;
    knn::cell_keys_kernel<<<grid, 256, 0, st>>>(points, n, g, u32(p.key_lo), u32(p.key_hi));
    if (int e = sort_pairs(u32(p.key_lo), nullptr, u32(p.keys_out), u32(p.vals_a), u32(p.keys_tmp), u32(p.vals_tmp), n_dev,
                           n, 0, 32, ws + p.sort_ws, false, st))
        return e;
    knn::gather_u32_kernel<<<grid, 256, 0, st>>>(u32(p.key_hi), u32(p.vals_a), n, u32(p.hi_perm));
    if (int e = sort_pairs(u32(p.hi_perm), u32(p.vals_a), u32(p.keys_out), u32(p.vals_b), u32(p.keys_tmp), u32(p.vals_tmp),
                           n_dev, n, 0, 32, ws + p.sort_ws + p.sort.total_bytes, false, st))
        return e;
    float4 *spos = reinterpret_cast<float4 *>(ws + p.spos);
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(ws + p.skey);
    unsigned long long *tkeys = reinterpret_cast<unsigned long long *>(ws + p.table_keys);
    const uint32_t mask = p.table_slots - 1;
    knn::bin_points_kernel<<<grid, 256, 0, st>>>(points, u32(p.vals_b), u32(p.key_lo), u32(p.key_hi), n, spos, skey, tkeys,
                                                 u32(p.table_vals), mask);
    int32_t *outliers = reinterpret_cast<int32_t *>(ws + p.outliers);
    int32_t *counter = reinterpret_cast<int32_t *>(ws + p.counter);
    knn::knn3_grid_kernel<<<(n + 127) / 128, 128, 0, st>>>(spos, skey, n, g, tkeys, u32(p.table_vals), mask, mean_dist2,
                                                           outliers, counter);
    const int max_outliers = n / 32 > 4096 ? n / 32 : 4096;
    knn::knn3_scan_kernel<<<148 * 4, 256, 0, st>>>(spos, n, outliers, counter, max_outliers, mean_dist2, status_dev);
    cudaMemcpyAsync(status_dev, counter, sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
    return check_launch(__func__);
}

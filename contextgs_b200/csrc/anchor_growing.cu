// One depth of anchor growing (scene/gaussian_model.py:778-816 inside `anchor_growing`):
//     all_xyz   = get_anchor[:,None] + _offset * get_scaling[:, :3][:,None]                       (:778)
//     cells     = round(all_xyz[candidate] / cur_size).int()                                       (:785-786)
//     unique    = lexicographically sorted unique cells, inverse                                   (:788)
//     fresh     = unique cells not occupied by round(get_anchor / cur_size).int()                  (:783, :790-802)
//     new_anchor = fresh cells * cur_size;  new_feat / new_hyper = per-cell maximum over the       (:803, :812-816)
//                  candidates of the cell of the SOURCE anchor's feature / hyper-latent row
// The reference tests occupancy with a chunked all-pairs comparison, O(|unique| x N) (seconds at 1.5 M anchors),
// and needs torch_scatter for the maximum.  Here: the existing anchors' cells go into an open-addressing hash set,
// the candidate cells are packed into order-preserving 63-bit keys and sorted with the library's own radix sort,
// one chained-scan pass marks group heads, probes the set and numbers the fresh cells, and one warp per fresh cell
// walks its (contiguous) group for the maximum.  No host synchronisation, no cub / thrust.
// Compiled with -fmad=false: `anchor + offset * scaling` is a multiply and an add in the reference.
#include "common.cuh"

extern "C" size_t cgs_compact_workspace_bytes(int N);
extern "C" int cgs_compact_indices(const uint8_t *mask, int N, int32_t *out_idx, int32_t *count_dev, void *workspace,
                                   size_t workspace_bytes, void *stream);

namespace cgs {
namespace grow {
constexpr int kBias = 1 << 20;                      // |cell coordinate| must stay below 2^20 (same packing as level_divide.cu)
constexpr unsigned long long kEmpty = ~0ull;        // never a valid key: keys use 63 bits
constexpr int kThreads = 256, kItems = 4, kTile = kThreads * kItems;

__device__ __forceinline__ bool cell_key(float x, float y, float z, float cs, unsigned long long &key)
{
    const float r[3] = {rintf(__fdiv_rn(x, cs)), rintf(__fdiv_rn(y, cs)), rintf(__fdiv_rn(z, cs))};
    bool ok = true;
    key = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ok &= fabsf(r[c]) < (float)kBias;
        key = (key << 21) | (unsigned long long)(uint32_t)(((int)r[c] + kBias) & 0x1fffff);
    }
    return ok;
}

__device__ __forceinline__ uint32_t slot_of(unsigned long long key, uint32_t mask)
{
    key ^= key >> 30; key *= 0xbf58476d1ce4e5b9ull;
    key ^= key >> 27; key *= 0x94d049bb133111ebull;
    key ^= key >> 31;
    return (uint32_t)key & mask;
}

__global__ void __launch_bounds__(256)
occupy_kernel(const float *__restrict__ anchor_q, int n, float cs, unsigned long long *table, uint32_t mask,
              int32_t *__restrict__ status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long key;
    if (!cell_key(anchor_q[3 * (size_t)i], anchor_q[3 * (size_t)i + 1], anchor_q[3 * (size_t)i + 2], cs, key))
        atomicExch(status + 1, 1);
    for (uint32_t h = slot_of(key, mask);; h = (h + 1) & mask) {
        const unsigned long long prev = atomicCAS(table + h, kEmpty, key);
        if (prev == kEmpty || prev == key) break;
    }
}

__device__ __forceinline__ bool occupied(const unsigned long long *__restrict__ table, uint32_t mask, unsigned long long key)
{
    for (uint32_t h = slot_of(key, mask);; h = (h + 1) & mask) {
        const unsigned long long t = table[h];
        if (t == key) return true;
        if (t == kEmpty) return false;
    }
}

__global__ void __launch_bounds__(256)
candidate_keys_kernel(const float *__restrict__ anchor_q, const float *__restrict__ offset, const float *__restrict__ scaling,
                      int scaling_stride, const int32_t *__restrict__ cand_slot, const int32_t *__restrict__ n_dev, int K,
                      float cs, uint32_t *__restrict__ key_lo, uint32_t *__restrict__ key_hi, int32_t *__restrict__ status)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= *n_dev) return;
    const int slot = cand_slot[c], a = slot / K;
    float p[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
        p[d] = anchor_q[3 * (size_t)a + d] + offset[3 * (size_t)slot + d] * scaling[(size_t)a * scaling_stride + d];
    unsigned long long key;
    if (!cell_key(p[0], p[1], p[2], cs, key)) atomicExch(status + 1, 1);
    key_lo[c] = (uint32_t)key;
    key_hi[c] = (uint32_t)(key >> 32);
}

// more candidates than the caller sized the workspace for: flag it and keep the first cand_cap (the host raises)
__global__ void clamp_count_kernel(int32_t *count, int cand_cap, int32_t *status)
{
    if (*count > cand_cap) {
        status[4] = 2;
        *count = cand_cap;
    }
}

__global__ void __launch_bounds__(256)
gather_hi_kernel(const uint32_t *__restrict__ key_hi, const uint32_t *__restrict__ idx, const int32_t *__restrict__ n_dev,
                 uint32_t *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *n_dev) dst[i] = key_hi[idx[i]];
}

// order[i] = candidate of the i-th smallest cell.  A group head that is not occupied is a FRESH cell; fresh cells are
// numbered in sorted order.  The chained scan carries (#heads | #fresh << 31) in one 62-bit value.
__global__ void __launch_bounds__(kThreads)
fresh_scan_kernel(const uint32_t *__restrict__ key_lo, const uint32_t *__restrict__ key_hi, const uint32_t *__restrict__ order,
                  const int32_t *__restrict__ n_dev, const unsigned long long *__restrict__ table, uint32_t mask, float cs,
                  int new_cap, int32_t *__restrict__ nid_sorted, float *__restrict__ new_anchor,
                  unsigned long long *scan_state, uint32_t *ticket, int32_t *__restrict__ status)
{
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_warp[kThreads / 32], s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile, n = *n_dev;
    if ((long long)tile * kTile >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base_i = tile * kTile + threadIdx.x * kItems;
    unsigned long long keys[kItems], prev = kEmpty;
    uint32_t heads = 0, fresh = 0;
    if (base_i > 0 && base_i < n) {
        const uint32_t p = order[base_i - 1];
        prev = ((unsigned long long)key_hi[p] << 32) | key_lo[p];
    }
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int i = base_i + k;
        keys[k] = kEmpty;
        if (i < n) {
            const uint32_t src = order[i];
            keys[k] = ((unsigned long long)key_hi[src] << 32) | key_lo[src];
            if (i == 0 || keys[k] != prev) {
                heads |= 1u << k;
                if (!occupied(table, mask, keys[k])) fresh |= 1u << k;
            }
            prev = keys[k];
        }
    }
    const unsigned long long mine = (unsigned long long)__popc(heads) | ((unsigned long long)__popc(fresh) << 31);
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long wex = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        wex += w < warp ? s_warp[w] : 0ull;
        total += s_warp[w];
    }
    if (warp == 0) {
        const unsigned long long excl = lookback_exclusive(scan_state, tile, total);
        if (lane == 0) {
            s_base = excl;
            if (tile == (n - 1) / kTile) {
                const unsigned long long all = excl + total;
                const int n_new = (int)(all >> 31);
                status[0] = n_new;
                status[2] = n;
                status[3] = (int)(all & 0x7fffffffu);
                if (n_new > new_cap) atomicOr(status + 4, 1);
            }
        }
    }
    __syncthreads();
    int id = (int)((s_base + wex + incl - mine) >> 31);          // fresh cells before this thread's first item
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int i = base_i + k;
        if (i >= n) break;
        int nid = -1;
        if (fresh & (1u << k)) {
            nid = id++;
            if (nid < new_cap) {
                // selected_grid_coords_unique[remove_duplicates] * cur_size (:803): int32 -> float32, one multiply
                new_anchor[3 * (size_t)nid + 0] = (float)((int)((keys[k] >> 42) & 0x1fffff) - kBias) * cs;
                new_anchor[3 * (size_t)nid + 1] = (float)((int)((keys[k] >> 21) & 0x1fffff) - kBias) * cs;
                new_anchor[3 * (size_t)nid + 2] = (float)((int)(keys[k] & 0x1fffff) - kBias) * cs;
            }
        }
        nid_sorted[i] = nid;
    }
}

// scatter_max over the group of every fresh cell (:812-816): one warp per sorted position, lanes over the channels
// [feat | hyper]; only fresh group heads do work and a group is contiguous in the sorted order.
__global__ void __launch_bounds__(256)
group_max_kernel(const uint32_t *__restrict__ key_lo, const uint32_t *__restrict__ key_hi, const uint32_t *__restrict__ order,
                 const int32_t *__restrict__ n_dev, const int32_t *__restrict__ nid_sorted, const int32_t *__restrict__ cand_slot,
                 int K, const float *__restrict__ feat, int feat_dim, const float *__restrict__ hyper, int hyper_dim,
                 int new_cap, float *__restrict__ new_feat, float *__restrict__ new_hyper)
{
    const int i = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31, n = *n_dev;
    if (i >= n) return;
    const int nid = nid_sorted[i];
    if (nid < 0 || nid >= new_cap) return;
    const uint32_t lo = key_lo[order[i]], hi = key_hi[order[i]];
    for (int c0 = 0; c0 < feat_dim + hyper_dim; c0 += 32) {
        const int c = c0 + lane;
        const bool live = c < feat_dim + hyper_dim;
        float m = -INFINITY;
        for (int j = i; j < n; ++j) {
            const uint32_t src = order[j];
            if (j > i && (key_lo[src] != lo || key_hi[src] != hi)) break;
            const int a = cand_slot[src] / K;
            if (live) m = fmaxf(m, c < feat_dim ? feat[(size_t)a * feat_dim + c] : hyper[(size_t)a * hyper_dim + (c - feat_dim)]);
        }
        if (live) {
            if (c < feat_dim) new_feat[(size_t)nid * feat_dim + c] = m;
            else new_hyper[(size_t)nid * hyper_dim + (c - feat_dim)] = m;
        }
    }
}

struct Plan {
    SortPlan sort;
    uint32_t table_slots;
    size_t sort_ws, scan_state, ticket, count, zero_bytes, table, compact_ws, cand_slot, key_lo, key_hi, hi_perm, keys_out,
        keys_tmp, vals_a, vals_b, vals_tmp, nid, total;
};

static Plan make_plan(int n_anchors, int n_offsets, int cand_cap)
{
    Plan p;
    const size_t M = (size_t)(cand_cap > 0 ? cand_cap : 1), N = (size_t)(n_anchors > 0 ? n_anchors : 1);
    p.sort = make_sort_plan((int64_t)M, 0, 32);
    p.table_slots = 1024;
    while ((size_t)p.table_slots < 2 * N) p.table_slots <<= 1;
    size_t off = 0;
    p.sort_ws = off; off += 2 * p.sort.total_bytes;
    p.scan_state = off; off += align_up(((M + kTile - 1) / kTile) * 8);
    p.ticket = off; off += align_up(16);
    p.count = off; off += align_up(16);
    p.zero_bytes = off;
    p.table = off; off += align_up((size_t)p.table_slots * 8);
    p.compact_ws = off; off += align_up(cgs_compact_workspace_bytes((int)(N * (size_t)n_offsets)));
    p.cand_slot = off; off += align_up(N * (size_t)n_offsets * 4);      // the compaction may emit every slot
    auto arr = [&](size_t &slot) { slot = off; off += align_up(M * 4); };
    arr(p.key_lo); arr(p.key_hi); arr(p.hi_perm); arr(p.keys_out); arr(p.keys_tmp); arr(p.vals_a);
    arr(p.vals_b); arr(p.vals_tmp); arr(p.nid);
    p.total = off;
    return p;
}
}  // namespace grow
}  // namespace cgs

using namespace cgs;

extern "C" size_t cgs_anchor_growing_workspace_bytes(int n_anchors, int n_offsets, int cand_cap)
{
    return grow::make_plan(n_anchors, n_offsets, cand_cap).total;
}

extern "C" int cgs_anchor_growing(const float *anchor_q, const float *offset, const float *scaling, int scaling_stride,
                                  const float *feat, int feat_dim, const float *hyper, int hyper_dim,
                                  const uint8_t *candidate, int n_anchors, int n_offsets, float cur_size, int cand_cap,
                                  float *new_anchor, float *new_feat, float *new_hyper, int new_cap, int32_t *status_dev,
                                  void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CGS_CHECK_PTR(status_dev);
    cudaMemsetAsync(status_dev, 0, 5 * sizeof(int32_t), st);
    if (n_anchors <= 0 || cand_cap <= 0) return check_launch(__func__);
    CGS_CHECK_PTR(anchor_q); CGS_CHECK_PTR(offset); CGS_CHECK_PTR(scaling); CGS_CHECK_PTR(feat); CGS_CHECK_PTR(hyper);
    CGS_CHECK_PTR(candidate); CGS_CHECK_PTR(new_anchor); CGS_CHECK_PTR(new_feat); CGS_CHECK_PTR(new_hyper);
    CGS_CHECK_PTR(workspace);
    if (!(cur_size > 0.f) || n_offsets <= 0 || feat_dim <= 0 || hyper_dim <= 0 || scaling_stride < 3 || new_cap <= 0) {
        set_error("%s: cur_size, n_offsets, feat_dim, hyper_dim, new_cap must be positive and scaling_stride >= 3", __func__);
        return -2;
    }
    if ((int64_t)n_anchors * n_offsets > INT32_MAX) {
        set_error("%s: n_anchors * n_offsets exceeds 2^31 - 1", __func__);
        return -2;
    }
    const grow::Plan p = grow::make_plan(n_anchors, n_offsets, cand_cap);
    if (workspace_bytes < p.total) {
        set_error("%s: workspace %zu < %zu bytes", __func__, workspace_bytes, p.total);
        return -3;
    }
    char *ws = static_cast<char *>(workspace);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(ws + off); };
    auto i32 = [&](size_t off) { return reinterpret_cast<int32_t *>(ws + off); };
    cudaMemsetAsync(ws, 0, p.zero_bytes, st);
    cudaMemsetAsync(ws + p.table, 0xff, (size_t)p.table_slots * 8, st);
    // candidates in slot order (cand_cap must cover them: the caller knows the count or passes n_anchors * n_offsets)
    int32_t *count = i32(p.count);
    const int n_slots = n_anchors * n_offsets;
    if (int e = cgs_compact_indices(candidate, n_slots, i32(p.cand_slot), count, ws + p.compact_ws,
                                    cgs_compact_workspace_bytes(n_slots), stream))
        return e;
    StageScope sc(ST_DENSIFY, st, 14);
    grow::clamp_count_kernel<<<1, 1, 0, st>>>(count, cand_cap, status_dev);
    unsigned long long *table = reinterpret_cast<unsigned long long *>(ws + p.table);
    const uint32_t mask = p.table_slots - 1;
    grow::occupy_kernel<<<(n_anchors + 255) / 256, 256, 0, st>>>(anchor_q, n_anchors, cur_size, table, mask, status_dev);
    const int grid = (cand_cap + 255) / 256;
    grow::candidate_keys_kernel<<<grid, 256, 0, st>>>(anchor_q, offset, scaling, scaling_stride, i32(p.cand_slot), count,
                                                      n_offsets, cur_size, u32(p.key_lo), u32(p.key_hi), status_dev);
    const uint32_t *n_dev = reinterpret_cast<const uint32_t *>(count);
    if (int e = sort_pairs(u32(p.key_lo), nullptr, u32(p.keys_out), u32(p.vals_a), u32(p.keys_tmp), u32(p.vals_tmp), n_dev,
                           cand_cap, 0, 32, ws + p.sort_ws, false, st))
        return e;
    grow::gather_hi_kernel<<<grid, 256, 0, st>>>(u32(p.key_hi), u32(p.vals_a), count, u32(p.hi_perm));
    if (int e = sort_pairs(u32(p.hi_perm), u32(p.vals_a), u32(p.keys_out), u32(p.vals_b), u32(p.keys_tmp), u32(p.vals_tmp),
                           n_dev, cand_cap, 0, 32, ws + p.sort_ws + p.sort.total_bytes, false, st))
        return e;
    grow::fresh_scan_kernel<<<(cand_cap + grow::kTile - 1) / grow::kTile, grow::kThreads, 0, st>>>(
        u32(p.key_lo), u32(p.key_hi), u32(p.vals_b), count, table, mask, cur_size, new_cap, i32(p.nid), new_anchor,
        reinterpret_cast<unsigned long long *>(ws + p.scan_state), u32(p.ticket), status_dev);
    grow::group_max_kernel<<<(int)(((size_t)cand_cap * 32 + 255) / 256), 256, 0, st>>>(
        u32(p.key_lo), u32(p.key_hi), u32(p.vals_b), count, i32(p.nid), i32(p.cand_slot), n_offsets, feat, feat_dim, hyper,
        hyper_dim, new_cap, new_feat, new_hyper);
    return check_launch(__func__);
}

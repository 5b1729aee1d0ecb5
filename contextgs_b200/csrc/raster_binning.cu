// Tile binning without a 64-bit sort, without a host read-back -- and without ever sorting the
// R (Gaussian, tile) instances.
//
// Upstream (rasterizer_impl.cu, not in the reference tree) emits one 64-bit key
// (tile<<32 | depth_bits) per (Gaussian, tile) instance and runs a 6-pass cub radix sort over all
// R instances, after a blocking D2H copy of R.  The order it produces is (tile, depth, id).
// Here the same point_list / ranges are produced, bit for bit, in two levels:
//
//   1. 4-pass radix sort of the P Gaussians by depth bits (stable -> ties keep ascending id);
//   2. one chained scan in that depth order which also EMITS one (super-tile, id) pair per covered
//      SUPER-TILE (8 x 4 tiles): ~1.4 pairs per Gaussian instead of ~4.8 instances, R stays on device;
//   3. stable radix sort of the pairs by super-tile id (one 8-bit pass up to 255 super-tiles, i.e.
//      1920x1080) -> per super-tile a depth-ordered list of the Gaussians touching it;
//   4. expansion: a Gaussian's footprint inside its super-tile is a 32-bit mask (one bit per tile), so
//      the stable rank of an instance within its tile is a popcount of a warp ballot:
//        bin_count   per 256-pair chunk, per tile bin: how many pairs of the chunk's last super-tile
//                    cover it (+ atomics into the per-tile totals for the few pairs of other super-tiles)
//        bin_runs    per (super-tile, bin): running sum over the chunks of that super-tile
//        tile_starts exclusive scan of the per-tile totals -> ranges
//        bin_write   point_list[start(tile) + chunks before + rank inside the chunk] = id
//
// Ranking costs ~3 instructions per instance here against ~170 per key and pass in the radix sort
// (which is issue bound on B200, not bandwidth bound), and the R-sized arrays are written exactly once.
#include "common.cuh"

namespace cgs {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

constexpr int kSuperX = CGS_SUPER_X, kSuperY = CGS_SUPER_Y, kBins = kSuperX * kSuperY;  // 8 x 4 tiles = 32 bins
static_assert(kBins == 32, "one bin per lane");
constexpr int kChunk = 256;             // pairs per chunk (one CTA iteration)
constexpr int kChunkWarps = kChunk / 32;

struct Rect {
    int x0, y0, x1, y1;
};
__device__ __forceinline__ Rect unpack_rect(uint2 r)
{
    Rect q;
    q.x0 = (int)(r.x & 0xffffu); q.x1 = (int)(r.x >> 16);
    q.y0 = (int)(r.y & 0xffffu); q.y1 = (int)(r.y >> 16);
    return q;
}

// footprint of a tile rectangle inside super-tile (sx, sy): bit (ly * 8 + lx)
__device__ __forceinline__ uint32_t supertile_mask(const Rect q, int sx, int sy)
{
    const int lx0 = max(q.x0 - kSuperX * sx, 0), lx1 = min(q.x1 - kSuperX * sx, kSuperX);
    const int ly0 = max(q.y0 - kSuperY * sy, 0), ly1 = min(q.y1 - kSuperY * sy, kSuperY);
    if (lx1 <= lx0 || ly1 <= ly0) return 0u;
    const uint32_t row = ((1u << (lx1 - lx0)) - 1u) << lx0;                        // 8 bits
    const uint32_t rows = (0xffffffffu >> (32 - 8 * (ly1 - ly0))) << (8 * ly0);    // 1..4 byte rows
    return (row * 0x01010101u) & rows;
}

// Chained scan over the Gaussians in depth order of the number of super-tiles each one touches; emits the
// (super-tile, id) pairs at their scanned offsets.  The instance count R (sum of tiles touched) needs no
// prefix, only a total: every CTA adds its share to a 64-bit accumulator and the last CTA to finish
// publishes R, the overflow flag and the pair count.
__global__ void __launch_bounds__(kScanThreads)
scan_emit_pairs_kernel(const uint32_t *__restrict__ order, const uint2 *__restrict__ rects,
                       const uint32_t *__restrict__ p_dev, int64_t R_cap, int sgx, uint32_t *__restrict__ pair_keys,
                       uint32_t *__restrict__ pair_vals, unsigned long long *state, uint32_t *ticket,
                       unsigned long long *r_acc, uint32_t *done, uint32_t *__restrict__ n_pairs_dev,
                       int32_t *__restrict__ status)
{
    const int P = (int)*p_dev;  // device-side Gaussian count (the grid is sized by the host's capacity)
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_warp[kScanThreads / 32], s_warp_tiles[kScanThreads / 32];
    __shared__ uint64_t s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (P <= 0) {  // nothing to draw: publish R = 0 once
        if (tile == 0 && threadIdx.x == 0) {
            status[CGS_STATUS_NUM_RENDERED] = 0;
            status[CGS_STATUS_OVERFLOW] = 0;
            status[CGS_STATUS_NUM_SORTED] = 0;
            *n_pairs_dev = 0;
        }
        return;
    }
    if ((int64_t)tile * kScanTile >= (int64_t)P) return;  // tiles beyond the device-side count
    const uint32_t num_tiles = (uint32_t)(((int64_t)P + kScanTile - 1) / kScanTile);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];      // inclusive pair counts inside the thread
    uint32_t gid[kScanItems];
    uint2 rr[kScanItems];   // packed rectangles (unpacked on use: registers decide this kernel's occupancy)
    uint32_t local = 0;
    uint64_t tiles_sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int idx = base + i;
        gid[i] = idx < P ? order[idx] : 0u;
    }
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int idx = base + i;
        rr[i] = idx < P ? rects[gid[i]] : make_uint2(0u, 0u);
        const Rect q = unpack_rect(rr[i]);
        const uint32_t t = (uint32_t)((q.x1 - q.x0) * (q.y1 - q.y0));
        uint32_t ns = 0;
        if (t) ns = (uint32_t)(((q.x1 - 1) / kSuperX - q.x0 / kSuperX + 1) * ((q.y1 - 1) / kSuperY - q.y0 / kSuperY + 1));
        tiles_sum += t;
        local += ns;
        v[i] = local;
    }
    uint64_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tiles_sum += __shfl_xor_sync(0xffffffffu, tiles_sum, o);
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_warp_tiles[warp] = tiles_sum;
    __syncthreads();
    uint64_t warp_excl = 0, total = 0, tiles_total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        warp_excl += w < warp ? s_warp[w] : 0ull;
        total += s_warp[w];
        tiles_total += s_warp_tiles[w];
    }
    if (warp == 0) {
        const uint64_t excl = lookback_exclusive(state, (int)tile, total);
        if (lane == 0) {
            s_prefix = excl;
            atomicAdd(r_acc, (unsigned long long)tiles_total);
            __threadfence();
            if (atomicAdd(done, 1u) == num_tiles - 1) {   // every CTA has contributed
                __threadfence();
                const uint64_t R = atomicAdd(r_acc, 0ull);
                const bool over = R > (uint64_t)R_cap;
                status[CGS_STATUS_NUM_RENDERED] = (int32_t)(R > 0x7fffffffull ? 0x7fffffff : R);
                status[CGS_STATUS_OVERFLOW] = over ? 1 : 0;
                status[CGS_STATUS_NUM_SORTED] = (int32_t)(over ? (uint64_t)R_cap : R);
            }
            if (tile == num_tiles - 1) {   // pairs <= R: only clipped on an instance overflow
                const uint64_t np = excl + total;
                *n_pairs_dev = (uint32_t)(np > (uint64_t)R_cap ? (uint64_t)R_cap : np);
                status[CGS_STATUS_NUM_PAIRS] = (int32_t)(np > 0x7fffffffull ? 0x7fffffff : np);
            }
        }
    }
    __syncthreads();
    const uint64_t thread_excl = s_prefix + warp_excl + incl - local;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const uint32_t prev = i ? v[i - 1] : 0u;
        if (v[i] == prev) continue;  // no tiles
        int64_t off = (int64_t)(thread_excl + prev);
        const Rect q = unpack_rect(rr[i]);
        const int sx0 = q.x0 / kSuperX, sx1 = (q.x1 - 1) / kSuperX;
        const int sy0 = q.y0 / kSuperY, sy1 = (q.y1 - 1) / kSuperY;
        for (int sy = sy0; sy <= sy1; ++sy)
            for (int sx = sx0; sx <= sx1; ++sx) {
                if (off < R_cap) {
                    pair_keys[off] = (uint32_t)(sy * sgx + sx);
                    pair_vals[off] = gid[i];
                }
                ++off;
            }
    }
}

// ---- expansion -------------------------------------------------------------------------------------
struct PairLane {
    bool valid;
    uint32_t st, gid, mask;
};

// Two-deep software pipeline over the chunks a CTA visits: (key, id) of chunk i+2 and the rectangle of chunk
// i+1 (a dependent gather) are in flight while chunk i is ranked.
struct PairStream {
    const uint32_t *keys, *vals;
    const uint2 *rects;
    uint32_t n, stride;
    uint32_t k1, v1, k2, v2;
    uint2 r1;
    __device__ __forceinline__ void fetch_kv(uint32_t e, uint32_t &k, uint32_t &v) const
    {
        k = 0xffffffffu; v = 0;
        if (e < n) { k = keys[e]; v = vals[e]; }
    }
    __device__ __forceinline__ void start(uint32_t e0)
    {
        fetch_kv(e0, k1, v1);
        r1 = k1 != 0xffffffffu ? rects[v1] : make_uint2(0u, 0u);
        fetch_kv(e0 + stride, k2, v2);
    }
    // returns the lane's pair of the chunk whose first element of this lane is `e`, and advances
    __device__ __forceinline__ PairLane next(uint32_t e, int sgx)
    {
        PairLane p;
        p.valid = e < n;
        p.st = k1; p.gid = v1; p.mask = 0;
        const uint2 r = r1;
        k1 = k2; v1 = v2;
        r1 = k1 != 0xffffffffu ? rects[v1] : make_uint2(0u, 0u);
        fetch_kv(e + 2 * stride, k2, v2);
        if (p.valid) p.mask = supertile_mask(unpack_rect(r), (int)(p.st % (uint32_t)sgx), (int)(p.st / (uint32_t)sgx));
        return p;
    }
};

// lane b receives the ballot of the `sel` lanes whose mask covers bin b
__device__ __forceinline__ uint32_t bin_ballots(bool sel, uint32_t mask, int lane)
{
    uint32_t mine = 0;
    const uint32_t m = sel ? mask : 0u;
#pragma unroll
    for (int b = 0; b < kBins; ++b) {
        const uint32_t bal = __ballot_sync(0xffffffffu, (m >> b) & 1u);
        if (lane == b) mine = bal;
    }
    return mine;
}

__device__ __forceinline__ int tile_of(uint32_t st, int bin, int sgx, int gx)
{
    const int sx = (int)(st % (uint32_t)sgx), sy = (int)(st / (uint32_t)sgx);
    return (kSuperY * sy + (bin >> 3)) * gx + kSuperX * sx + (bin & 7);   // only used for covered bins: always inside the grid
}

// Per chunk c (256 sorted pairs): tail_cnt[c][b] = pairs of the chunk's LAST super-tile covering bin b.  Pairs of
// other super-tiles in the chunk (a super-tile list ends inside it) go straight into the per-tile totals.  Also
// records where each super-tile's list begins and ends.
__global__ void __launch_bounds__(kChunk)
bin_count_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, const uint2 *__restrict__ rects,
                 const uint32_t *__restrict__ n_dev, int sgx, int gx, uint32_t *__restrict__ tail_cnt,
                 uint32_t *__restrict__ seg_begin, uint32_t *__restrict__ seg_end, uint32_t *__restrict__ tile_total)
{
    __shared__ uint32_t s_acc[kBins];
    const uint32_t n = *n_dev;
    const uint32_t chunks = (n + kChunk - 1) / kChunk;
    const int lane = threadIdx.x & 31;
    PairStream ps{keys, vals, rects, n, gridDim.x * (uint32_t)kChunk};
    ps.start(blockIdx.x * kChunk + threadIdx.x);
    for (uint32_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        if (threadIdx.x < kBins) s_acc[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t e = c * kChunk + threadIdx.x;
        const PairLane p = ps.next(e, sgx);
        const uint32_t tail_st = keys[min(c * kChunk + kChunk - 1, n - 1)];
        // list boundaries: the neighbours' keys come from the adjacent lanes (global memory only across warps)
        uint32_t prev = __shfl_up_sync(0xffffffffu, p.st, 1), nxt = __shfl_down_sync(0xffffffffu, p.st, 1);
        if (p.valid) {
            if (lane == 0) prev = e ? keys[e - 1] : 0xffffffffu;
            if (lane == 31) nxt = e + 1 < n ? keys[e + 1] : 0xffffffffu;
            if (e + 1 == n) nxt = 0xffffffffu;
            if (prev != p.st) seg_begin[p.st] = e;
            if (nxt != p.st) seg_end[p.st] = e + 1;
        }
        const uint32_t cnt = __popc(bin_ballots(p.valid && p.st == tail_st, p.mask, lane));
        if (cnt) atomicAdd(&s_acc[lane], cnt);
        // pairs of other super-tiles: one pass per distinct super-tile in the warp (rare)
        uint32_t todo = __ballot_sync(0xffffffffu, p.valid && p.st != tail_st);
        while (todo) {
            const uint32_t s = __shfl_sync(0xffffffffu, p.st, __ffs(todo) - 1);
            const bool sel = p.valid && p.st == s;
            todo &= ~__ballot_sync(0xffffffffu, sel);
            const uint32_t k = __popc(bin_ballots(sel, p.mask, lane));
            if (k) atomicAdd(&tile_total[tile_of(s, lane, sgx, gx)], k);
        }
        __syncthreads();
        if (threadIdx.x < kBins) tail_cnt[(size_t)c * kBins + threadIdx.x] = s_acc[threadIdx.x];
        __syncthreads();
    }
}

// One thread per (super-tile, bin): running sum of tail_cnt over the chunks whose last pair belongs to the
// super-tile.  run_prefix[c + 1][b] = pairs of that super-tile covering bin b in chunks <= c -- the offset of the
// pairs at the head of chunk c + 1 that continue the list.  The run total joins the per-tile totals.
__global__ void __launch_bounds__(256)
bin_runs_kernel(const uint32_t *__restrict__ n_dev, const uint32_t *__restrict__ seg_begin,
                const uint32_t *__restrict__ seg_end, int num_super, int sgx, int gx,
                const uint32_t *__restrict__ tail_cnt, uint32_t *__restrict__ run_prefix,
                uint32_t *__restrict__ tile_total)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = t / kBins, b = t % kBins;
    if (s >= num_super) return;
    const uint32_t n = *n_dev;
    const uint32_t begin = seg_begin[s], end = seg_end[s];
    if (end <= begin || n == 0) return;
    const uint32_t last_chunk = (n - 1) / kChunk;
    // chunk c ends in this list iff its last pair min(256c + 255, n - 1) lies in [begin, end)
    const int64_t c_first = begin / kChunk;
    const int64_t c_last = end == n ? (int64_t)last_chunk : (int64_t)(end / kChunk) - 1;
    uint32_t acc = 0;
    for (int64_t c = c_first; c <= c_last; ++c) {
        acc += tail_cnt[(size_t)c * kBins + b];
        run_prefix[(size_t)(c + 1) * kBins + b] = acc;
    }
    if (acc) atomicAdd(&tile_total[tile_of((uint32_t)s, b, sgx, gx)], acc);
}

// Exclusive scan of the per-tile totals in tile order -> ranges (empty tiles keep upstream's {0, 0}).
__global__ void __launch_bounds__(1024)
tile_starts_kernel(const uint32_t *__restrict__ tile_total, int tiles, int64_t R_cap, uint32_t *__restrict__ tile_start,
                   uint32_t *__restrict__ ranges)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t v = t < tiles ? tile_total[t] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = s_carry;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        const uint32_t start = before + incl - v;
        if (t < tiles) {
            tile_start[t] = start;
            // on an instance overflow the lists are clipped to the capacity (the caller re-runs the frame)
            const uint32_t lo = (uint32_t)min((int64_t)start, R_cap), hi = (uint32_t)min((int64_t)start + v, R_cap);
            ranges[2 * t] = v ? lo : 0u;
            ranges[2 * t + 1] = v ? hi : 0u;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + incl;
        __syncthreads();
    }
}

// point_list[start(tile) + pairs of the same list in earlier chunks + rank inside the chunk] = id
__global__ void __launch_bounds__(kChunk)
bin_write_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, const uint2 *__restrict__ rects,
                 const uint32_t *__restrict__ n_dev, int sgx, int gx, const uint32_t *__restrict__ run_prefix,
                 const uint32_t *__restrict__ tile_start, int64_t R_cap, uint32_t *__restrict__ point_list)
{
    __shared__ uint32_t s_bal[kChunkWarps][kBins];    // per warp, per bin: ballot of the pairs of the warp's LAST super-tile
    __shared__ uint32_t s_bal2[kChunkWarps][kBins];   // the same for another super-tile of the warp (a list ends inside it)
    __shared__ uint32_t s_base[kChunkWarps][kBins];
    __shared__ uint32_t s_wtail[kChunkWarps];
    const uint32_t n = *n_dev;
    const uint32_t chunks = (n + kChunk - 1) / kChunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    PairStream ps{keys, vals, rects, n, gridDim.x * (uint32_t)kChunk};
    ps.start(blockIdx.x * kChunk + threadIdx.x);
    for (uint32_t c = blockIdx.x; c < chunks; c += gridDim.x) {
        const uint32_t e = c * kChunk + threadIdx.x;
        const PairLane p = ps.next(e, sgx);
        const uint32_t head_st = keys[c * kChunk];
        const bool continues = c > 0 && keys[c * kChunk - 1] == head_st;   // the chunk starts inside a list
        const uint32_t vmask = __ballot_sync(0xffffffffu, p.valid);
        const uint32_t wtail = vmask ? __shfl_sync(0xffffffffu, p.st, 31 - __clz(vmask)) : 0xffffffffu;
        s_bal[warp][lane] = bin_ballots(p.valid && p.st == wtail, p.mask, lane);
        if (lane == 0) s_wtail[warp] = wtail;
        __syncthreads();
        uint32_t todo = vmask;
        while (todo) {
            const uint32_t s = __shfl_sync(0xffffffffu, p.st, __ffs(todo) - 1);
            const bool sel = p.valid && p.st == s;
            todo &= ~__ballot_sync(0xffffffffu, sel);
            const uint32_t *bal = s_bal[warp];
            if (s != wtail) {   // warp-uniform, rare
                s_bal2[warp][lane] = bin_ballots(sel, p.mask, lane);
                bal = s_bal2[warp];
            }
            __syncwarp();
            // lane b: offset of bin b for this warp's pairs of super-tile s
            uint32_t base = 0;
#pragma unroll
            for (int w = 0; w < kChunkWarps; ++w)
                if (w < warp && s_wtail[w] == s) base += __popc(s_bal[w][lane]);   // earlier warps (lists are contiguous)
            if (continues && s == head_st) base += run_prefix[(size_t)c * kBins + lane];
            if (bal[lane]) base += tile_start[tile_of(s, lane, sgx, gx)];
            s_base[warp][lane] = base;
            __syncwarp();
            uint32_t m = sel ? p.mask : 0u;   // ~3 tiles per pair: every lane walks its own bits
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t pos = s_base[warp][b] + __popc(bal[b] & lt);
                if ((int64_t)pos < R_cap) point_list[pos] = p.gid;
            }
            __syncwarp();
        }
        __syncthreads();
    }
}

// ---- launchers ---------------------------------------------------------------------------------------
void launch_scan_emit_pairs(const uint32_t *order, const uint2 *rects, int P_cap, const uint32_t *p_dev, int64_t R_cap,
                            int sgx, uint32_t *pair_keys, uint32_t *pair_vals, unsigned long long *state,
                            uint32_t *ticket, unsigned long long *r_acc, uint32_t *done, uint32_t *n_pairs_dev,
                            int32_t *status, cudaStream_t st)
{
    if (P_cap <= 0) return;
    scan_emit_pairs_kernel<<<(P_cap + kScanTile - 1) / kScanTile, kScanThreads, 0, st>>>(
        order, rects, p_dev, R_cap, sgx, pair_keys, pair_vals, state, ticket, r_acc, done, n_pairs_dev, status);
}

int bin_chunks_cap(int64_t R_cap) { return (int)ceil_div64(R_cap > 0 ? R_cap : 1, kChunk); }

void launch_bin_expand(const uint32_t *keys, const uint32_t *vals, const uint2 *rects, const uint32_t *n_pairs_dev,
                       int64_t R_cap, int gx, int gy, uint32_t *tail_cnt, uint32_t *run_prefix, uint32_t *seg_begin,
                       uint32_t *seg_end, uint32_t *tile_total, uint32_t *tile_start, uint32_t *ranges,
                       uint32_t *point_list, cudaStream_t st)
{
    const int sgx = (gx + kSuperX - 1) / kSuperX, sgy = (gy + kSuperY - 1) / kSuperY;
    const int num_super = sgx * sgy;
    // persistent grids: the pair count lives on the device, a capacity-sized grid would mostly launch empty CTAs
    const int grid = (int)min((int64_t)kNumSMs * 8, (int64_t)bin_chunks_cap(R_cap));
    bin_count_kernel<<<grid, kChunk, 0, st>>>(keys, vals, rects, n_pairs_dev, sgx, gx, tail_cnt, seg_begin, seg_end,
                                              tile_total);
    bin_runs_kernel<<<(num_super * kBins + 255) / 256, 256, 0, st>>>(n_pairs_dev, seg_begin, seg_end, num_super, sgx, gx,
                                                                     tail_cnt, run_prefix, tile_total);
    tile_starts_kernel<<<1, 1024, 0, st>>>(tile_total, gx * gy, R_cap, tile_start, ranges);
    bin_write_kernel<<<grid, kChunk, 0, st>>>(keys, vals, rects, n_pairs_dev, sgx, gx, run_prefix, tile_start, R_cap,
                                              point_list);
}

int scan_tiles_count(int P) { return (P + kScanTile - 1) / kScanTile; }

}  // namespace cgs

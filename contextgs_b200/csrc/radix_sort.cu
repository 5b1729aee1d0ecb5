// Library-free (no cub/thrust) stable LSD radix sort of (u32 key, u32 value) pairs, one-sweep
// style: one upfront histogram read for all digit places, then one chained-scan ("decoupled
// look-back") scatter kernel per 8-bit digit.  Per pass each key is read once and written once.
//
// B200 sizing: 4096 keys per CTA (256 threads x 16 keys), tiles handed out by an atomic ticket so
// that look-back predecessors are always resident; the element count lives on the device
// (*n_dev) so the rasterizer never has to read `num_rendered` back to the host.
//
// Replaces the cub::DeviceRadixSort::SortPairs call of the upstream rasterizer
// (rasterizer_impl.cu, not in the reference tree) -- see raster_binning.cu for how the 64-bit
// (tile<<32 | depth) sort is decomposed into two narrower sorts with an identical result.
#include "common.cuh"

namespace cgs {

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1;
constexpr int kSortWarps = kSortThreads / 32;

// hist[pass][256]: digit counts of every pass, one read of the keys.
__global__ void __launch_bounds__(256)
radix_histogram_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ n_dev, uint32_t n_cap,
                       int begin_bit, int end_bit, int npass, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t sh[4 * kRadix];
    for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t n = min(*n_dev, n_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        for (int p = 0; p < npass; ++p) {
            const int shift = begin_bit + p * kRadixBits;
            const int bits = min(kRadixBits, end_bit - shift);
            atomicAdd(&sh[p * kRadix + ((k >> shift) & ((1u << bits) - 1))], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// In-place exclusive scan of each 256-bin histogram (one CTA of 256 threads per pass).
__global__ void __launch_bounds__(256) radix_scan_hist_kernel(uint32_t *__restrict__ hist)
{
    __shared__ uint32_t warp_sums[8];
    uint32_t *h = hist + blockIdx.x * kRadix;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t v = h[threadIdx.x];
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += warp_sums[w];
    h[threadIdx.x] = base + incl - v;
}

// One digit pass.  lookback[tile][256] holds (flag | count) words; ticket hands out tiles.
//
// Per tile of 4096 pairs: (1) warp-local stable ranking with match.any, (2) thread d owns digit d:
// prefix over the 8 warps, a block scan over the 256 digits (tile-local sorted position of each
// digit's run) and the chained look-back across tiles, (3) the pairs are first scattered into SHARED
// memory in tile-sorted order and only then copied out, so that consecutive threads write
// consecutive global addresses (one run per digit) instead of 32 scattered 4-byte stores per warp.
__global__ void __launch_bounds__(kSortThreads, 4)
onesweep_pass_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                     const uint32_t *__restrict__ n_dev, uint32_t n_cap, int shift, int bits,
                     const uint32_t *__restrict__ global_base, uint32_t *lookback, uint32_t tiles_cap,
                     uint32_t *ticket)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t warp_hist[kSortWarps][kRadix];
    __shared__ uint32_t digit_start[kRadix];   // tile-local sorted position of the first key of each digit
    __shared__ uint32_t digit_gbase[kRadix];   // global position of that key minus digit_start
    __shared__ uint32_t s_wsum[kSortWarps];
    __shared__ uint32_t s_key[kSortTile];
    __shared__ uint32_t s_val[kSortTile];

    const uint32_t n = min(*n_dev, n_cap);
    const uint32_t num_tiles = (n + kSortTile - 1) / kSortTile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= num_tiles) return;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane_lt = (1u << lane) - 1;
    const uint32_t mask = (1u << bits) - 1;
    const uint32_t tile_base = tile * kSortTile;
    const uint32_t warp_base = tile_base + warp * (32 * kSortItems);
    const uint32_t tile_count = min((uint32_t)kSortTile, n - tile_base);

    uint32_t key[kSortItems];
    uint16_t rank[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t idx = warp_base + i * 32 + lane;
        key[i] = idx < n ? keys_in[idx] : 0xFFFFFFFFu;
    }
    // warp-local stable ranking (match-any), items visited in memory order
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t idx = warp_base + i * 32 + lane;
        const bool valid = idx < n;
        const uint32_t digit = valid ? ((key[i] >> shift) & mask) : kRadix;  // invalid lanes form their own group
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid) {
            old = warp_hist[warp][digit];
            warp_hist[warp][digit] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[i] = (uint16_t)(old + __popc(peers & lane_lt));
        __syncwarp();
    }
    __syncthreads();

    // thread d owns digit d: prefix over the warps, aggregate published EARLY, block scan over digits
    uint32_t my_sum;
    {
        const int d = threadIdx.x;
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = warp_hist[w][d];
            warp_hist[w][d] = sum;
            sum += c;
        }
        my_sum = sum;
        volatile uint32_t *lb = lookback;  // layout [tile][digit]
        lb[(size_t)tile * kRadix + d] = (tile == 0 ? kFlagInclusive : kFlagAggregate) | sum;
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) before += w < warp ? s_wsum[w] : 0u;
        digit_start[d] = before + incl - sum;
    }
    __syncthreads();

    // scatter into shared memory in tile-sorted order (frees the key / rank registers)
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t idx = warp_base + i * 32 + lane;
        if (idx < n) {
            const uint32_t digit = (key[i] >> shift) & mask;
            const uint32_t lp = digit_start[digit] + warp_hist[warp][digit] + rank[i];
            s_key[lp] = key[i];
            s_val[lp] = vals_in ? vals_in[idx] : idx;
        }
    }

    // Chained look-back, one thread per digit.  (A warp-wide variant that reads 32 predecessor tiles
    // x 32 digits per round trip was measured 2x SLOWER here: the extra L2 traffic costs more than the
    // shorter chains save -- the aggregates are published early, so chains are short already.)
    {
        const int d = threadIdx.x;
        volatile uint32_t *lb = lookback;
        uint32_t excl = 0;
        if (tile != 0) {
            int64_t t = (int64_t)tile - 1;
            while (true) {
                uint32_t v = lb[(size_t)t * kRadix + d];
                while ((v >> 30) == 0) v = lb[(size_t)t * kRadix + d];
                excl += v & kValueMask;
                if ((v >> 30) == 2u) break;
                --t;
            }
            lb[(size_t)tile * kRadix + d] = kFlagInclusive | (excl + my_sum);
        }
        digit_gbase[d] = global_base[d] + excl - digit_start[d];
    }
    __syncthreads();
    // coalesced copy-out: position j of the tile-sorted order goes to digit_gbase[digit] + j
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t j = i * kSortThreads + threadIdx.x;
        if (j < tile_count) {
            const uint32_t k = s_key[j];
            const uint32_t pos = digit_gbase[(k >> shift) & mask] + j;
            keys_out[pos] = k;
            vals_out[pos] = s_val[j];
        }
    }
}

SortPlan make_sort_plan(int64_t n_cap, int begin_bit, int end_bit)
{
    SortPlan p;
    const int nbits = end_bit > begin_bit ? end_bit - begin_bit : 1;
    p.npass = (nbits + kRadixBits - 1) / kRadixBits;
    if (p.npass > 4) p.npass = 4;
    p.tiles_cap = ceil_div64(n_cap > 0 ? n_cap : 1, kSortTile);
    size_t off = 0;
    p.hist_off = off;
    off += align_up((size_t)p.npass * kRadix * sizeof(uint32_t));
    p.ticket_off = off;
    off += align_up(8 * sizeof(uint32_t));
    p.lookback_off = off;
    off += align_up((size_t)p.npass * (size_t)p.tiles_cap * kRadix * sizeof(uint32_t));
    p.zero_bytes = off;
    p.total_bytes = off;
    return p;
}

int sort_pairs(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
               uint32_t *keys_tmp, uint32_t *vals_tmp, const uint32_t *n_dev, int64_t n_cap, int begin_bit,
               int end_bit, void *ws, bool zero_ws, cudaStream_t stream)
{
    if (n_cap <= 0) return 0;
    if (n_cap >= (int64_t)kValueMask) {
        set_error("sort_pairs: n_cap %lld exceeds the 30-bit look-back counter", (long long)n_cap);
        return -2;
    }
    if (end_bit - begin_bit > 32 || end_bit > 32 || begin_bit < 0) {
        set_error("sort_pairs: invalid bit range [%d,%d)", begin_bit, end_bit);
        return -2;
    }
    const SortPlan plan = make_sort_plan(n_cap, begin_bit, end_bit);
    char *base = static_cast<char *>(ws);
    if (zero_ws) cudaMemsetAsync(base, 0, plan.zero_bytes, stream);
    uint32_t *hist = reinterpret_cast<uint32_t *>(base + plan.hist_off);
    uint32_t *ticket = reinterpret_cast<uint32_t *>(base + plan.ticket_off);
    uint32_t *lookback = reinterpret_cast<uint32_t *>(base + plan.lookback_off);
    if (end_bit <= begin_bit) end_bit = begin_bit + 1;

    const int hist_grid = (int)min((int64_t)kNumSMs * 8, ceil_div64(n_cap, 256 * 4));
    radix_histogram_kernel<<<hist_grid, 256, 0, stream>>>(keys_in, n_dev, (uint32_t)n_cap, begin_bit, end_bit,
                                                          plan.npass, hist);
    radix_scan_hist_kernel<<<plan.npass, 256, 0, stream>>>(hist);

    const uint32_t *kin = keys_in;
    const uint32_t *vin = vals_in;
    for (int p = 0; p < plan.npass; ++p) {
        const bool last = (p == plan.npass - 1);
        // choose outputs so that the last pass lands in (keys_out, vals_out)
        const bool to_out = ((plan.npass - 1 - p) % 2 == 0);
        uint32_t *ko = to_out ? keys_out : keys_tmp;
        uint32_t *vo = to_out ? vals_out : vals_tmp;
        (void)last;
        const int shift = begin_bit + p * kRadixBits;
        const int bits = min(kRadixBits, end_bit - shift);
        onesweep_pass_kernel<<<(unsigned)plan.tiles_cap, kSortThreads, 0, stream>>>(
            kin, vin, ko, vo, n_dev, (uint32_t)n_cap, shift, bits, hist + p * kRadix,
            lookback + (size_t)p * plan.tiles_cap * kRadix, (uint32_t)plan.tiles_cap, ticket + p);
        kin = ko;
        vin = vo;
    }
    return check_launch("sort_pairs");
}

}  // namespace cgs

extern "C" size_t cgs_sort_workspace_bytes(int64_t n_cap, int begin_bit, int end_bit)
{
    return cgs::make_sort_plan(n_cap, begin_bit, end_bit).total_bytes;
}

extern "C" int cgs_sort_pairs_u32(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out,
                                  uint32_t *vals_out, uint32_t *keys_tmp, uint32_t *vals_tmp, const uint32_t *n_dev,
                                  int64_t n_cap, int begin_bit, int end_bit, void *workspace, size_t workspace_bytes,
                                  void *stream)
{
    if (n_cap <= 0) return 0;
    CGS_CHECK_PTR(keys_in);
    CGS_CHECK_PTR(keys_out);
    CGS_CHECK_PTR(vals_out);
    CGS_CHECK_PTR(n_dev);
    CGS_CHECK_PTR(workspace);
    const cgs::SortPlan plan = cgs::make_sort_plan(n_cap, begin_bit, end_bit);
    if (workspace_bytes < plan.total_bytes) {
        cgs::set_error("cgs_sort_pairs_u32: workspace %zu < %zu bytes", workspace_bytes, plan.total_bytes);
        return -3;
    }
    if (plan.npass > 1) {
        CGS_CHECK_PTR(keys_tmp);
        CGS_CHECK_PTR(vals_tmp);
    }
    return cgs::sort_pairs(keys_in, vals_in, keys_out, vals_out, keys_tmp, vals_tmp, n_dev, n_cap, begin_bit, end_bit,
                           workspace, true, static_cast<cudaStream_t>(stream));
}

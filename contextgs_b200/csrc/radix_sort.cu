// Library-free (no cub/thrust) stable LSD radix sort of (u32 key, u32 value) pairs, one-sweep
// style: one upfront histogram read for all digit places, then one chained-scan ("decoupled
// look-back") scatter kernel per 8-bit digit.  Per pass each key is read once and written once.
//
// B200 sizing: 6144 keys per CTA (384 threads x 16 keys), tiles handed out by an atomic ticket so
// that look-back predecessors are always resident; the element count lives on the device
// (*n_dev) so the rasterizer never has to read `num_rendered` back to the host.
//
// Replaces the cub::DeviceRadixSort::SortPairs call of the upstream rasterizer
// (rasterizer_impl.cu, not in the reference tree) -- see raster_binning.cu for how the 64-bit
// (tile<<32 | depth) sort is decomposed into two narrower sorts with an identical result.
#include <stdlib.h>

#include "common.cuh"

namespace cgs {

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1;

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
// Per tile of THREADS x ITEMS pairs: (1) warp-local stable ranking: the match.any masks of a chunk of
// items are computed back to back, then one leader lane per digit group bumps the per-warp digit counter; (2) thread d owns digit d: prefix over the warps, a block scan over the 256 digits
// (tile-local sorted position of each digit's run) and the chained look-back across tiles; (3) the pairs
// are first scattered into SHARED memory in tile-sorted order and only then copied out, so that
// consecutive threads write consecutive global addresses (one run per digit) instead of 32 scattered
// 4-byte stores per warp.  Values are fetched right after the keys, so their HBM latency hides behind
// the ranking.
template <int THREADS, int ITEMS>
struct SortSmem {
    static constexpr int kWarps = THREADS / 32, kTileKeys = THREADS * ITEMS;
    uint32_t warp_hist[kWarps][kRadix];
    uint32_t digit_start[kRadix];   // tile-local sorted position of the first key of each digit
    uint32_t digit_gbase[kRadix];   // global position of that key minus digit_start
    uint32_t wsum[8];
    uint32_t tile;
    uint32_t key[kTileKeys];
    uint32_t val[kTileKeys];
};

template <int THREADS, int ITEMS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
onesweep_pass_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                     const uint32_t *__restrict__ n_dev, uint32_t n_cap, int shift, int bits,
                     const uint32_t *__restrict__ global_base, uint32_t *lookback, uint32_t tiles_cap,
                     uint32_t *ticket)
{
    using Smem = SortSmem<THREADS, ITEMS>;
    constexpr int kWarps = Smem::kWarps, kTileKeys = Smem::kTileKeys;
    constexpr int kChunk = ITEMS < 8 ? ITEMS : 8;
    static_assert(THREADS >= kRadix && ITEMS % kChunk == 0 && ITEMS % 2 == 0, "tile shape");
    extern __shared__ __align__(16) unsigned char sort_smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(sort_smem_raw);

    const uint32_t n = min(*n_dev, n_cap);
    const uint32_t num_tiles = (n + kTileKeys - 1) / kTileKeys;
    if (threadIdx.x == 0) S.tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < kWarps * kRadix; i += THREADS) (&S.warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = S.tile;
    if (tile >= num_tiles) return;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane_lt = (1u << lane) - 1;
    const uint32_t mask = (1u << bits) - 1;
    const uint32_t tile_base = tile * kTileKeys;
    const uint32_t warp_base = tile_base + warp * (32 * ITEMS);
    const uint32_t tile_count = min((uint32_t)kTileKeys, n - tile_base);
    const bool full = tile_count == (uint32_t)kTileKeys;

    uint32_t key[ITEMS], val[ITEMS];
    uint32_t rank2[ITEMS / 2];  // two 16-bit warp-local ranks per register
    if (full) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = keys_in[warp_base + i * 32 + lane];
        if (vals_in) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) val[i] = vals_in[warp_base + i * 32 + lane];
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) val[i] = warp_base + i * 32 + lane;
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t idx = warp_base + i * 32 + lane;
            key[i] = idx < n ? keys_in[idx] : 0xFFFFFFFFu;
            val[i] = idx < n ? (vals_in ? vals_in[idx] : idx) : 0u;
        }
    }
    // warp-local stable ranking (match-any), items visited in memory order, kChunk items in flight
    uint32_t *wh = S.warp_hist[warp];
#pragma unroll
    for (int c = 0; c < ITEMS; c += kChunk) {
        uint32_t peers[kChunk];
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
            // Peers = lanes holding the same digit, built from one ballot per digit bit.  MATCH.ANY retires one
            // distinct value per step, and a warp of depth- or tile-ordered keys holds ~25 distinct digits:
            // ncu (profiles/r01_sort3_*) showed ~1000 stall cycles per item on its consumer.  Eight independent
            // ballots cost more issue slots but no serialisation.
            const bool valid = full || (warp_base + (c + i) * 32 + lane) < n;
            const uint32_t digit = (key[c + i] >> shift) & mask;
            uint32_t p = full ? 0xffffffffu : __ballot_sync(0xffffffffu, valid);
#pragma unroll
            for (int b = 0; b < kRadixBits; ++b) {
                if (b < bits) {   // warp-uniform
                    const bool bit = (digit >> b) & 1u;
                    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                    p &= bit ? bal : ~bal;
                }
            }
            peers[i] = valid ? p : (1u << lane);   // invalid lanes: a group of their own, never written
        }
#pragma unroll
        for (int i = 0; i < kChunk; ++i) {
            const bool valid = full || (warp_base + (c + i) * 32 + lane) < n;
            const int leader = __ffs(peers[i]) - 1;
            uint32_t o = 0;
            if (lane == leader && valid) {   // one lane per digit group: plain read-modify-write (ATOMS costs 2 cyc/lane)
                const uint32_t digit = (key[c + i] >> shift) & mask;
                o = wh[digit];
                wh[digit] = o + __popc(peers[i]);
            }
            o = __shfl_sync(0xffffffffu, o, leader);
            const uint32_t r = o + __popc(peers[i] & lane_lt);
            if ((i & 1) == 0) rank2[(c + i) / 2] = r;
            else rank2[(c + i) / 2] |= r << 16;
            __syncwarp();
        }
    }
    __syncthreads();

    // thread d owns digit d: prefix over the warps, aggregate published EARLY, block scan over digits
    uint32_t my_sum = 0;
    if (threadIdx.x < kRadix) {
        const int d = threadIdx.x;
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = S.warp_hist[w][d];
            S.warp_hist[w][d] = sum;
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
        if (lane == 31) S.wsum[warp] = incl;
        // (the second half of the scan needs all eight warp sums)
        asm volatile("bar.sync 1, 256;");
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < kRadix / 32; ++w) before += w < warp ? S.wsum[w] : 0u;
        S.digit_start[d] = before + incl - sum;
    }
    __syncthreads();

    // scatter into shared memory in tile-sorted order (frees the key / rank registers)
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (full || (warp_base + i * 32 + lane) < n) {
            const uint32_t digit = (key[i] >> shift) & mask;
            const uint32_t lp = S.digit_start[digit] + wh[digit] + ((rank2[i / 2] >> (16 * (i & 1))) & 0xffffu);
            S.key[lp] = key[i];
            S.val[lp] = val[i];
        }
    }

    // Chained look-back, one thread per digit.  (A warp-wide variant that reads 32 predecessor tiles
    // x 32 digits per round trip was measured 2x SLOWER here: the extra L2 traffic costs more than the
    // shorter chains save -- the aggregates are published early, so chains are short already.)
    if (threadIdx.x < kRadix) {
        const int d = threadIdx.x;
        volatile uint32_t *lb = lookback;
        uint32_t excl = 0;
        if (tile != 0) {
            // The walk reads kLook predecessors per round trip (independent loads in flight) and consumes them
            // in order: with hundreds of tiles resident the nearest INCLUSIVE prefix is often dozens of tiles
            // back, and one dependent L2 round trip per predecessor was the critical path of the whole pass.
            constexpr int kLook = 8;
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t v[kLook];
#pragma unroll
                for (int j = 0; j < kLook; ++j)
                    v[j] = t - j >= 0 ? lb[(size_t)(t - j) * kRadix + d] : (2u << 30);  // virtual inclusive zero before tile 0
                int used = 0;
#pragma unroll
                for (int j = 0; j < kLook; ++j) {
                    if (!done && used == j) {
                        if ((v[j] >> 30) != 0) {
                            excl += v[j] & kValueMask;
                            done = (v[j] >> 30) == 2u;
                            used = j + 1;
                        }
                    }
                }
                t -= used;   // used == 0: the nearest predecessor has not published yet -- poll again
            }
            lb[(size_t)tile * kRadix + d] = kFlagInclusive | (excl + my_sum);
        }
        S.digit_gbase[d] = global_base[d] + excl - S.digit_start[d];
    }
    __syncthreads();
    // coalesced copy-out: position j of the tile-sorted order goes to digit_gbase[digit] + j
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = i * THREADS + threadIdx.x;
        if (j < tile_count) {
            const uint32_t k = S.key[j];
            const uint32_t pos = S.digit_gbase[(k >> shift) & mask] + j;
            keys_out[pos] = k;
            vals_out[pos] = S.val[j];
        }
    }
}

// Tile shape of the digit pass.  The default was chosen by measurement on B200 (profiles/); the
// environment variable CGS_SORT_VARIANT (read once) selects another one for experiments.
struct SortVariant {
    int threads, items;
};
static const SortVariant kSortVariants[] = {{256, 16}, {512, 16}, {512, 8}, {256, 24}, {384, 16}, {1024, 8}, {256, 16}, {512, 16}};
static int sort_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CGS_SORT_VARIANT");
        v = e ? atoi(e) : 4;   // 384 x 16: fastest of the measured shapes (profiles/r01_sort_variants.txt)
        if (v < 0 || v >= (int)(sizeof(kSortVariants) / sizeof(kSortVariants[0]))) v = 4;
    }
    return v;
}
static int sort_tile_keys() { return kSortVariants[sort_variant()].threads * kSortVariants[sort_variant()].items; }

template <int THREADS, int ITEMS, int MIN_BLOCKS>
static void launch_onesweep(unsigned grid, cudaStream_t stream, const uint32_t *kin, const uint32_t *vin, uint32_t *ko,
                            uint32_t *vo, const uint32_t *n_dev, uint32_t n_cap, int shift, int bits,
                            const uint32_t *gbase, uint32_t *lookback, uint32_t tiles_cap, uint32_t *ticket)
{
    auto kern = onesweep_pass_kernel<THREADS, ITEMS, MIN_BLOCKS>;
    constexpr size_t smem = sizeof(SortSmem<THREADS, ITEMS>);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    kern<<<grid, THREADS, smem, stream>>>(kin, vin, ko, vo, n_dev, n_cap, shift, bits, gbase, lookback, tiles_cap, ticket);
}

SortPlan make_sort_plan(int64_t n_cap, int begin_bit, int end_bit)
{
    SortPlan p;
    const int nbits = end_bit > begin_bit ? end_bit - begin_bit : 1;
    p.npass = (nbits + kRadixBits - 1) / kRadixBits;
    if (p.npass > 4) p.npass = 4;
    p.tiles_cap = ceil_div64(n_cap > 0 ? n_cap : 1, sort_tile_keys());
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
        const unsigned grid = (unsigned)plan.tiles_cap;
        const uint32_t *gb = hist + p * kRadix;
        uint32_t *lb = lookback + (size_t)p * plan.tiles_cap * kRadix;
        const uint32_t tc = (uint32_t)plan.tiles_cap;
        switch (sort_variant()) {
        default:
        case 0: launch_onesweep<256, 16, 3>(grid, stream, kin, vin, ko, vo, n_dev, (uint32_t)n_cap, shift, bits, gb, lb, tc, ticket + p); break;
        case 1: launch_onesweep<512, 16, 2>(grid, stream, kin, vin, ko, vo, n_dev, (uint32_t)n_cap, shift, bits, gb, lb, tc, ticket + p); break;
        case 2: launch_onesweep<512, 8, 2>(grid, stream, kin, vin, ko, vo, n_dev, (uint32_t)n_cap, shift, bits, gb, lb, tc, ticket + p); break;
        case 3: launch_onesweep<256, 24, 2>(grid, stream, kin, vin, ko, vo, n_dev, (uint32_t)n_cap, shift, bits, gb, lb, tc, ticket + p); break;
        case 4: launch_onesweep<384, 16, 2>(grid, stream, kin, vin, ko, vo, n_dev, (uint32_t)n_cap, shift, bits, gb, lb, tc, ticket + p); break;
        case 5: launch_onesweep<1024, 8, 1>(grid, stream, kin, vin, ko, vo, n_dev, (uint32_t)n_cap, shift, bits, gb, lb, tc, ticket + p); break;
        case 6: launch_onesweep<256, 16, 2>(grid, stream, kin, vin, ko, vo, n_dev, (uint32_t)n_cap, shift, bits, gb, lb, tc, ticket + p); break;
        case 7: launch_onesweep<512, 16, 1>(grid, stream, kin, vin, ko, vo, n_dev, (uint32_t)n_cap, shift, bits, gb, lb, tc, ticket + p); break;
        }
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

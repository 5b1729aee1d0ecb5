// Tile binning without a 64-bit sort and without a host read-back.
//
// Upstream (rasterizer_impl.cu, not in the reference tree) emits one 64-bit key
// (tile<<32 | depth_bits) per (Gaussian, tile) instance and runs a 6-pass cub radix sort over all
// R instances, after a blocking D2H copy of R.  Here the same ordering is produced as
//   1. 4-pass sort of the P Gaussians by depth bits (stable -> ties keep ascending id);
//   2. chained-scan of tiles_touched in that depth order  -> instance offsets, R stays on device;
//   3. emit (tile id, Gaussian id) in depth order;
//   4. stable sort of the R instances by tile id only (ceil(log2(tiles)/8) = 2 passes at 1080p).
// A stable sort by tile of a depth-ordered stream is ordered by (tile, depth, id): exactly the
// order of the upstream 64-bit LSD sort, so point_list and ranges are bit-identical to it while
// moving ~4x fewer bytes (8-byte pairs x 2 passes over R instead of 12-byte pairs x 6 passes).
#include "common.cuh"

namespace cgs {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// Inclusive scan of tiles_touched gathered in depth order; the last tile publishes R.
__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const uint32_t *__restrict__ order, const float *__restrict__ geom,
                  const uint32_t *__restrict__ p_dev, int64_t R_cap, uint32_t *__restrict__ offsets,
                  unsigned long long *state, uint32_t *ticket, int32_t *__restrict__ status)
{
    const int P = (int)*p_dev;  // device-side Gaussian count (the grid is sized by the host's capacity)
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_warp[kScanThreads / 32];
    __shared__ uint64_t s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (P <= 0) {  // nothing to draw: publish R = 0 once
        if (tile == 0 && threadIdx.x == 0) {
            status[CGS_STATUS_NUM_RENDERED] = 0;
            status[CGS_STATUS_OVERFLOW] = 0;
            status[CGS_STATUS_NUM_SORTED] = 0;
        }
        return;
    }
    if ((int64_t)tile * kScanTile >= (int64_t)P) return;  // tiles beyond the device-side count
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = tile * kScanTile + threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t local = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int idx = base + i;
        uint32_t t = 0;
        if (idx < P) t = __float_as_uint(geom[(size_t)order[idx] * kGeomStride + G_TILES]);
        local += t;
        v[i] = local;
    }
    uint64_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t warp_excl = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        warp_excl += w < warp ? s_warp[w] : 0ull;
        total += s_warp[w];
    }
    if (warp == 0) {
        const uint64_t excl = lookback_exclusive(state, (int)tile, total);
        if (lane == 0) {
            s_prefix = excl;
            if ((int64_t)tile == ((int64_t)P - 1) / kScanTile) {
                const uint64_t R = excl + total;
                status[CGS_STATUS_NUM_RENDERED] = (int32_t)(R > 0x7fffffffull ? 0x7fffffff : R);
                status[CGS_STATUS_OVERFLOW] = R > (uint64_t)R_cap ? 1 : 0;
                status[CGS_STATUS_NUM_SORTED] = (int32_t)(R > (uint64_t)R_cap ? (uint64_t)R_cap : R);
            }
        }
    }
    __syncthreads();
    const uint64_t thread_excl = s_prefix + warp_excl + incl - local;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int idx = base + i;
        if (idx < P) {
            const uint64_t o = thread_excl + v[i];
            offsets[idx] = (uint32_t)(o > 0xffffffffull ? 0xffffffffull : o);
        }
    }
}

// Emit (tile id, Gaussian id) for every covered tile, y-major / x-minor, in depth order.
__global__ void __launch_bounds__(256)
emit_instances_kernel(const uint32_t *__restrict__ order, const float *__restrict__ geom,
                      const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ p_dev, int grid_x, int grid_y,
                      int64_t R_cap, uint32_t *__restrict__ tile_keys, uint32_t *__restrict__ inst_vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int)*p_dev) return;
    const uint32_t gid = order[i];
    const float4 g0 = *reinterpret_cast<const float4 *>(geom + (size_t)gid * kGeomStride);
    const float4 g2 = *reinterpret_cast<const float4 *>(geom + (size_t)gid * kGeomStride + 8);
    const int radius = __float_as_int(g2.z);
    if (radius <= 0) return;
    int64_t off = i == 0 ? 0 : (int64_t)offsets[i - 1];
    int x0, y0, x1, y1;
    get_rect(g0.x, g0.y, radius, grid_x, grid_y, x0, y0, x1, y1);
    for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
            if (off < R_cap) {
                tile_keys[off] = (uint32_t)(y * grid_x + x);
                inst_vals[off] = gid;
            }
            ++off;
        }
}

__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint32_t *__restrict__ tile_keys, const int32_t *__restrict__ status,
                   uint32_t *__restrict__ ranges)
{
    const uint32_t n = (uint32_t)status[CGS_STATUS_NUM_SORTED];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t t = tile_keys[i];
        if (i == 0) {
            ranges[2 * t] = 0;
        } else {
            const uint32_t prev = tile_keys[i - 1];
            if (prev != t) {
                ranges[2 * prev + 1] = i;
                ranges[2 * t] = i;
            }
        }
        if (i == n - 1) ranges[2 * t + 1] = n;
    }
}

void launch_scan_tiles(const uint32_t *order, const float *geom, int P_cap, const uint32_t *p_dev, int64_t R_cap,
                       uint32_t *offsets, unsigned long long *state, uint32_t *ticket, int32_t *status, cudaStream_t st)
{
    if (P_cap <= 0) return;
    scan_tiles_kernel<<<(P_cap + kScanTile - 1) / kScanTile, kScanThreads, 0, st>>>(order, geom, p_dev, R_cap, offsets,
                                                                                  state, ticket, status);
}

void launch_emit_instances(const uint32_t *order, const float *geom, const uint32_t *offsets, int P_cap,
                           const uint32_t *p_dev, int grid_x, int grid_y, int64_t R_cap, uint32_t *tile_keys,
                           uint32_t *inst_vals, cudaStream_t st)
{
    if (P_cap <= 0) return;
    emit_instances_kernel<<<(P_cap + 255) / 256, 256, 0, st>>>(order, geom, offsets, p_dev, grid_x, grid_y, R_cap,
                                                               tile_keys, inst_vals);
}

void launch_tile_ranges(const uint32_t *tile_keys, const int32_t *status, int64_t R_cap, uint32_t *ranges,
                        cudaStream_t st)
{
    const int grid = (int)min((int64_t)kNumSMs * 8, ceil_div64(R_cap > 0 ? R_cap : 1, 256));
    tile_ranges_kernel<<<grid, 256, 0, st>>>(tile_keys, status, ranges);
}

int scan_tiles_count(int P) { return (P + kScanTile - 1) / kScanTile; }

}  // namespace cgs

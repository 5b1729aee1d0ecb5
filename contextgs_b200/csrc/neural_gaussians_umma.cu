// Fused anchor -> neural-Gaussian generation on the tcgen05 tensor cores.
//
// Same contract as neural_gaussians.cu (reference: gaussian_renderer/__init__.py:106-145 and the
// decoder MLPs scene/gaussian_model.py:153-174), but the two MLP layers run as 3xTF32 tcgen05.mma
// (fp32-grade accuracy, see umma.cuh) with every activation resident in TENSOR MEMORY:
//
//   persistent CTA (one per SM), tile = 128 anchors = 128 TMEM lanes, 512 threads in TWO WARP GROUPS that
//   work on DIFFERENT tiles at the same time (the kernel is bound by the latency of its serial phases, not by
//   the tensor core, HBM or issue slots -- doubling the threads per phase bought nothing, overlapping the
//   phases of two tiles does):
//     FRONT (warps 0-7):  stage the layer-1 input of tile t+1 -> layer-1 MMA -> ReLU epilogue in TMEM
//                         -> issue the layer-2 MMAs as soon as BACK has released the accumulators of tile t
//                         -> coalesced copy-out of tile t from the staging area BACK has just filled;
//     BACK  (warps 8-15): selection, ordered compaction (block scan + decoupled look-back across tiles) and
//                         post-processing of tile t into the staging area.
//   Inside a group, thread (row = 32*(warp%4) + lane, half = (warp%8)/4): two threads share a row.
//
//   TMEM columns (496 of 512):
//     [  0,176)  layer-1 input  x_hi [0,56) | x_lo [56,112)      -> later hidden_lo [0,176)
//     [176,352)  layer-1 accumulator D1 (3 heads x 56: 50 units + 6 zero pads, + 8 pad)
//                -> overwritten IN PLACE by hidden_hi = tf32(relu(D1 + b1))
//     [352,496)  layer-2 accumulators: opacity 16 | color 48 | cov 80
//   Columns [0,352) belong to FRONT (free again once the layer-2 MMAs of the tile have completed),
//   [352,496) go back and forth: written by the layer-2 MMAs, read by BACK, released through an mbarrier.
//   shared memory (145 KB weights + 77 KB staging): W1 hi/lo [14][176][4], W2 per head hi/lo [14][N_h][4], biases.
//
// HBM traffic is the algorithmic minimum (446 B per visible anchor + 56 B per Gaussian): nothing
// but the final attributes is written.
#include "umma.cuh"

namespace cgs {

namespace ngu {
constexpr int kFeat = 50, kK = 10;
constexpr int kRows = 128;                  // anchors per tile
constexpr int kThreads = 512, kGroup = 256; // two warp groups of 8 warps
constexpr int kK1 = 56;                     // 54 inputs padded to a multiple of 8
constexpr int kHeadStride = 56;             // hidden units per head incl. zero pads
constexpr int kN1 = 176;                    // 3 * 56 = 168 padded to a multiple of 16
// layer-2 output widths.  Output columns are arranged so that every thread's tcgen05.ld starts on
// an aligned column: opacity of offset k at column 8*(k/5) + k%5, colour channel c of offset k
// at 4k + c, covariance value i of offset k at 8k + i (the host packs W2 / b2 rows accordingly).
constexpr int kNo = 16, kNc = 48, kNv = 80;
// TMEM columns
constexpr uint32_t kColXHi = 0, kColXLo = 56, kColHLo = 0, kColD1 = 176, kColDo = 352, kColDc = 368, kColDv = 416;
constexpr uint32_t kTmemCols = 512;
// packed weight block (floats)
constexpr int kW1 = (kK1 / 4) * kN1 * 4;    // 9856
constexpr int kW2o = (kK1 / 4) * kNo * 4;   // 896
constexpr int kW2c = (kK1 / 4) * kNc * 4;   // 2688
constexpr int kW2v = (kK1 / 4) * kNv * 4;   // 4480
constexpr int kOffW1Hi = 0, kOffW1Lo = kOffW1Hi + kW1;
constexpr int kOffW2oHi = kOffW1Lo + kW1, kOffW2oLo = kOffW2oHi + kW2o;
constexpr int kOffW2cHi = kOffW2oLo + kW2o, kOffW2cLo = kOffW2cHi + kW2c;
constexpr int kOffW2vHi = kOffW2cLo + kW2c, kOffW2vLo = kOffW2vHi + kW2v;
constexpr int kOffB1 = kOffW2vLo + kW2v;
constexpr int kOffB2o = kOffB1 + kN1, kOffB2c = kOffB2o + kNo, kOffB2v = kOffB2c + kNc;
constexpr int kPacked = kOffB2v + kNv;      // 36160 floats = 144640 B

constexpr int kTileGauss = kRows * kK;       // 1280 (anchor, offset) pairs per tile

// mbarriers
enum { BAR_L1 = 0, BAR_L2O, BAR_L2ALL, BAR_ACCFREE, BAR_STAGEFREE, BAR_COUNT };

struct Smem {
    float w[kPacked];
    // per-tile output staging: Gaussians are written here at their tile-local rank and then copied
    // to HBM as contiguous, fully coalesced segments
    float o_xyz[kTileGauss * 3], o_color[kTileGauss * 3], o_opacity[kTileGauss], o_scaling[kTileGauss * 3];
    float4 o_rot[kTileGauss];
    uint32_t cnt[kGroup];         // kept Gaussians per (row, half), index = row*2 + half
    uint32_t excl[kGroup];
    uint32_t wsum[kGroup / 32];
    uint32_t tile_base[2], tile_total[2];   // by tile parity: FRONT copies tile t out while BACK already scans tile t+1
    uint32_t tmem;
    int timeout;
    alignas(8) uint64_t bar[BAR_COUNT];
};

__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kGroup) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

// FRONT: this thread's share of the layer-1 input of one row.  Loaded one tile AHEAD into registers.
struct FrontInputs {
    float x[32];      // half 0: feat[0..31]; half 1: feat[32..49], view dir (3), distance, 0, 0, ...
    float anchor[3];  // half 1 only (view direction)
    int a;            // source anchor (-1: padding row)
};
// BACK: what the last epilogue needs for the five offsets of this thread.
struct BackInputs {
    float mask[5];
    float off[15];
    float anchor[3];
    float sc[6];
    int a;
};

__device__ __forceinline__ void load_front(FrontInputs &r, int a, int half, const float *__restrict__ anchor,
                                           const float *__restrict__ feat)
{
    r.a = a;
#pragma unroll
    for (int j = 0; j < 32; ++j) r.x[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) r.anchor[i] = 0.f;
    if (a < 0) return;
    const float2 *f2 = reinterpret_cast<const float2 *>(feat + (size_t)a * kFeat);
    if (half == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float2 v = __ldg(f2 + j);
            r.x[2 * j] = v.x;
            r.x[2 * j + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float2 v = __ldg(f2 + 16 + j);
            r.x[2 * j] = v.x;
            r.x[2 * j + 1] = v.y;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) r.anchor[i] = __ldg(anchor + 3 * (size_t)a + i);
    }
}

// view direction / distance (gaussian_renderer/__init__.py:106-110) -- computed when the row is staged
__device__ __forceinline__ void finish_front(FrontInputs &r, int half, float cx, float cy, float cz)
{
    if (half == 1 && r.a >= 0) {
        const float vx = r.anchor[0] - cx, vy = r.anchor[1] - cy, vz = r.anchor[2] - cz;
        const float d = sqrtf(vx * vx + vy * vy + vz * vz);
        r.x[18] = vx / d; r.x[19] = vy / d; r.x[20] = vz / d; r.x[21] = d;
    }
}

__device__ __forceinline__ void load_back(BackInputs &r, int a, int half, const float *__restrict__ anchor,
                                          const float *__restrict__ offsets, const float *__restrict__ scaling,
                                          const float *__restrict__ mask)
{
    r.a = a;
#pragma unroll
    for (int j = 0; j < 5; ++j) r.mask[j] = 0.f;
#pragma unroll
    for (int j = 0; j < 15; ++j) r.off[j] = 0.f;
    if (a < 0) return;
#pragma unroll
    for (int i = 0; i < 3; ++i) r.anchor[i] = __ldg(anchor + 3 * (size_t)a + i);
    {
        const float2 *s2 = reinterpret_cast<const float2 *>(scaling + 6 * (size_t)a);   // 24-byte rows: 8-byte aligned
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float2 v = __ldg(s2 + i);
            r.sc[2 * i] = v.x;
            r.sc[2 * i + 1] = v.y;
        }
    }
    const int kbase = 5 * half;
    const float *mp = mask + (size_t)a * kK + kbase;               // +0 / +20 B: 4-byte aligned in general
    const float *op = offsets + ((size_t)a * kK + kbase) * 3;      // +0 / +60 B
#pragma unroll
    for (int j = 0; j < 5; ++j) r.mask[j] = __ldg(mp + j);
#pragma unroll
    for (int j = 0; j < 15; ++j) r.off[j] = __ldg(op + j);
}

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// tanh(x) = 1 - 2 / (exp(2x) + 1): absolute error ~1e-7 (the selection `tanh(x) * mask > 0` can only
// flip where |x| is at rounding level, as for any fp32 evaluation of the MLP)
__device__ __forceinline__ float fast_tanh(float x)
{
    const float e = __expf(2.0f * fminf(fmaxf(x, -15.0f), 15.0f));
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// What the training-mode forward leaves behind for the tcgen05 backward (neural_gaussians_bwd_umma.cu), per VISIBLE
// row r (position in the visible list): all NULL in inference.
struct SaveOut {
    float *h;             // [Nv][176] hidden activations relu(D1 + b1) in the kernel's column layout (head h at 56h)
    uint32_t *hmask;      // [Nv][6]   bit (8c + j) of word group `half`: hidden unit 88*half + 8c + j is positive
    float *pre2;          // [Nv][144] layer-2 pre-activations incl. bias: opacity 16 | colour 48 | covariance 80
    uint32_t *rowpos;     // [Nv][2]   tile-local rank of the first kept Gaussian of (row, half)
    uint32_t *tilebase;   // [tiles]   global rank of the tile's first Gaussian
};

// hidden = relu(D1 + b1) of 8 accumulator columns, split into TF32 hi / lo, written back to TMEM
template <bool kSave>
__device__ __forceinline__ uint32_t relu_split_store(const Smem &S, uint32_t tl, uint32_t col, const uint32_t (&v)[8],
                                                     float *save_row)
{
    uint32_t hi[8], lo[8];
    float h[8];
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        h[j] = fmaxf(__uint_as_float(v[j]) + S.w[kOffB1 + col + j], 0.f);
        umma::split_tf32(h[j], hi[j], lo[j]);
        if (kSave) bits |= h[j] > 0.f ? (1u << j) : 0u;
    }
    umma::tmem_st8(tl + kColD1 + col, hi);
    umma::tmem_st8(tl + kColHLo + col, lo);
    if (kSave && save_row) {
        reinterpret_cast<float4 *>(save_row + col)[0] = make_float4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<float4 *>(save_row + col)[1] = make_float4(h[4], h[5], h[6], h[7]);
    }
    return bits;
}
}  // namespace ngu

template <bool kSave>
__global__ void __launch_bounds__(ngu::kThreads, 1)
neural_gaussians_umma_kernel(const float *__restrict__ packed_w, const int *__restrict__ vis_idx, int Nv_cap,
                             const int *__restrict__ nv_dev, uint32_t out_cap, const float *__restrict__ anchor, const float *__restrict__ feat,
                             const float *__restrict__ offsets, const float *__restrict__ scaling,
                             const float *__restrict__ mask, float cx, float cy, float cz,
                             float *__restrict__ o_xyz, float *__restrict__ o_color, float *__restrict__ o_opacity,
                             float *__restrict__ o_scaling, float *__restrict__ o_rot,
                             float *__restrict__ o_neural_opacity, uint8_t *__restrict__ o_mask,
                             unsigned long long *scan_state, uint32_t *ctrl, int32_t *__restrict__ count_out,
                             ngu::SaveOut save)
{
    using namespace ngu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool front = tid < kGroup;
    const int uwarp = umma::uniform_warp();                   // warp index the compiler knows to be warp-uniform
    const int gtid = tid & (kGroup - 1), gwarp = gtid >> 5;   // inside the group
    const int half = gwarp >> 2;
    const int row = 32 * (warp & 3) + lane;                   // warp % 4 == gwarp % 4: the lane quadrant this warp may touch
    // the number of visible anchors may live on the device (no host read-back between the stages)
    const int Nv = nv_dev ? min(max(__ldg(nv_dev), 0), Nv_cap) : Nv_cap;
    const int num_tiles = (Nv + kRows - 1) / kRows;
    if (num_tiles == 0) {
        if (blockIdx.x == 0 && tid == 0) *count_out = 0;
        return;
    }

    if (warp == 0) umma::tmem_alloc(&S.tmem, kTmemCols);
    if (tid == 0) {
        for (int i = 0; i < BAR_COUNT; ++i) umma::mbar_init(&S.bar[i], i == BAR_L2ALL ? 3 : 1);   // three issuing threads commit layer 2
        umma::fence_mbar_init();
        S.timeout = 0;
    }
    {
        const float4 *s4 = reinterpret_cast<const float4 *>(packed_w);
        float4 *d4 = reinterpret_cast<float4 *>(S.w);
        for (int i = tid; i < kPacked / 4; i += kThreads) d4[i] = __ldg(s4 + i);
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = S.tmem;
    const uint32_t tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16);  // this warp's lane quadrant
    // Static round-robin tile order: the grid never exceeds the SM count and a CTA needs a whole SM
    // (222 KB shared memory, all 512 TMEM columns), so every CTA is resident and the tiles of one
    // round run concurrently -- the look-back predecessor of a tile is at most one round behind.
    const int stride = (int)gridDim.x;

    auto source_of = [&](int t) -> int {
        const int g = t * kRows + row;
        if (t >= num_tiles || g >= Nv) return -1;
        return vis_idx ? __ldg(vis_idx + g) : g;
    };

    if (front) {
        // =============================== FRONT: input staging, layer 1, ReLU epilogue, MMA issue ===============
        int tile = blockIdx.x;
        FrontInputs cur;
        load_front(cur, source_of(tile), half, anchor, feat);
        int a_next = source_of(tile + stride);
        // coalesced copy-out of a staged tile (its base / size were parked by tile parity)
        auto copy_out = [&](uint32_t par) {
            const size_t base = S.tile_base[par];
            // Gaussians beyond the output capacity are dropped (count_out still reports the true total)
            const uint32_t n = base >= out_cap ? 0u : min(S.tile_total[par], (uint32_t)(out_cap - base));
            for (uint32_t i = gtid; i < 3 * n; i += kGroup) {
                o_xyz[3 * base + i] = S.o_xyz[i];
                o_color[3 * base + i] = S.o_color[i];
                o_scaling[3 * base + i] = S.o_scaling[i];
            }
            for (uint32_t i = gtid; i < n; i += kGroup) {
                o_opacity[base + i] = S.o_opacity[i];
                reinterpret_cast<float4 *>(o_rot)[base + i] = S.o_rot[i];
            }
        };
        // exclusive prefix of a tile whose count BACK has published: one FRONT warp walks the chain (the predecessors
        // published long ago), everybody else picks the result up after the group barrier
        auto resolve_base = [&](int t, uint32_t par) {
            if (gwarp == 0) {
                const uint32_t total = S.tile_total[par];
                const uint64_t excl = lookback_walk(scan_state, t, total);
                if (lane == 0) {
                    S.tile_base[par] = (uint32_t)excl;
                    if (kSave) save.tilebase[t] = (uint32_t)excl;
                    if (t == num_tiles - 1) *count_out = (int32_t)(excl + total);
                }
            }
            group_sync(1);
        };
        uint32_t last_it = 0;
        bool any = false;
        for (uint32_t it = 0; tile < num_tiles; ++it, tile += stride) {
            const uint32_t parity = it & 1u;
            any = true;
            // columns [0,352) are free once the layer-2 MMAs of the previous tile have read the hidden activations
            if (it > 0) {
                if (!umma::mbar_wait(&S.bar[BAR_L2ALL], parity ^ 1u)) S.timeout = 1;
                umma::fence_after_thread_sync();
            }
            // ---- stage the layer-1 input row: half 0 -> k in [0,32), half 1 -> k in [32,56) ---------
            finish_front(cur, half, cx, cy, cz);
            {
                const uint32_t k0 = half == 0 ? 0u : 32u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < 3 || half == 0) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) umma::split_tf32(cur.x[8 * c + j], hi[j], lo[j]);
                        umma::tmem_st8(tl + kColXHi + k0 + 8 * c, hi);
                        umma::tmem_st8(tl + kColXLo + k0 + 8 * c, lo);
                    }
                }
            }
            umma::tmem_wait_st();
            umma::fence_before_thread_sync();
            group_sync(1);
            // ---- layer 1: D1[128 x 176] = x[128 x 56] * W1^T ----------------------------------------
            if (uwarp == 0) {      // warp-uniform branch + elect: back-to-back tcgen05.mma (see umma.cuh)
                if (umma::elect_one_sync()) {
                    umma::fence_after_thread_sync();
                    umma::gemm_3xtf32(tbase + kColD1, tbase + kColXHi, tbase + kColXLo, S.w + kOffW1Hi, S.w + kOffW1Lo, kN1,
                                      kK1, true);
                    umma::umma_commit(&S.bar[BAR_L1]);
                }
                __syncwarp();
            }
            // while the tensor core works: the next tile's input rows start travelling (index fetched a tile earlier)
            FrontInputs nxt;
            load_front(nxt, a_next, half, anchor, feat);
            a_next = source_of(tile + 2 * stride);
            if (!umma::mbar_wait(&S.bar[BAR_L1], parity)) S.timeout = 1;
            umma::fence_after_thread_sync();
            // ---- epilogue 1: hidden = relu(D1 + b1), split, back into TMEM (cols 88*half .. +88) ------
            // software pipelined: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed
            {
                const uint32_t col0 = (uint32_t)(88 * half);
                const int grow = tile * kRows + row;
                float *save_row = (kSave && grow < Nv) ? save.h + (size_t)grow * kN1 : nullptr;
                uint32_t hm[3] = {0u, 0u, 0u};
                uint32_t va[8], vb[8];
                umma::tmem_ld8(tl + kColD1 + col0, va);
                umma::tmem_wait_ld8(va);
#pragma unroll
                for (int c = 0; c < 10; c += 2) {
                    umma::tmem_ld8(tl + kColD1 + col0 + 8 * (c + 1), vb);
                    hm[(8 * c) >> 5] |= relu_split_store<kSave>(S, tl, col0 + 8 * c, va, save_row) << ((8 * c) & 31);
                    umma::tmem_wait_ld8(vb);
                    umma::tmem_ld8(tl + kColD1 + col0 + 8 * (c + 2), va);
                    hm[(8 * (c + 1)) >> 5] |= relu_split_store<kSave>(S, tl, col0 + 8 * (c + 1), vb, save_row) << ((8 * (c + 1)) & 31);
                    umma::tmem_wait_ld8(va);
                }
                hm[2] |= relu_split_store<kSave>(S, tl, col0 + 80, va, save_row) << 16;
                if (kSave && save_row) {
                    uint32_t *m = save.hmask + (size_t)grow * 6 + 3 * half;
                    m[0] = hm[0]; m[1] = hm[1]; m[2] = hm[2];
                }
            }
            umma::tmem_wait_st();
            umma::fence_before_thread_sync();
            group_sync(1);
            // ---- layer 2: opacity head first (it decides the selection), then colour and covariance ----
            if (it > 0) {   // BACK has finished reading the accumulators of the previous tile and filled the staging area
                if (!umma::mbar_wait(&S.bar[BAR_ACCFREE], parity ^ 1u)) S.timeout = 1;
                umma::fence_after_thread_sync();
            }
            // The three heads are independent GEMMs.  One thread issues a tcgen05.mma every ~95 cycles whatever its shape, and
            // MMAs issued by different warps overlap (scripts/umma_rate_probe.py): three warps issue one head each, 21 MMAs
            // per warp instead of 63 from one thread.  Every issuing thread commits its own MMAs (BAR_L2ALL counts three).
            if (uwarp < 3 && umma::elect_one_sync()) {
                umma::fence_after_thread_sync();
                if (uwarp == 0) {
                    umma::gemm_3xtf32(tbase + kColDo, tbase + kColD1 + 0 * kHeadStride, tbase + kColHLo + 0 * kHeadStride,
                                      S.w + kOffW2oHi, S.w + kOffW2oLo, kNo, kK1, true);
                    umma::umma_commit(&S.bar[BAR_L2O]);
                } else if (uwarp == 1) {
                    umma::gemm_3xtf32(tbase + kColDc, tbase + kColD1 + 1 * kHeadStride, tbase + kColHLo + 1 * kHeadStride,
                                      S.w + kOffW2cHi, S.w + kOffW2cLo, kNc, kK1, true);
                } else {
                    umma::gemm_3xtf32(tbase + kColDv, tbase + kColD1 + 2 * kHeadStride, tbase + kColHLo + 2 * kHeadStride,
                                      S.w + kOffW2vHi, S.w + kOffW2vLo, kNv, kK1, true);
                }
                umma::umma_commit(&S.bar[BAR_L2ALL]);
            }
            if (it > 0) {   // copy the previous tile out while the tensor core and BACK work on
                resolve_base(tile - stride, (it - 1) & 1u);
                copy_out((it - 1) & 1u);
                group_sync(1);
                if (gtid == 0) mbar_arrive(&S.bar[BAR_STAGEFREE]);
            }
            last_it = it;
            cur = nxt;
        }
        if (any) {   // the last tile of this CTA
            if (!umma::mbar_wait(&S.bar[BAR_ACCFREE], last_it & 1u)) S.timeout = 1;
            resolve_base(tile - stride, last_it & 1u);
            copy_out(last_it & 1u);
        }
    } else {
        // =============================== BACK: selection, ordered compaction, post-processing, copy-out ==========
        int tile = blockIdx.x;
        BackInputs cur;
        load_back(cur, source_of(tile), half, anchor, offsets, scaling, mask);
        int a_next = source_of(tile + stride);
        const int kbase = 5 * half;
        for (uint32_t it = 0; tile < num_tiles; ++it, tile += stride) {
            const uint32_t parity = it & 1u;
            const int a = cur.a;
            // the next tile's rows travel while this tile is post-processed
            BackInputs nxt;
            load_back(nxt, a_next, half, anchor, offsets, scaling, mask);
            a_next = source_of(tile + 2 * stride);
            if (!umma::mbar_wait(&S.bar[BAR_L2O], parity)) S.timeout = 1;
            umma::fence_after_thread_sync();
            // ---- epilogue 2a: selection of offsets k = 5*half + j, ordered ranks -------------------------
            float nop[5];
            uint32_t keepbits = 0;
            {
                uint32_t v[8];
                umma::tmem_ld8(tl + kColDo + 8 * half, v);
                umma::tmem_wait_ld8(v);
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    nop[j] = 0.f;
                    if (a >= 0) {
                        nop[j] = fast_tanh(__uint_as_float(v[j]) + S.w[kOffB2o + 8 * half + j]) * cur.mask[j];
                        keepbits |= nop[j] > 0.0f ? (1u << j) : 0u;
                    }
                    if (o_neural_opacity && tile * kRows + row < Nv) {   // training-side outputs (gaussian_renderer/__init__.py:147-148)
                        const size_t gp = ((size_t)tile * kRows + row) * kK + kbase + j;
                        o_neural_opacity[gp] = nop[j];
                        o_mask[gp] = (keepbits >> j) & 1u;
                    }
                }
                if (kSave && tile * kRows + row < Nv) {
                    float p[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) p[j] = j < 5 ? __uint_as_float(v[j]) + S.w[kOffB2o + 8 * half + j] : 0.f;
                    float4 *dst = reinterpret_cast<float4 *>(save.pre2 + ((size_t)tile * kRows + row) * 144 + 8 * half);
                    dst[0] = make_float4(p[0], p[1], p[2], p[3]);
                    dst[1] = make_float4(p[4], p[5], p[6], p[7]);
                }
            }
            S.cnt[row * 2 + half] = __popc(keepbits);  // order index = row*2 + half
            group_sync(2);
            uint32_t total = 0;
            {
                const uint32_t v = S.cnt[gtid];
                uint32_t incl = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane == 31) S.wsum[gwarp] = incl;
                group_sync(2);
                uint32_t before = 0;
#pragma unroll
                for (int w = 0; w < kGroup / 32; ++w) {
                    const uint32_t c = S.wsum[w];
                    before += w < gwarp ? c : 0u;
                    total += c;
                }
                S.excl[gtid] = before + incl - v;
            }
            group_sync(2);  // tile-local ranks complete
            if (gtid == 0) {
                // publish this tile's count right away; the prefix is resolved later by FRONT, when the copy-out
                // needs it -- BACK never waits for another CTA
                S.tile_total[parity] = total;
                volatile unsigned long long *st = scan_state;
                st[tile] = (tile == 0 ? kLbInclusive : kLbAggregate) | (unsigned long long)total;
            }
            uint32_t pos = S.excl[row * 2 + half];
            if (kSave && tile * kRows + row < Nv) save.rowpos[((size_t)tile * kRows + row) * 2 + half] = pos;

            // ---- epilogue 2b: post-process the kept offsets into the staging buffers -----------------------
            if (!umma::mbar_wait(&S.bar[BAR_L2ALL], parity)) S.timeout = 1;
            umma::fence_after_thread_sync();
            if (it > 0 && !umma::mbar_wait(&S.bar[BAR_STAGEFREE], parity ^ 1u)) S.timeout = 1;   // FRONT has copied tile t-1 out
            {
                // tcgen05.ld of offset j+1 is in flight while offset j is post-processed
                uint32_t vc[2][4], vv[2][8];
                umma::tmem_ld4(tl + kColDc + 4 * kbase, vc[0]);
                umma::tmem_ld8(tl + kColDv + 8 * kbase, vv[0]);
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const int k = kbase + j;
                    const int b = j & 1;
                    umma::tmem_wait_ld4(vc[b]);
                    umma::tmem_wait_ld8(vv[b]);
                    if (j + 1 < 5) {
                        umma::tmem_ld4(tl + kColDc + 4 * (k + 1), vc[b ^ 1]);
                        umma::tmem_ld8(tl + kColDv + 8 * (k + 1), vv[b ^ 1]);
                    }
                    if (kSave && tile * kRows + row < Nv) {
                        float *dst = save.pre2 + ((size_t)tile * kRows + row) * 144;
                        reinterpret_cast<float4 *>(dst + 16 + 4 * k)[0] =
                            make_float4(__uint_as_float(vc[b][0]) + S.w[kOffB2c + 4 * k + 0],
                                        __uint_as_float(vc[b][1]) + S.w[kOffB2c + 4 * k + 1],
                                        __uint_as_float(vc[b][2]) + S.w[kOffB2c + 4 * k + 2], 0.f);
                        reinterpret_cast<float4 *>(dst + 64 + 8 * k)[0] =
                            make_float4(__uint_as_float(vv[b][0]) + S.w[kOffB2v + 8 * k + 0],
                                        __uint_as_float(vv[b][1]) + S.w[kOffB2v + 8 * k + 1],
                                        __uint_as_float(vv[b][2]) + S.w[kOffB2v + 8 * k + 2],
                                        __uint_as_float(vv[b][3]) + S.w[kOffB2v + 8 * k + 3]);
                        reinterpret_cast<float4 *>(dst + 64 + 8 * k)[1] =
                            make_float4(__uint_as_float(vv[b][4]) + S.w[kOffB2v + 8 * k + 4],
                                        __uint_as_float(vv[b][5]) + S.w[kOffB2v + 8 * k + 5],
                                        __uint_as_float(vv[b][6]) + S.w[kOffB2v + 8 * k + 6], 0.f);
                    }
                    if (keepbits & (1u << j)) {
                        const uint32_t p = pos++;
                        S.o_xyz[3 * p + 0] = cur.anchor[0] + cur.off[3 * j + 0] * cur.sc[0];
                        S.o_xyz[3 * p + 1] = cur.anchor[1] + cur.off[3 * j + 1] * cur.sc[1];
                        S.o_xyz[3 * p + 2] = cur.anchor[2] + cur.off[3 * j + 2] * cur.sc[2];
                        S.o_color[3 * p + 0] = fast_sigmoid(__uint_as_float(vc[b][0]) + S.w[kOffB2c + 4 * k + 0]);
                        S.o_color[3 * p + 1] = fast_sigmoid(__uint_as_float(vc[b][1]) + S.w[kOffB2c + 4 * k + 1]);
                        S.o_color[3 * p + 2] = fast_sigmoid(__uint_as_float(vc[b][2]) + S.w[kOffB2c + 4 * k + 2]);
                        S.o_opacity[p] = nop[j];
                        float cv[7];
#pragma unroll
                        for (int i = 0; i < 7; ++i) cv[i] = __uint_as_float(vv[b][i]) + S.w[kOffB2v + 8 * k + i];
                        S.o_scaling[3 * p + 0] = cur.sc[3] * fast_sigmoid(cv[0]);
                        S.o_scaling[3 * p + 1] = cur.sc[4] * fast_sigmoid(cv[1]);
                        S.o_scaling[3 * p + 2] = cur.sc[5] * fast_sigmoid(cv[2]);
                        const float ss = cv[3] * cv[3] + cv[4] * cv[4] + cv[5] * cv[5] + cv[6] * cv[6];
                        const float inv = rsqrtf(fmaxf(ss, 1e-24f));  // F.normalize: v / max(|v|, 1e-12)
                        S.o_rot[p] = make_float4(cv[3] * inv, cv[4] * inv, cv[5] * inv, cv[6] * inv);
                    }
                }
            }
            umma::fence_before_thread_sync();
            group_sync(2);  // staging complete, tile_base published, all TMEM reads of this tile done
            if (gtid == 0) mbar_arrive(&S.bar[BAR_ACCFREE]);   // FRONT may overwrite the layer-2 accumulators and copy the tile out
            cur = nxt;
        }
    }

    umma::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
        // a tensor-core completion that timed out poisons the result: the LAST CTA to leave reports
        // it through the count (-1), which the host reads anyway
        if (S.timeout) atomicExch(&ctrl[1], 1u);
        __threadfence();
        if (atomicAdd(&ctrl[2], 1u) == gridDim.x - 1 && atomicAdd(&ctrl[1], 0u) != 0u) *count_out = -1;
    }
    if (warp == 0) umma::tmem_dealloc(tbase, kTmemCols);
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_neural_gaussians_umma_packed_floats(void) { return ngu::kPacked; }

extern "C" size_t cgs_neural_gaussians_umma_workspace_bytes(int Nv)
{
    const size_t tiles = (size_t)(Nv > 0 ? (Nv + ngu::kRows - 1) / ngu::kRows : 1);
    return align_up(tiles * 8) + align_up(16);
}

extern "C" int cgs_neural_gaussians_umma_forward(const float *packed_weights, const int32_t *vis_idx, int Nv,
                                                 const float *anchor, const float *feat, const float *offsets,
                                                 const float *scaling, const float *mask, const float *campos_host,
                                                 float *o_xyz, float *o_color, float *o_opacity, float *o_scaling,
                                                 float *o_rot, float *o_neural_opacity, uint8_t *o_mask,
                                                 int32_t *count_dev, void *workspace, size_t workspace_bytes,
                                                 void *stream)
{
    if (Nv > 0) {
        CGS_CHECK_PTR(o_neural_opacity);
        CGS_CHECK_PTR(o_mask);
    }
    return cgs_neural_gaussians_umma_forward_dev(packed_weights, vis_idx, Nv, nullptr, (int64_t)Nv * ngu::kK, anchor, feat,
                                                 offsets, scaling, mask, campos_host, o_xyz, o_color, o_opacity, o_scaling,
                                                 o_rot, o_neural_opacity, o_mask, count_dev, workspace, workspace_bytes,
                                                 stream);
}

static int launch_g1(const char *func, const float *packed_weights, const int32_t *vis_idx, int Nv, const int32_t *nv_dev,
                     int64_t out_cap, const float *anchor, const float *feat, const float *offsets, const float *scaling,
                     const float *mask, const float *campos_host, float *o_xyz, float *o_color, float *o_opacity,
                     float *o_scaling, float *o_rot, float *o_neural_opacity, uint8_t *o_mask, int32_t *count_dev,
                     void *workspace, size_t workspace_bytes, void *stream, const ngu::SaveOut &save);

extern "C" int cgs_neural_gaussians_umma_forward_dev(const float *packed_weights, const int32_t *vis_idx, int Nv,
                                                     const int32_t *nv_dev, int64_t out_cap, const float *anchor,
                                                     const float *feat, const float *offsets, const float *scaling,
                                                     const float *mask, const float *campos_host, float *o_xyz,
                                                     float *o_color, float *o_opacity, float *o_scaling, float *o_rot,
                                                     float *o_neural_opacity, uint8_t *o_mask, int32_t *count_dev,
                                                     void *workspace, size_t workspace_bytes, void *stream)
{
    return launch_g1(__func__, packed_weights, vis_idx, Nv, nv_dev, out_cap, anchor, feat, offsets, scaling, mask,
                     campos_host, o_xyz, o_color, o_opacity, o_scaling, o_rot, o_neural_opacity, o_mask, count_dev, workspace,
                     workspace_bytes, stream, ngu::SaveOut{nullptr, nullptr, nullptr, nullptr, nullptr});
}

extern "C" int cgs_neural_gaussians_umma_forward_train(const float *packed_weights, const int32_t *vis_idx, int Nv,
                                                       const float *anchor, const float *feat, const float *offsets,
                                                       const float *scaling, const float *mask, const float *campos_host,
                                                       float *o_xyz, float *o_color, float *o_opacity, float *o_scaling,
                                                       float *o_rot, float *o_neural_opacity, uint8_t *o_mask,
                                                       int32_t *count_dev, float *save_h, uint32_t *save_hmask,
                                                       float *save_pre2, uint32_t *save_rowpos, uint32_t *save_tilebase,
                                                       void *workspace, size_t workspace_bytes, void *stream)
{
    if (Nv > 0) {
        CGS_CHECK_PTR(o_neural_opacity); CGS_CHECK_PTR(o_mask); CGS_CHECK_PTR(save_h); CGS_CHECK_PTR(save_hmask);
        CGS_CHECK_PTR(save_pre2); CGS_CHECK_PTR(save_rowpos); CGS_CHECK_PTR(save_tilebase);
    }
    return launch_g1(__func__, packed_weights, vis_idx, Nv, nullptr, (int64_t)Nv * ngu::kK, anchor, feat, offsets, scaling,
                     mask, campos_host, o_xyz, o_color, o_opacity, o_scaling, o_rot, o_neural_opacity, o_mask, count_dev,
                     workspace, workspace_bytes, stream,
                     ngu::SaveOut{save_h, save_hmask, save_pre2, save_rowpos, save_tilebase});
}

static int launch_g1(const char *func, const float *packed_weights, const int32_t *vis_idx, int Nv, const int32_t *nv_dev,
                     int64_t out_cap, const float *anchor, const float *feat, const float *offsets, const float *scaling,
                     const float *mask, const float *campos_host, float *o_xyz, float *o_color, float *o_opacity,
                     float *o_scaling, float *o_rot, float *o_neural_opacity, uint8_t *o_mask, int32_t *count_dev,
                     void *workspace, size_t workspace_bytes, void *stream, const ngu::SaveOut &save)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CGS_CHECK_PTR(count_dev);
    if (Nv <= 0) {
        cudaMemsetAsync(count_dev, 0, sizeof(int32_t), st);
        return check_launch(__func__);
    }
    CGS_CHECK_PTR(packed_weights);
    CGS_CHECK_PTR(anchor);
    CGS_CHECK_PTR(feat);
    CGS_CHECK_PTR(offsets);
    CGS_CHECK_PTR(scaling);
    CGS_CHECK_PTR(mask);
    CGS_CHECK_PTR(campos_host);
    CGS_CHECK_PTR(o_xyz);
    CGS_CHECK_PTR(o_color);
    CGS_CHECK_PTR(o_opacity);
    CGS_CHECK_PTR(o_scaling);
    CGS_CHECK_PTR(o_rot);
    CGS_CHECK_PTR(workspace);
    if ((o_neural_opacity == nullptr) != (o_mask == nullptr)) {
        set_error("%s: o_neural_opacity and o_mask must both be given or both be NULL", __func__);
        return -2;
    }
    if (out_cap < 0 || out_cap > 0x7fffffffll) {
        set_error("%s: invalid output capacity", __func__);
        return -2;
    }
    if (workspace_bytes < cgs_neural_gaussians_umma_workspace_bytes(Nv)) {
        set_error("%s: workspace too small", __func__);
        return -3;
    }
    const int tiles = (Nv + ngu::kRows - 1) / ngu::kRows;
    char *ws = static_cast<char *>(workspace);
    unsigned long long *scan_state = reinterpret_cast<unsigned long long *>(ws);
    uint32_t *ctrl = reinterpret_cast<uint32_t *>(ws + align_up((size_t)tiles * 8));  // ticket, error, finished
    cudaMemsetAsync(ws, 0, align_up((size_t)tiles * 8) + align_up(16), st);
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(neural_gaussians_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(ngu::Smem));
        cudaFuncSetAttribute(neural_gaussians_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(ngu::Smem));
        if (sm_count <= 0) sm_count = kNumSMs;
    }
    const int grid = tiles < sm_count ? tiles : sm_count;
    StageScope sc(ST_G1_FWD, st, 1);
    if (save.h)
        neural_gaussians_umma_kernel<true><<<grid, ngu::kThreads, sizeof(ngu::Smem), st>>>(
            packed_weights, vis_idx, Nv, nv_dev, (uint32_t)out_cap, anchor, feat, offsets, scaling, mask, campos_host[0],
            campos_host[1], campos_host[2], o_xyz, o_color, o_opacity, o_scaling, o_rot, o_neural_opacity, o_mask,
            scan_state, ctrl, count_dev, save);
    else
        neural_gaussians_umma_kernel<false><<<grid, ngu::kThreads, sizeof(ngu::Smem), st>>>(
            packed_weights, vis_idx, Nv, nv_dev, (uint32_t)out_cap, anchor, feat, offsets, scaling, mask, campos_host[0],
            campos_host[1], campos_host[2], o_xyz, o_color, o_opacity, o_scaling, o_rot, o_neural_opacity, o_mask,
            scan_state, ctrl, count_dev, save);
    return check_launch(func);
}

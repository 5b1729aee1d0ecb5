// Backward of the fused anchor -> neural-Gaussian generation on the tcgen05 tensor cores.
//
// Replaces what autograd does for gaussian_renderer/__init__.py:106-145 and the decoder MLPs
// (scene/gaussian_model.py:153-174), like neural_gaussians_bwd.cu, but with every GEMM on the tensor cores and
// NOTHING recomputed: the training-mode forward (neural_gaussians_umma.cu, kSave) leaves the hidden activations,
// their sign bits, the layer-2 pre-activations and the compaction ranks behind (1.3 kB per visible anchor).
//
//   kernel 1  neural_gaussians_dgrad_umma_kernel   (data gradients; persistent, 128 anchors per tile, 640 threads:
//             five threads per row, each owning two of the ten offsets; index-type inputs are fetched one tile ahead and
//             every other load of a tile is issued before the first use -- the first version, 256 threads with loads
//             inside the per-offset branches, sat on the DRAM latency with 12 % of the issue slots busy)
//       E1  gradient of the post-processing per kept (anchor, offset) pair -> dOut[128 x 144] (layer-2 output layout of
//           the forward: opacity 16 | colour 48 | covariance 80), split hi / lo into TMEM, fp32 copy to HBM for kernel 2;
//           direct gradients d_offsets, d_mask, partial d_scaling / d_anchor
//       M1  dH = dOut W2 per head: 3xTF32 tcgen05.mma, A = dOut (TMEM), B = W2^T head (shared memory, K-major)
//       E2  dPre = dH * (H > 0) (saved sign bits), split, IN PLACE as the next A operand; fp32 copy to HBM
//       M2  dX = dPre W1 (K = 176, N = 64)
//       E3  d_feat, d_anchor (view direction / distance), d_scaling
//   kernel 2  neural_gaussians_wgrad_umma_kernel   (weight gradients; the contraction runs over the ROWS)
//       dW2_h^T = dOut_h^T H_h and dW1^T = X^T dPre as SS-form tcgen05.mma with the ROW index playing K: both operands
//       in the K-major no-swizzle core-matrix layout [row / 4][feature][4 rows] (leading dimension = 1 mod 8 features:
//       conflict-free scalar stores), 3xTF32, accumulators resident in TMEM over all slabs of a persistent CTA, one
//       flush (atomicAdd) per CTA.  (The MN-major descriptor form returned zeros in a probe on B200 and is not used.)
//       Bias gradients ride along: a constant-one column in the padding of X (-> db1) and of every head of H (-> db2).
//
// TMEM columns of kernel 1 (480 of 512):
//   [  0,288)  dOut hi [0,144) | lo [144,288)   -> later dPre lo [0,192) and the dX accumulator [192,256)
//   [288,480)  dH accumulator, head h at 64h (N = 64 per head: 50 units + zero pads) -> overwritten in place by dPre hi
#include "umma.cuh"

namespace cgs {
namespace ngbu {
constexpr int kFeat = 50, kK = 10;
constexpr int kRows = 128, kThreads = 640;   // five threads per row
constexpr int kOutP = 144, kHidP = 176, kInP = 64;
constexpr int kHidT = 192;                                  // hidden columns in TMEM: head h at 64h (HBM rows: 56h)
constexpr int kKo = 16, kKc = 48, kKv = 80;                 // K of the three dH GEMMs (= padded head widths)
constexpr uint32_t kColAHi = 0, kColALo = 144, kColPLo = 0, kColDX = 192, kColD1 = 288;
constexpr uint32_t kTmemCols = 512;
// packed transposed weights (floats): B operands [K/4][64][4], hi then lo
constexpr int kW2To = (kKo / 4) * 64 * 4, kW2Tc = (kKc / 4) * 64 * 4, kW2Tv = (kKv / 4) * 64 * 4, kW1T = (kHidT / 4) * 64 * 4;
constexpr int kOffW2ToHi = 0, kOffW2ToLo = kOffW2ToHi + kW2To;
constexpr int kOffW2TcHi = kOffW2ToLo + kW2To, kOffW2TcLo = kOffW2TcHi + kW2Tc;
constexpr int kOffW2TvHi = kOffW2TcLo + kW2Tc, kOffW2TvLo = kOffW2TvHi + kW2Tv;
constexpr int kOffW1THi = kOffW2TvLo + kW2Tv, kOffW1TLo = kOffW1THi + kW1T;
constexpr int kPacked = kOffW1TLo + kW1T;                  // 43008 floats = 172032 B

struct Smem {
    float w[kPacked];
    uint32_t tmem;
    int timeout;
    alignas(8) uint64_t bar[2];
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ void st_split8(uint32_t tl, uint32_t col_hi, uint32_t col_lo, const float (&v)[8])
{
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) umma::split_tf32(v[j], hi[j], lo[j]);
    umma::tmem_st8(tl + col_hi, hi);
    umma::tmem_st8(tl + col_lo, lo);
}
__device__ __forceinline__ void st_split4(uint32_t tl, uint32_t col_hi, uint32_t col_lo, const float (&v)[4])
{
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) umma::split_tf32(v[j], hi[j], lo[j]);
    umma::tmem_st4(tl + col_hi, hi);
    umma::tmem_st4(tl + col_lo, lo);
}
}  // namespace ngbu

// kernel 0: gradient of the post-processing, one thread per (visible row, offset) -- plain elementwise work at full
// occupancy (it used to be the first phase of kernel 1, where 20 warps per SM could not cover the DRAM latency of its
// gathers).  Writes the layer-2 output gradient dOut[Nv,144] (opacity 16 | colour 48 | covariance 80, padding columns
// zero), d_offsets, d_mask and accumulates the per-anchor parts of d_scaling / d_anchor (rows zero-filled by the caller).
__global__ void __launch_bounds__(256)
neural_gaussians_postproc_backward_kernel(const int *__restrict__ vis_idx, int Nv, const float *__restrict__ offsets,
                                          const float *__restrict__ scaling, const float *__restrict__ mask,
                                          const uint8_t *__restrict__ keep_mask, const float *__restrict__ save_pre2,
                                          const uint32_t *__restrict__ save_rowpos,
                                          const uint32_t *__restrict__ save_tilebase, const float *__restrict__ g_xyz,
                                          const float *__restrict__ g_color, const float *__restrict__ g_opacity,
                                          const float *__restrict__ g_scaling, const float *__restrict__ g_rot,
                                          float *__restrict__ d_anchor, float *__restrict__ d_offsets,
                                          float *__restrict__ d_scaling, float *__restrict__ d_mask,
                                          float *__restrict__ d_out)
{
    using namespace ngbu;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)Nv * 16) return;
    const int g = (int)(e >> 4), k = (int)(e & 15);          // 16 threads per row: offsets 0..9, 10..15 write the padding
    float *dst = d_out + (size_t)g * kOutP;
    if (k >= kK) {
        // padding columns: opacity 5..7, 13..15 and colour 56..63
        if (k == 10) { dst[5] = 0.f; dst[6] = 0.f; dst[7] = 0.f; }
        if (k == 11) { dst[13] = 0.f; dst[14] = 0.f; dst[15] = 0.f; }
        if (k == 12) *reinterpret_cast<float4 *>(dst + 56) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k == 13) *reinterpret_cast<float4 *>(dst + 60) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const int hk = k >= 5 ? 1 : 0;
    const uint32_t ocol = (uint32_t)(8 * hk + (k - 5 * hk));
    float dO = 0.f;
    float4 dC = make_float4(0.f, 0.f, 0.f, 0.f), dVa = dC, dVb = dC;
    if (keep_mask[(size_t)g * kK + k]) {
        const int a = vis_idx ? __ldg(vis_idx + g) : g;
        uint32_t before = 0;
        for (int i = 5 * hk; i < k; ++i) before += keep_mask[(size_t)g * kK + i] ? 1u : 0u;
        const size_t pos = (size_t)__ldg(save_tilebase + (g >> 7)) + __ldg(save_rowpos + (size_t)g * 2 + hk) + before;
        const size_t ak = (size_t)a * kK + k;
        const float *p2 = save_pre2 + (size_t)g * kOutP;
        const float gx = __ldg(g_xyz + 3 * pos), gy = __ldg(g_xyz + 3 * pos + 1), gz = __ldg(g_xyz + 3 * pos + 2);
        const float o0 = __ldg(offsets + 3 * ak), o1 = __ldg(offsets + 3 * ak + 1), o2 = __ldg(offsets + 3 * ak + 2);
        const float gc0 = __ldg(g_color + 3 * pos), gc1 = __ldg(g_color + 3 * pos + 1), gc2 = __ldg(g_color + 3 * pos + 2);
        const float gs0 = __ldg(g_scaling + 3 * pos), gs1 = __ldg(g_scaling + 3 * pos + 1), gs2 = __ldg(g_scaling + 3 * pos + 2);
        const float go = __ldg(g_opacity + pos), mk = __ldg(mask + ak), po = __ldg(p2 + ocol);
        const float4 pc = __ldg(reinterpret_cast<const float4 *>(p2 + 16 + 4 * k));
        const float4 pa = __ldg(reinterpret_cast<const float4 *>(p2 + 64 + 8 * k));
        const float4 pb = __ldg(reinterpret_cast<const float4 *>(p2 + 64 + 8 * k) + 1);
        const float4 gr = __ldg(reinterpret_cast<const float4 *>(g_rot) + pos);
        float sc[6];
        {
            const float2 *s2 = reinterpret_cast<const float2 *>(scaling + 6 * (size_t)a);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float2 v = __ldg(s2 + i);
                sc[2 * i] = v.x;
                sc[2 * i + 1] = v.y;
            }
        }
        d_offsets[ak * 3 + 0] = gx * sc[0];
        d_offsets[ak * 3 + 1] = gy * sc[1];
        d_offsets[ak * 3 + 2] = gz * sc[2];
        // opacity = tanh(pre) * mask
        const float t = tanhf(po);
        d_mask[ak] = go * t;
        dO = go * mk * (1.0f - t * t);
        // colour = sigmoid(pre)
        const float c0 = sigmoidf_(pc.x), c1 = sigmoidf_(pc.y), c2 = sigmoidf_(pc.z);
        dC = make_float4(gc0 * c0 * (1.0f - c0), gc1 * c1 * (1.0f - c1), gc2 * c2 * (1.0f - c2), 0.f);
        // scaling = sc[3:6] * sigmoid(pre[0:3]); rot = normalize(pre[3:7])
        const float s0 = sigmoidf_(pa.x), s1 = sigmoidf_(pa.y), s2 = sigmoidf_(pa.z);
        const float q0 = pa.w, q1 = pb.x, q2 = pb.y, q3 = pb.z;
        const float nrm = fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);
        const float r0 = q0 / nrm, r1 = q1 / nrm, r2 = q2 / nrm, r3 = q3 / nrm;
        const float dot = r0 * gr.x + r1 * gr.y + r2 * gr.z + r3 * gr.w;
        dVa = make_float4(gs0 * sc[3] * s0 * (1.0f - s0), gs1 * sc[4] * s1 * (1.0f - s1), gs2 * sc[5] * s2 * (1.0f - s2),
                          (gr.x - r0 * dot) / nrm);
        dVb = make_float4((gr.y - r1 * dot) / nrm, (gr.z - r2 * dot) / nrm, (gr.w - r3 * dot) / nrm, 0.f);
        // per-anchor sums over the offsets (ten threads of one row: distinct addresses across rows, light contention)
        atomicAdd(d_anchor + 3 * (size_t)a + 0, gx);
        atomicAdd(d_anchor + 3 * (size_t)a + 1, gy);
        atomicAdd(d_anchor + 3 * (size_t)a + 2, gz);
        atomicAdd(d_scaling + 6 * (size_t)a + 0, gx * o0);
        atomicAdd(d_scaling + 6 * (size_t)a + 1, gy * o1);
        atomicAdd(d_scaling + 6 * (size_t)a + 2, gz * o2);
        atomicAdd(d_scaling + 6 * (size_t)a + 3, gs0 * s0);
        atomicAdd(d_scaling + 6 * (size_t)a + 4, gs1 * s1);
        atomicAdd(d_scaling + 6 * (size_t)a + 5, gs2 * s2);
    }
    dst[ocol] = dO;
    *reinterpret_cast<float4 *>(dst + 16 + 4 * k) = dC;
    *reinterpret_cast<float4 *>(dst + 64 + 8 * k) = dVa;
    *(reinterpret_cast<float4 *>(dst + 64 + 8 * k) + 1) = dVb;
}

__global__ void __launch_bounds__(ngbu::kThreads, 1)
neural_gaussians_dgrad_umma_kernel(const float *__restrict__ packed_w, const int *__restrict__ vis_idx, int Nv,
                                   const float *__restrict__ anchor, float cx, float cy, float cz,
                                   const uint32_t *__restrict__ save_hmask, const float *__restrict__ d_out,
                                   float *__restrict__ d_anchor, float *__restrict__ d_feat, float *__restrict__ d_pre,
                                   int32_t *__restrict__ err)
{
    using namespace ngbu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fifth = warp >> 2;                       // five threads per row
    const int row = 32 * (warp & 3) + lane;
    const int uwarp = umma::uniform_warp();
    const int num_tiles = (Nv + kRows - 1) / kRows;

    if (warp == 0) umma::tmem_alloc(&S.tmem, kTmemCols);
    if (tid == 0) {
        umma::mbar_init(&S.bar[0], 3);      // three issuing threads commit the head GEMMs
        umma::mbar_init(&S.bar[1], 1);
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
    const uint32_t tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16);

    // Everything a tile reads from HBM is fetched one tile AHEAD into registers: this thread's share of the dOut row
    // (8-column chunks fifth, fifth + 5, ... of 18) and the sign bits of the hidden layer.
    struct Pre {
        float4 v[4][2];
        uint32_t hm[6];
        int a;
    };
    auto prefetch = [&](int tile, Pre &p) {
        const int g = tile * kRows + row;
        p.a = -1;
#pragma unroll
        for (int i = 0; i < 6; ++i) p.hm[i] = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) p.v[i][0] = p.v[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tile >= num_tiles || g >= Nv) return;
        p.a = vis_idx ? __ldg(vis_idx + g) : g;
        const float4 *src = reinterpret_cast<const float4 *>(d_out + (size_t)g * kOutP);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = fifth + 5 * i;
            if (c < 18) {
                p.v[i][0] = __ldg(src + 2 * c);
                p.v[i][1] = __ldg(src + 2 * c + 1);
            }
        }
        const uint2 *m = reinterpret_cast<const uint2 *>(save_hmask + (size_t)g * 6);
        const uint2 m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2);
        p.hm[0] = m0.x; p.hm[1] = m0.y; p.hm[2] = m1.x; p.hm[3] = m1.y; p.hm[4] = m2.x; p.hm[5] = m2.y;
    };
    Pre nxt;
    prefetch(blockIdx.x, nxt);

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t parity = it & 1u;
        const int g = tile * kRows + row;
        const Pre cur = nxt;
        const bool valid = cur.a >= 0;
        prefetch(tile + (int)gridDim.x, nxt);

        // ================= stage dOut (kernel 0) as the A operand: hi / lo into TMEM =================
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = fifth + 5 * i;
            if (c < 18) {
                const float f[8] = {cur.v[i][0].x, cur.v[i][0].y, cur.v[i][0].z, cur.v[i][0].w,
                                    cur.v[i][1].x, cur.v[i][1].y, cur.v[i][1].z, cur.v[i][1].w};
                st_split8(tl, kColAHi + 8 * c, kColALo + 8 * c, f);
            }
        }
        umma::tmem_wait_st();
        umma::fence_before_thread_sync();
        __syncthreads();

        // ================= M1: dH_h = dOut_h W2_h =================
        // three independent heads, one issuing warp each (MMAs of different warps overlap; one thread issues one every ~95
        // cycles): each issuing thread commits its own MMAs, bar[0] counts three arrivals
        if (uwarp < 3 && umma::elect_one_sync()) {
            umma::fence_after_thread_sync();
            if (uwarp == 0)
                umma::gemm_3xtf32(tbase + kColD1 + 0, tbase + kColAHi + 0, tbase + kColALo + 0, S.w + kOffW2ToHi,
                                  S.w + kOffW2ToLo, 64, kKo, true);
            else if (uwarp == 1)
                umma::gemm_3xtf32(tbase + kColD1 + 64, tbase + kColAHi + 16, tbase + kColALo + 16, S.w + kOffW2TcHi,
                                  S.w + kOffW2TcLo, 64, kKc, true);
            else
                umma::gemm_3xtf32(tbase + kColD1 + 128, tbase + kColAHi + 64, tbase + kColALo + 64, S.w + kOffW2TvHi,
                                  S.w + kOffW2TvLo, 64, kKv, true);
            umma::umma_commit(&S.bar[0]);
        }
        if (!umma::mbar_wait(&S.bar[0], parity)) S.timeout = 1;
        umma::fence_after_thread_sync();

        // ================= E2: dPre = dH * (H > 0) =================
        // TMEM chunk q (8 columns at 64h + 8cc) <-> hidden columns 56h + 8cc .. +7 of the saved layout (cc = 7: padding)
#pragma unroll
        for (int qq = 0; qq < 5; ++qq) {
            const int q = fifth + 5 * qq;
            if (q < 24) {
                const int h = q >> 3, cc = q & 7;
                const uint32_t tcol = (uint32_t)(64 * h + 8 * cc);
                float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (cc < 7) {
                    uint32_t v[8];
                    umma::tmem_ld8(tl + kColD1 + tcol, v);
                    umma::tmem_wait_ld8(v);
                    const int hcol = 56 * h + 8 * cc, hf = hcol >= 88 ? 1 : 0, b = hcol - 88 * hf;
                    const int wi = 3 * hf + (b >> 5);      // selects instead of a dynamically indexed (local-memory) array
                    const uint32_t word = wi == 0 ? cur.hm[0] : wi == 1 ? cur.hm[1] : wi == 2 ? cur.hm[2]
                                        : wi == 3 ? cur.hm[3] : wi == 4 ? cur.hm[4] : cur.hm[5];
                    const uint32_t bits = (word >> (b & 31)) & 0xffu;
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = (bits >> j) & 1u ? __uint_as_float(v[j]) : 0.f;
                    if (valid) {
                        float *dst = d_pre + (size_t)g * kHidP + hcol;
                        *reinterpret_cast<float4 *>(dst) = make_float4(f[0], f[1], f[2], f[3]);
                        *(reinterpret_cast<float4 *>(dst) + 1) = make_float4(f[4], f[5], f[6], f[7]);
                    }
                }
                st_split8(tl, kColD1 + tcol, kColPLo + tcol, f);
            }
        }
        if (fifth == 4 && valid) {   // hidden columns 168..175 of the HBM row are padding
            float *dst = d_pre + (size_t)g * kHidP + 168;
            *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            *(reinterpret_cast<float4 *>(dst) + 1) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        umma::tmem_wait_st();
        umma::fence_before_thread_sync();
        __syncthreads();

        // ================= M2: dX = dPre W1 =================
        if (uwarp == 0 && umma::elect_one_sync()) {
            umma::fence_after_thread_sync();
            umma::gemm_3xtf32(tbase + kColDX, tbase + kColD1, tbase + kColPLo, S.w + kOffW1THi, S.w + kOffW1TLo, kInP,
                              kHidT, true);
            umma::umma_commit(&S.bar[1]);
        }
        if (!umma::mbar_wait(&S.bar[1], parity)) S.timeout = 1;
        umma::fence_after_thread_sync();

        // ================= E3: d_feat, d_anchor, d_scaling =================
        // dX chunk c (8 columns): fifth f takes chunk f and, for f < 2, chunk f + 5; chunk 6 = feat 48, 49 | view 3 | dist
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = fifth + 5 * cc;
            if (c < 7) {
                uint32_t v[8];
                umma::tmem_ld8(tl + kColDX + 8 * c, v);
                umma::tmem_wait_ld8(v);
                if (valid) {
                    if (c < 6) {
                        float2 *df = reinterpret_cast<float2 *>(d_feat + (size_t)cur.a * kFeat + 8 * c);
#pragma unroll
                        for (int j = 0; j < 4; ++j) df[j] = make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                    } else {
                        float2 *df = reinterpret_cast<float2 *>(d_feat + (size_t)cur.a * kFeat + 48);
                        df[0] = make_float2(__uint_as_float(v[0]), __uint_as_float(v[1]));
                        // view = u / |u|, dist = |u|, u = anchor - cam  (gaussian_renderer/__init__.py:106-108)
                        const float ax = __ldg(anchor + 3 * (size_t)cur.a), ay = __ldg(anchor + 3 * (size_t)cur.a + 1),
                                    az = __ldg(anchor + 3 * (size_t)cur.a + 2);
                        const float ux = ax - cx, uy = ay - cy, uz = az - cz;
                        const float d = sqrtf(ux * ux + uy * uy + uz * uz);
                        const float vx = ux / d, vy = uy / d, vz = uz / d;
                        const float gvx = __uint_as_float(v[2]), gvy = __uint_as_float(v[3]), gvz = __uint_as_float(v[4]),
                                    gd = __uint_as_float(v[5]);
                        const float dot = vx * gvx + vy * gvy + vz * gvz;
                        // the per-offset parts were accumulated by kernel 0; this is the only writer of the view path
                        atomicAdd(d_anchor + 3 * (size_t)cur.a + 0, (gvx - vx * dot) / d + gd * vx);
                        atomicAdd(d_anchor + 3 * (size_t)cur.a + 1, (gvy - vy * dot) / d + gd * vy);
                        atomicAdd(d_anchor + 3 * (size_t)cur.a + 2, (gvz - vz * dot) / d + gd * vz);
                    }
                }
            }
        }
        umma::fence_before_thread_sync();
        __syncthreads();   // dX consumed: the next tile may overwrite columns [0,288)
        umma::fence_after_thread_sync();
    }

    umma::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0 && S.timeout) atomicExch(err, 1);
    if (warp == 0) umma::tmem_dealloc(tbase, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------------------
// kernel 2: weight gradients.  Per slab of kSlab visible rows the CTA stages, split into TF32 hi / lo, the features
//     X    56  layer-1 input  [feat 50 | view 3 | dist | 1 | 0]      (the 1 makes row 54 of dW1^T the bias gradient)
//     dOut 144 (kernel 1)
//     H    176 (forward), with a 1 in column 56h + 55 of every head (-> db2 in column 55 of dW2_h^T)
//     dPre 176 (kernel 1)
// as ONE K-major operand block [row / 4][feature (kLd = 553)][row % 4] and issues, per 8 rows,
//     D_h[out, hid] += dOut_h^T H_h  (M = 128 from feature {56, 72, 120}, N = 64 from feature 200 + 56h)
//     D_x[in,  hid] += X^T dPre      (M = 128 from feature 0,             N = 176 from feature 376).
// An M = 128 operand that is narrower than 128 features simply runs on into the next features (finite values, their
// accumulator lanes are never read).  TMEM: D_o | D_c | D_v at columns 0 / 64 / 128, D_x at [192,368).
//
// Pipeline (the first version loaded through registers and sat on the DRAM latency with 8 % of the issue slots busy):
//   * the fp32 rows of dOut / H / dPre travel HBM -> shared memory as TMA bulk copies (cp.async.bulk, one per row into
//     a padded row stride that keeps the later reads conflict free), EIGHT 8-row slabs ahead (two 16-row slabs ahead
//     left the converters waiting for the copies 54 % of the time), completion on one mbarrier per stage;
//   * 16 converter warps split the staged rows into TF32 hi / lo and scatter them into the operand block (X, 10 % of the
//     bytes, is gathered through the visible-anchor list straight from HBM, one slab ahead);
//   * one extra warp issues the bulk copies and the tcgen05.mma (named barriers: converters only ARRIVE, never wait for it).
namespace ngwu {
// 8 rows = one K step per slab; 8 slabs of raw rows in flight; 2 operand buffers (a third one, a second MMA-issuing warp and
// bulk copies issued from two warps were each measured at +-0.5 %: scripts/wgrad_probe.py)
constexpr int kConv = 512, kThreads = kConv + 32, kSlab = 8, kStages = 8, kBufs = 2;
constexpr int kGX = 14, kGO = 36, kGH = 44, kGP = 44, kGroups = kGX + kGO + kGH + kGP;   // float4 groups per row: 138
constexpr int kOffX = 0, kOffO = kGX, kOffH = kOffO + kGO, kOffP = kOffH + kGH;
constexpr int kLd = 4 * kGroups + 1;      // 553 features per 4-row chunk: = 1 (mod 8) spreads the chunks over the banks
constexpr uint32_t kColDo = 0, kColDc = 64, kColDv = 128, kColDx = 192, kTmemCols = 512;
// staged raw rows: padded strides (bytes / 4 = 20 mod 32: sixteen rows at the same column hit distinct banks)
constexpr int kRawO = 148, kRawH = 180, kRawP = 180;             // floats per staged row (144 + 4, 176 + 4)
constexpr uint32_t kBytesO = 144 * 4, kBytesH = 176 * 4;
// forward-layout weight block of the SIMT kernels (contextgs_b200/neural_gaussians.py pack_decoder_weights): the
// gradient is returned in this layout
constexpr int kIn = 54, kLd1 = 152, kLdO = 12, kLdC = 32, kLdV = 72;
constexpr int kOffW1 = 0, kOffB1 = kOffW1 + kIn * kLd1, kOffW2o = kOffB1 + kLd1, kOffB2o = kOffW2o + 50 * kLdO;
constexpr int kOffW2c = kOffB2o + kLdO, kOffB2c = kOffW2c + 50 * kLdC, kOffW2v = kOffB2c + kLdC;
constexpr int kOffB2v = kOffW2v + 50 * kLdV;

struct Buf {
    float hi[(kSlab / 4) * kLd * 4];     // [row / 4][feature][row % 4]
    float lo[(kSlab / 4) * kLd * 4];
};
struct Raw {
    float o[kSlab * kRawO];
    float h[kSlab * kRawH];
    float p[kSlab * kRawP];
};
struct Smem {
    Buf buf[kBufs];
    Raw raw[kStages];
    uint32_t tmem;
    int timeout;
    alignas(8) uint64_t mma_done[kBufs];   // the MMAs that read buf[b] have completed
    alignas(8) uint64_t full[kStages];     // the bulk copies into raw[stage] have landed
};

__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     umma::smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(umma::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
}  // namespace ngwu

static int g_wgrad_dbg = 0;

__global__ void __launch_bounds__(ngwu::kThreads, 1)
neural_gaussians_wgrad_umma_kernel(const int *__restrict__ vis_idx, int Nv, const float *__restrict__ anchor,
                                   const float *__restrict__ feat, float cx, float cy, float cz,
                                   const float *__restrict__ save_h, const float *__restrict__ d_out,
                                   const float *__restrict__ d_pre, float *__restrict__ d_w, int32_t *__restrict__ err, int dbg)
{
    using namespace ngwu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int num_slabs = (Nv + kSlab - 1) / kSlab;
    const int stride = (int)gridDim.x;
    const bool issuer = umma::uniform_warp() >= kConv / 32;     // warp-uniform role

    if (warp == 0) umma::tmem_alloc(&S.tmem, kTmemCols);
    if (tid == 0) {
        for (int i = 0; i < kBufs; ++i) umma::mbar_init(&S.mma_done[i], 1);
        for (int i = 0; i < kStages; ++i) umma::mbar_init(&S.full[i], 1);
        umma::fence_mbar_init();
        S.timeout = 0;
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = S.tmem;
    // number of slabs this CTA processes
    const int n_it = blockIdx.x < num_slabs ? (num_slabs - 1 - (int)blockIdx.x) / stride + 1 : 0;

    if (issuer) {
        // =============================== issue warp: TMA bulk copies + tcgen05.mma ===============================
        const uint32_t idesc64 = umma::idesc_tf32(128, 64), idesc176 = umma::idesc_tf32(128, 176);
        // K-major no-swizzle: LBO = stride between the two 4-row chunks of one K = 8 step, SBO = stride between 8-feature groups
        const uint32_t lbo = (uint32_t)kLd * 16u, sbo = 128u;
        auto stage_rows = [&](int it) {     // bulk copies of slab `it` into raw[it % kStages]
            const int b = it % kStages, row0 = ((int)blockIdx.x + it * stride) * kSlab;
            const int rows = min(kSlab, Nv - row0);
            // ONE elected lane issues every copy of the slab with warp-uniform operands (per-lane copies make the compiler
            // serialise the lanes through a vote / BRA.U.ANY loop around each UBLKCP)
            if (umma::elect_one_sync()) {
                mbar_expect_tx(&S.full[b], (dbg & 4) ? 0u : (uint32_t)rows * (kBytesO + 2u * kBytesH));
                for (int r = 0; r < rows && !(dbg & 4); ++r) {
                    const size_t g = (size_t)(row0 + r);
                    bulk_g2s(S.raw[b].o + r * kRawO, d_out + g * 144, kBytesO, &S.full[b]);
                    bulk_g2s(S.raw[b].h + r * kRawH, save_h + g * 176, kBytesH, &S.full[b]);
                    bulk_g2s(S.raw[b].p + r * kRawP, d_pre + g * 176, kBytesH, &S.full[b]);
                }
            }
            __syncwarp();
        };
        for (int i = 0; i < kStages && i < n_it; ++i) stage_rows(i);
        for (int it = 0; it < n_it; ++it) {
            const uint32_t b = (uint32_t)it % kBufs;
            named_sync(1 + (int)b, kThreads);     // the converters have written buf[b] and are done with raw[b]
            umma::fence_after_thread_sync();
            if (umma::elect_one_sync()) {
                const uint32_t hi = umma::smem_u32(S.buf[b].hi), lo = umma::smem_u32(S.buf[b].lo);
                auto desc = [&](uint32_t base, int group, int kstep) {   // operand starting at float4 group `group`, rows 8 kstep ..
                    return umma::smem_desc_kmajor(base + (uint32_t)kstep * 2u * lbo + (uint32_t)group * 64u, lbo, sbo);
                };
                const uint32_t acc0 = it > 0 ? 1u : 0u;
#pragma unroll
                for (int s = 0; s < kSlab / 8 && !(dbg & 1); ++s) {
                    const uint32_t acc = (s > 0) ? 1u : acc0;
                    const int og[3] = {0, 4, 16};
                    const uint32_t dcol[3] = {kColDo, kColDc, kColDv};
                    // one TF32 product of each of the four accumulators in turn (independent dependency chains overlap)
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const uint32_t abase = p == 1 ? lo : hi, bbase = p == 0 ? lo : hi;
#pragma unroll
                        for (int h = 0; h < 3; ++h)
                            umma::mma_tf32_ss(tbase + dcol[h], desc(abase, kOffO + og[h], s), desc(bbase, kOffH + 14 * h, s),
                                              idesc64, p == 0 ? acc : 1u);
                        umma::mma_tf32_ss(tbase + kColDx, desc(abase, kOffX, s), desc(bbase, kOffP, s), idesc176,
                                          p == 0 ? acc : 1u);
                    }
                }
                umma::umma_commit(&S.mma_done[b]);
            }
            __syncwarp();
            if (it + kStages < n_it) stage_rows(it + kStages);    // raw[it % kStages] is free again
        }
    } else {
        // =============================== converter warps ===============================================================
        // X of the first slab travels while the first bulk copies land
        // Items are float2 (two features of one row): with 8 rows per slab a warp covers 8 rows x 4 feature pairs, and
        // both the staged-row reads (row stride = 20 banks) and the operand stores (8 c + 4 (r / 4) + r % 4) are conflict free.
        constexpr int kPairs = 2 * kGroups, kPX = 2 * kGX, kPO = 2 * kOffO, kPH = 2 * kOffH, kPP = 2 * kOffP;
        // the visible-anchor index is fetched TWO slabs ahead and the values ONE slab ahead: neither dependent load is waited for
        const bool xthread = tid < kPX * kSlab;
        const int xr = tid & (kSlab - 1), xc = tid / kSlab;      // xc: feature pair 0..27
        auto load_idx = [&](int it, int &a) {
            a = -1;
            if (!xthread || it >= n_it) return;
            const int g = ((int)blockIdx.x + it * stride) * kSlab + xr;
            if (g >= Nv) return;
            a = vis_idx ? __ldg(vis_idx + g) : g;
        };
        auto load_x = [&](int a, float2 &v) {
            v = make_float2(0.f, 0.f);
            if (a < 0) return;
            if (xc < 25) {
                v = __ldg(reinterpret_cast<const float2 *>(feat + (size_t)a * 50) + xc);
            } else {
                const float ux = __ldg(anchor + 3 * (size_t)a) - cx, uy = __ldg(anchor + 3 * (size_t)a + 1) - cy,
                            uz = __ldg(anchor + 3 * (size_t)a + 2) - cz;
                const float d = sqrtf(ux * ux + uy * uy + uz * uz);
                if (xc == 25) v = make_float2(ux / d, uy / d);
                else if (xc == 26) v = make_float2(uz / d, d);
                else v = make_float2(1.0f, 0.f);                   // feature 54 = 1: its row of dW1^T is the bias gradient
            }
        };
        auto store_item = [&](Buf &B, int r, int c, const float2 &v) {
            uint32_t h[2], l[2];
            umma::split_tf32(v.x, h[0], l[0]);
            umma::split_tf32(v.y, h[1], l[1]);
            const int dst = ((r >> 2) * kLd + 2 * c) * 4 + (r & 3);
            B.hi[dst] = __uint_as_float(h[0]);
            B.hi[dst + 4] = __uint_as_float(h[1]);
            B.lo[dst] = __uint_as_float(l[0]);
            B.lo[dst + 4] = __uint_as_float(l[1]);
        };
        float2 xv;
        int a1, a2;
        load_idx(0, a1);
        load_x(a1, xv);
        load_idx(1, a1);
        load_idx(2, a2);
        for (int it = 0; it < n_it; ++it) {
            const uint32_t b = (uint32_t)it % kBufs;
            const int st = it % kStages;
            Buf &B = S.buf[b];
            const Raw &R = S.raw[st];
            const int row0 = ((int)blockIdx.x + it * stride) * kSlab;
            // the MMAs that read buf[b] kBufs slabs ago must have completed
            if (it >= kBufs) {
                if (!umma::mbar_wait(&S.mma_done[b], (uint32_t)(it / kBufs - 1) & 1u)) S.timeout = 1;
                umma::fence_after_thread_sync();
            }
            // X (prefetched one slab ahead), then the next slab's X starts travelling
            if (xthread && !(dbg & 2)) store_item(B, xr, xc, xv);
            load_x(a1, xv);
            a1 = a2;
            load_idx(it + 3, a2);
            if (tid < (kSlab / 4) * 4) {   // the skew feature (index 552) of every chunk: keep it finite
                const int dst = ((tid >> 2) * kLd + 4 * kGroups) * 4 + (tid & 3);
                B.hi[dst] = 0.f;
                B.lo[dst] = 0.f;
            }
            // the staged rows of dOut / H / dPre
            if (!umma::mbar_wait(&S.full[st], (uint32_t)(it / kStages) & 1u)) S.timeout = 1;
            // One item = one feature of FOUR consecutive rows: four scalar reads of the staged rows (row stride = 20 banks,
            // consecutive lanes on consecutive features: conflict free) and ONE 16-byte store each for the hi and the lo
            // operand -- the four rows of a feature are adjacent in the K-major layout.  (A flat (row, feature-pair) item
            // needed four scalar stores and twice the index arithmetic per value; the converters' instruction stream is
            // what bounds this kernel, scripts/wgrad_probe.py.)
            constexpr int kFeat = 4 * (kGroups - kGX);     // 496 staged features: dOut 144 | H 176 | dPre 176
#pragma unroll
            for (int k = 0; k < (kFeat * (kSlab / 4) + kConv - 1) / kConv; ++k) {
                const int i = tid + k * kConv;
                if (i >= kFeat * (kSlab / 4)) break;
                const int q = i >= kFeat ? 1 : 0, f = i - q * kFeat;     // (kSlab / 4 == 2 row quads)
                const float *src;
                int ldr;
                float one = 0.f;                                          // columns 56h + 55 of H: bias rows of the heads
                if (f < 4 * kGO) { src = R.o + f; ldr = kRawO; }
                else if (f < 4 * (kGO + kGH)) {
                    const int fh = f - 4 * kGO;
                    src = R.h + fh; ldr = kRawH;
                    one = (fh == 55 || fh == 111 || fh == 167) ? 1.0f : 0.f;
                } else { src = R.p + (f - 4 * (kGO + kGH)); ldr = kRawP; }
                float v[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int row = 4 * q + r;
                    v[r] = row0 + row < Nv ? (one != 0.f ? 1.0f : src[row * ldr]) : 0.f;
                }
                if (!(dbg & 2)) {
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) umma::split_tf32(v[r], h[r], l[r]);
                    const int dst = (q * kLd + 4 * kGX + f) * 4;
                    *reinterpret_cast<uint4 *>(B.hi + dst) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4 *>(B.lo + dst) = make_uint4(l[0], l[1], l[2], l[3]);
                }
            }
            umma::fence_proxy_async_smem();
            umma::fence_before_thread_sync();
            named_arrive(1 + (int)b, kThreads);
        }
    }
    // ---- drain: the last commit of each buffer covers every earlier MMA (commits complete in order) ----------------
    for (int k = 0; k < kBufs && k < n_it; ++k) {
        const uint32_t idx = (uint32_t)(n_it - 1 - k);
        if (!umma::mbar_wait(&S.mma_done[idx % kBufs], (idx / kBufs) & 1u)) S.timeout = 1;
    }
    umma::fence_after_thread_sync();

    // ---- flush: lane = output unit (dW2 heads) / input unit (dW1), column = hidden unit ---------------------------------
    if (n_it >= 1 && !issuer) {
        const uint32_t tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
        const int L = 32 * (warp & 3) + lane;     // accumulator lane of this thread
        const int cq = warp >> 2;                 // four warps share a lane quadrant: split the columns
#pragma unroll 1
        for (int h = 0; h < 3; ++h) {
            int n = -1, ld = 0, offW = 0, offB = 0;
            if (h == 0) { if (L < 16 && (L & 7) < 5) n = 5 * (L >> 3) + (L & 7); ld = kLdO; offW = kOffW2o; offB = kOffB2o; }
            if (h == 1) { if (L < 40 && (L & 3) < 3) n = 3 * (L >> 2) + (L & 3); ld = kLdC; offW = kOffW2c; offB = kOffB2c; }
            if (h == 2) { if (L < 80 && (L & 7) < 7) n = 7 * (L >> 3) + (L & 7); ld = kLdV; offW = kOffW2v; offB = kOffB2v; }
            const uint32_t dcol = h == 0 ? kColDo : (h == 1 ? kColDc : kColDv);
#pragma unroll 1
            for (int c = cq; c < 7; c += 4) {
                uint32_t v[8];
                umma::tmem_ld8(tl + dcol + 8 * c, v);
                umma::tmem_wait_ld8(v);
                if (n >= 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int hh = 8 * c + j;
                        if (hh < 50) atomicAdd(d_w + offW + hh * ld + n, __uint_as_float(v[j]));
                        else if (hh == 55) atomicAdd(d_w + offB + n, __uint_as_float(v[j]));
                    }
                }
            }
        }
        // dW1 (rows 0..53) and b1 (row 54, the constant-one input)
#pragma unroll 1
        for (int c = cq; c < 22; c += 4) {
            uint32_t v[8];
            umma::tmem_ld8(tl + kColDx + 8 * c, v);
            umma::tmem_wait_ld8(v);
            if (L <= kIn) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = 8 * c + j, h = col / 56, u = col - 56 * h;
                    if (h < 3 && u < 50) {
                        if (L < kIn) atomicAdd(d_w + kOffW1 + L * kLd1 + 50 * h + u, __uint_as_float(v[j]));
                        else atomicAdd(d_w + kOffB1 + 50 * h + u, __uint_as_float(v[j]));
                    }
                }
            }
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0 && S.timeout) atomicExch(err, 1);
    if (warp == 0) umma::tmem_dealloc(tbase, kTmemCols);
}

}  // namespace cgs

using namespace cgs;

/* diagnostic switches (timing experiments only: results are wrong when set): key 1 = weight-gradient kernel of G1,
 * bit 0 skips the tcgen05.mma instructions, bit 1 the converters' hi / lo stores, bit 2 the bulk copies
 * (scripts/wgrad_probe.py: with all three skipped the kernel still takes 0.66 of its 0.95 ms -- it is bound by the
 * converters' instruction stream, index arithmetic + LDS + split, not by MMA issue, TMA issue or operand buffers) */
extern "C" int cgs_debug_set(int key, int value)
{
    if (key == 1) { g_wgrad_dbg = value; return 0; }
    return -1;
}

extern "C" int cgs_neural_gaussians_bwd_umma_packed_floats(void) { return ngbu::kPacked; }

/* scratch rows of the forward -> backward hand-over, in floats / words per visible anchor */
extern "C" int cgs_neural_gaussians_save_floats(int what)
{
    switch (what) {
    case 0: return 176;   /* save_h      */
    case 1: return 6;     /* save_hmask  */
    case 2: return 144;   /* save_pre2   */
    case 3: return 2;     /* save_rowpos */
    case 4: return ngbu::kRows;   /* rows per tile (save_tilebase has one word per tile) */
    }
    return -1;
}

extern "C" int cgs_neural_gaussians_backward_umma(const float *packed_bwd, const int32_t *vis_idx, int Nv,
                                                  const float *anchor, const float *feat, const float *offsets,
                                                  const float *scaling, const float *mask, const float *campos_host,
                                                  const uint8_t *keep_mask, const float *save_h, const uint32_t *save_hmask,
                                                  const float *save_pre2, const uint32_t *save_rowpos,
                                                  const uint32_t *save_tilebase, const float *g_xyz, const float *g_color,
                                                  const float *g_opacity, const float *g_scaling, const float *g_rot,
                                                  float *d_anchor, float *d_feat, float *d_offsets, float *d_scaling,
                                                  float *d_mask, float *d_packed_fwd, float *scratch_dout,
                                                  float *scratch_dpre, int32_t *err, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (Nv <= 0) return 0;
    CGS_CHECK_PTR(packed_bwd); CGS_CHECK_PTR(anchor); CGS_CHECK_PTR(feat); CGS_CHECK_PTR(offsets); CGS_CHECK_PTR(scaling);
    CGS_CHECK_PTR(mask); CGS_CHECK_PTR(campos_host); CGS_CHECK_PTR(keep_mask); CGS_CHECK_PTR(save_h);
    CGS_CHECK_PTR(save_hmask); CGS_CHECK_PTR(save_pre2); CGS_CHECK_PTR(save_rowpos); CGS_CHECK_PTR(save_tilebase);
    CGS_CHECK_PTR(g_xyz); CGS_CHECK_PTR(g_color); CGS_CHECK_PTR(g_opacity); CGS_CHECK_PTR(g_scaling); CGS_CHECK_PTR(g_rot);
    CGS_CHECK_PTR(d_anchor); CGS_CHECK_PTR(d_feat); CGS_CHECK_PTR(d_offsets); CGS_CHECK_PTR(d_scaling); CGS_CHECK_PTR(d_mask);
    CGS_CHECK_PTR(d_packed_fwd); CGS_CHECK_PTR(scratch_dout); CGS_CHECK_PTR(scratch_dpre); CGS_CHECK_PTR(err);
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(neural_gaussians_dgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(ngbu::Smem));
        cudaFuncSetAttribute(neural_gaussians_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(ngwu::Smem));
        if (sm_count <= 0) sm_count = kNumSMs;
    }
    StageScope sc(ST_G1_BWD, st, 3);
    {
        const size_t threads = (size_t)Nv * 16;
        neural_gaussians_postproc_backward_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
            vis_idx, Nv, offsets, scaling, mask, keep_mask, save_pre2, save_rowpos, save_tilebase, g_xyz, g_color, g_opacity,
            g_scaling, g_rot, d_anchor, d_offsets, d_scaling, d_mask, scratch_dout);
    }
    {
        const int tiles = (Nv + ngbu::kRows - 1) / ngbu::kRows;
        const int grid = tiles < sm_count ? tiles : sm_count;
        neural_gaussians_dgrad_umma_kernel<<<grid, ngbu::kThreads, sizeof(ngbu::Smem), st>>>(
            packed_bwd, vis_idx, Nv, anchor, campos_host[0], campos_host[1], campos_host[2], save_hmask, scratch_dout,
            d_anchor, d_feat, scratch_dpre, err);
    }
    {
        const int slabs = (Nv + ngwu::kSlab - 1) / ngwu::kSlab;
        const int grid = slabs < sm_count ? slabs : sm_count;
        neural_gaussians_wgrad_umma_kernel<<<grid, ngwu::kThreads, sizeof(ngwu::Smem), st>>>(
            vis_idx, Nv, anchor, feat, campos_host[0], campos_host[1], campos_host[2], save_h, scratch_dout, scratch_dpre,
            d_packed_fwd, err, g_wgrad_dbg);
    }
    return check_launch(__func__);
}

// Backward of one level of the anchor-level context model on the tcgen05 tensor cores:
// what autograd does in the reference for scene/gaussian_model.py:1596-1652 + 1666-1670 with training=True,
// predict_bpp=True.  Same contract as context_level_backward_kernel (context_model_bwd.cu), but nothing is recomputed
// and every GEMM runs as 3xTF32 tcgen05.mma: the training-mode forward (context_model_umma.cu,
// cgs_context_level_umma_forward_train) leaves (mean, scale, Q) of every coded value, the hidden activations and their
// sign bits behind (1.17 kB per level row).  All rows of the level take the same path (the ~85 % of rows that are not
// chosen for the bit-rate term simply have an output gradient that is zero outside the three step columns), so there
// are no row lists and every array is walked contiguously.
//
//   kernel 0  context_level_bwd_elem_kernel      one WARP per level row: bit-rate gradient (Low_bound, clamps, mask),
//             x_q = x + n Q, step gradient dQ -> dOut[n,176] in the MLP's output layout; G_* updated in place
//   kernel 1  context_level_dgrad_umma_kernel    dH = dOut W2 (K = 176), dPre = dH * (H > 0), dX = dPre W1, scatter-ADD of
//             dX onto the quantised attributes of the context sources (the level chain), d_hyper, d_anchor
//   kernel 2  context_level_wgrad_umma_kernel    dW2^T = H^T dOut, dW1^T = X^T dPre: SS-form MMAs with the row index as K,
//             operands staged by TMA bulk copies (8 slabs of 8 rows in flight) and converted by 16 warps; one warp
//             issues the copies, one the MMAs; accumulators resident in TMEM, one flush per CTA
#include "entropy_math.cuh"
#include "umma.cuh"

namespace cgs {
namespace cbu {
constexpr int kOut = 176, kHid = 112, kHidK = 104;
__device__ __forceinline__ float q0_of(int g) { return g == 0 ? kQf0 : (g == 1 ? kQs0 : kQo0); }

// derivatives of bits = -log2(max(|Phi_hi - Phi_lo|, 1e-6))   (Low_bound: zero below the bound); as in context_model_bwd.cu
__device__ __forceinline__ void bits_grad(float x0, float mu, float s0, float q, float x_mean, float &bits, float &gx,
                                          float &gm, float &gs, float &gq)
{
    const float lo_b = x_mean - kClampSteps * q, hi_b = x_mean + kClampSteps * q;
    const float xc = fminf(fmaxf(x0, lo_b), hi_b);
    const bool x_pass = x0 >= lo_b && x0 <= hi_b;
    const float s = fmaxf(s0, 1e-9f);
    const bool s_pass = s0 >= 1e-9f;
    const float inv = __frcp_rn(s);
    const float dh = xc + 0.5f * q - mu, dl = xc - 0.5f * q - mu;
    const float zh = __fdiv_rn(dh * inv, 1.41421356237309515f), zl = __fdiv_rn(dl * inv, 1.41421356237309515f);
    const float diff = 0.5f * (1.0f + erff(zh)) - 0.5f * (1.0f + erff(zl));
    const float lk = fabsf(diff);
    bits = -log2f(fmaxf(lk, 1e-6f));
    gx = gm = gs = gq = 0.f;
    if (lk >= 1e-6f) {
        const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const float c = 0.3989422804014327f * inv;
        const float ph = c * expf(-zh * zh), pl = c * expf(-zl * zl);
        const float gl = -sg / (lk * 0.69314718055994531f);
        gx = x_pass ? gl * (ph - pl) : 0.f;
        gm = -gl * (ph - pl);
        gq = gl * 0.5f * (ph + pl);
        gs = s_pass ? gl * (-(ph * dh - pl * dl) * inv) : 0.f;
    }
}

struct ElemArgs {
    const int *orig_idx;
    int row0, n_rows;                              // level rows [row0, row0 + n_rows)
    const float *params;                           // [n_rows,176] forward: mean[86] | scale[86] | Q[3] | 0
    const float *feat_q, *scaling_q, *offsets_q;   // forward outputs [N,*]
    const float *mask;                             // [N,10]
    const uint8_t *choose;                         // [N]
    const float *noise;                            // [n_rows,86]
    float feat_mean, scaling_mean, offset_mean;
    const float *g_bits_dev;                       // device scalar: dL / d bit_per_param
    float bits_factor;                             // rate / (n_chosen * 86)
    float *G_feat, *G_scaling, *G_offsets;         // [N,*] in: grad of the quantised values; out: grad of x
    float *d_mask;                                 // [N,10] +=
    float *d_out;                                  // [n_rows,176] gradient of the context MLP's output
};

// LITE: rows that are NOT chosen for the bit-rate term (the training forward orders every level so that the chosen rows come
// first): only the three step columns of dOut are non-zero, so the 8-column chunk [168,176) is all that is written.
template <bool LITE>
__global__ void __launch_bounds__(256) context_level_bwd_elem_kernel(ElemArgs A)
{
    const int lane = threadIdx.x & 31;
    const int row = A.row0 + (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= A.row0 + A.n_rows) return;
    const int o = __ldg(A.orig_idx + row);
    const float wbits = (!LITE && A.g_bits_dev) ? __ldg(A.g_bits_dev) * A.bits_factor : 0.f;
    const bool chosen = !LITE && wbits != 0.f && (A.choose ? A.choose[o] != 0 : true);
    const float *prow = A.params + (size_t)row * kOut;
    float *drow = A.d_out + (size_t)row * kOut;
    const float Qg[3] = {__ldg(prow + 172), __ldg(prow + 173), __ldg(prow + 174)};
    float dq[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int j = lane + 32 * t;
        if (j >= kCE) break;
        int grp, k, mcol, scol;
        float *G;
        const float *XQ;
        float x_mean;
        if (j < kCF) {
            grp = 0; k = j; mcol = j; scol = kCF + j; x_mean = A.feat_mean;
            G = A.G_feat + (size_t)o * kCF + k; XQ = A.feat_q + (size_t)o * kCF + k;
        } else if (j < kCF + kCS) {
            grp = 1; k = j - kCF; mcol = 2 * kCF + k; scol = 2 * kCF + kCS + k; x_mean = A.scaling_mean;
            G = A.G_scaling + (size_t)o * kCS + k; XQ = A.scaling_q + (size_t)o * kCS + k;
        } else {
            grp = 2; k = j - kCF - kCS; mcol = 2 * kCF + 2 * kCS + k; scol = 2 * kCF + 2 * kCS + kCO + k;
            x_mean = A.offset_mean;
            G = A.G_offsets + (size_t)o * kCO + k; XQ = A.offsets_q + (size_t)o * kCO + k;
        }
        float gx_total = *G;
        float d_mean = 0.f, d_scale = 0.f, dqj = 0.f;
        if (chosen) {
            float keep = 1.f;
            if (grp == 2) keep = __ldg(A.mask + (size_t)o * 10 + k / 3);
            float bits, gx, gm, gs, gq;
            bits_grad(__ldg(XQ), __ldg(prow + j), __ldg(prow + kCE + j), Qg[grp], x_mean, bits, gx, gm, gs, gq);
            const float wk = wbits * keep;
            gx_total += wk * gx;
            d_mean = wk * gm;
            d_scale = wk * gs;
            dqj = wk * gq;
            if (grp == 2) atomicAdd(A.d_mask + (size_t)o * 10 + k / 3, wbits * bits);
            *G = gx_total;      // x_q = x + n Q  ->  d x = d x_q
        }
        dqj += __ldg(A.noise + (size_t)row * kCE + j) * gx_total;
        dq[grp] += dqj;
        if (!LITE) {
            drow[mcol] = d_mean;
            drow[scol] = d_scale;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        dq[0] += __shfl_xor_sync(0xffffffffu, dq[0], off);
        dq[1] += __shfl_xor_sync(0xffffffffu, dq[1], off);
        dq[2] += __shfl_xor_sync(0xffffffffu, dq[2], off);
    }
    if (lane < 3) {
        // Q = max(Q0 (1 + tanh a), 1e-9): dQ/da = Q0 (1 - tanh^2) = Q0 u (2 - u), u = Q / Q0 = 1 + tanh a
        const float q = Qg[lane], q0 = q0_of(lane), u = q / q0;
        const float d = lane == 0 ? dq[0] : (lane == 1 ? dq[1] : dq[2]);
        drow[2 * kCE + lane] = q > 1e-9f ? d * q0 * u * (2.0f - u) : 0.f;
    } else if (lane == 3) {
        drow[175] = 0.f;
    } else if (LITE && lane < 8) {
        drow[164 + lane] = 0.f;      // columns 168..171 of the chunk the LITE data-gradient kernel stages
    }
}

// ------------------------------------------------------------------------------------------------------------------
template <int K1>
struct DLayout {
    static constexpr int kN1 = K1 == 71 ? 80 : 16;                      // dX width (inputs padded to a multiple of 16)
    static constexpr int kW2T = (kOut / 4) * kHid * 4;                  // floats per hi / lo part: [K = 176][N = 112]
    static constexpr int kW1T = (kHidK / 4) * kN1 * 4;                  //                           [K = 104][N = kN1]
    static constexpr int kOffW2THi = 0, kOffW2TLo = kW2T, kOffW1THi = 2 * kW2T, kOffW1TLo = 2 * kW2T + kW1T;
    static constexpr int kPacked = 2 * kW2T + 2 * kW1T;
};
constexpr int kRows = 128, kDThreads = 640;
constexpr uint32_t kColAHi = 0, kColALo = 176, kColD1 = 352, kColPLo = 0, kColDX = 112, kTmemCols = 512;

template <int K1>
struct DSmem {
    float w[DLayout<K1>::kPacked];
    uint32_t tmem;
    int timeout;
    alignas(8) uint64_t bar[2];
};

struct DArgs {
    const float *packed_w;
    const int *orig_idx, *ctx_src;
    int row0, n_rows;                // level rows [row0, row0 + n_rows)
    const uint32_t *save_hmask;      // [n_rows,4]
    const float *d_out;              // [n_rows,176] (kernel 0)
    float *d_pre;                    // [n_rows,112] -> kernel 2
    float *G_feat, *G_scaling;       // [N,*] += onto the context sources
    float *d_hyper_q, *d_anchor;     // [N,12] (=), [N,3] (+=)
    int32_t *err;
};

__device__ __forceinline__ void st_split8(uint32_t tl, uint32_t col_hi, uint32_t col_lo, const float (&v)[8])
{
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) umma::split_tf32(v[j], hi[j], lo[j]);
    umma::tmem_st8(tl + col_hi, hi);
    umma::tmem_st8(tl + col_lo, lo);
}

// LITE (rows not chosen for the bit-rate term): dOut is zero outside the chunk [168,176), so one chunk is staged and the
// first GEMM shrinks from K = 176 (66 tcgen05.mma) to K = 8 (3).
template <int K1, bool LITE>
__global__ void __launch_bounds__(kDThreads, 1) context_level_dgrad_umma_kernel(DArgs A)
{
    using LY = DLayout<K1>;
    using SM = DSmem<K1>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SM &S = *reinterpret_cast<SM *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fifth = warp >> 2;                       // five threads per row
    const int row = 32 * (warp & 3) + lane;
    const int uwarp = umma::uniform_warp();
    const int num_tiles = (A.n_rows + kRows - 1) / kRows;

    if (warp == 0) umma::tmem_alloc(&S.tmem, kTmemCols);
    if (tid == 0) {
        umma::mbar_init(&S.bar[0], 1);
        umma::mbar_init(&S.bar[1], 1);
        umma::fence_mbar_init();
        S.timeout = 0;
    }
    {
        const float4 *s4 = reinterpret_cast<const float4 *>(A.packed_w);
        float4 *d4 = reinterpret_cast<float4 *>(S.w);
        for (int i = tid; i < LY::kPacked / 4; i += kDThreads) d4[i] = __ldg(s4 + i);
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = S.tmem;
    const uint32_t tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16);

    // everything a tile reads from HBM is fetched one tile AHEAD: this thread's chunks (fifth, fifth + 5, ...) of the
    // 22 eight-column chunks of its dOut row, the sign bits of the hidden layer and the two scatter targets
    struct Pre {
        float4 v[5][2];
        uint32_t hm[4];
        int o, s;
    };
    struct Cur {          // what the later phases of a tile still need after its operand has been staged
        uint32_t hm[4];
        int o, s;
    };
    auto prefetch = [&](int tile, Pre &p) {
        const int g = A.row0 + tile * kRows + row;
        p.o = -1; p.s = -1;
#pragma unroll
        for (int i = 0; i < 4; ++i) p.hm[i] = 0u;
#pragma unroll
        for (int i = 0; i < 5; ++i) p.v[i][0] = p.v[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tile >= num_tiles || g >= A.row0 + A.n_rows) return;
        p.o = __ldg(A.orig_idx + g);
        p.s = K1 == 71 ? __ldg(A.ctx_src + g) : p.o;
        const float4 *src = reinterpret_cast<const float4 *>(A.d_out + (size_t)g * kOut);
        if (LITE) {
            if (fifth == 0) {
                p.v[0][0] = __ldg(src + 42);
                p.v[0][1] = __ldg(src + 43);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int c = fifth + 5 * i;
                if (c < 22) {
                    p.v[i][0] = __ldg(src + 2 * c);
                    p.v[i][1] = __ldg(src + 2 * c + 1);
                }
            }
        }
        const uint4 m = __ldg(reinterpret_cast<const uint4 *>(A.save_hmask) + g);
        p.hm[0] = m.x; p.hm[1] = m.y; p.hm[2] = m.z; p.hm[3] = m.w;
    };
    Pre nxt;
    prefetch(blockIdx.x, nxt);

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t parity = it & 1u;
        const int g = A.row0 + tile * kRows + row;
        Cur cur;
#pragma unroll
        for (int i = 0; i < 4; ++i) cur.hm[i] = nxt.hm[i];
        cur.o = nxt.o; cur.s = nxt.s;
        const bool valid = cur.o >= 0;

        // ---- stage dOut as the A operand: hi / lo into TMEM; then the next tile's row starts travelling -----------------
        if (LITE) {
            if (fifth == 0) {
                const float f[8] = {nxt.v[0][0].x, nxt.v[0][0].y, nxt.v[0][0].z, nxt.v[0][0].w,
                                    nxt.v[0][1].x, nxt.v[0][1].y, nxt.v[0][1].z, nxt.v[0][1].w};
                st_split8(tl, kColAHi + 168, kColALo + 168, f);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int c = fifth + 5 * i;
                if (c < 22) {
                    const float f[8] = {nxt.v[i][0].x, nxt.v[i][0].y, nxt.v[i][0].z, nxt.v[i][0].w,
                                        nxt.v[i][1].x, nxt.v[i][1].y, nxt.v[i][1].z, nxt.v[i][1].w};
                    st_split8(tl, kColAHi + 8 * c, kColALo + 8 * c, f);
                }
            }
        }
        prefetch(tile + (int)gridDim.x, nxt);
        umma::tmem_wait_st();
        umma::fence_before_thread_sync();
        __syncthreads();
        // ---- M1: dH = dOut W2 ------------------------------------------------------------------------------------------
        if (uwarp == 0 && umma::elect_one_sync()) {     // warp-uniform branch + elect: back-to-back tcgen05.mma (umma.cuh)
            umma::fence_after_thread_sync();
            if (LITE)   // K chunk [168,176): B chunks 42, 43 of W2^T
                umma::gemm_3xtf32(tbase + kColD1, tbase + kColAHi + 168, tbase + kColALo + 168, S.w + LY::kOffW2THi + 42 * kHid * 4,
                                  S.w + LY::kOffW2TLo + 42 * kHid * 4, kHid, 8, true);
            else
                umma::gemm_3xtf32(tbase + kColD1, tbase + kColAHi, tbase + kColALo, S.w + LY::kOffW2THi, S.w + LY::kOffW2TLo,
                                  kHid, kOut, true);
            umma::umma_commit(&S.bar[0]);
        }
        if (!umma::mbar_wait(&S.bar[0], parity)) S.timeout = 1;
        umma::fence_after_thread_sync();
        // ---- E2: dPre = dH * (H > 0): hi in place, lo to [0,112), fp32 copy to HBM for the weight gradients ---------------
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {
            const int q = fifth + 5 * qq;
            if (q < 14) {
                const uint32_t col = (uint32_t)(8 * q);
                uint32_t v[8];
                umma::tmem_ld8(tl + kColD1 + col, v);
                umma::tmem_wait_ld8(v);
                const int hf = q >= 7 ? 1 : 0, b = 8 * q - 56 * hf;
                const int wi = 2 * hf + (b >> 5);
                const uint32_t word = wi == 0 ? cur.hm[0] : wi == 1 ? cur.hm[1] : wi == 2 ? cur.hm[2] : cur.hm[3];
                const uint32_t bits = (word >> (b & 31)) & 0xffu;
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = (bits >> j) & 1u ? __uint_as_float(v[j]) : 0.f;
                st_split8(tl, kColD1 + col, kColPLo + col, f);
                if (valid) {
                    float *dst = A.d_pre + (size_t)g * kHid + col;
                    *reinterpret_cast<float4 *>(dst) = make_float4(f[0], f[1], f[2], f[3]);
                    *(reinterpret_cast<float4 *>(dst) + 1) = make_float4(f[4], f[5], f[6], f[7]);
                }
            }
        }
        umma::tmem_wait_st();
        umma::fence_before_thread_sync();
        __syncthreads();
        // ---- M2: dX = dPre W1 --------------------------------------------------------------------------------------------
        if (uwarp == 0 && umma::elect_one_sync()) {
            umma::fence_after_thread_sync();
            umma::gemm_3xtf32(tbase + kColDX, tbase + kColD1, tbase + kColPLo, S.w + LY::kOffW1THi, S.w + LY::kOffW1TLo,
                              LY::kN1, kHidK, true);
            umma::umma_commit(&S.bar[1]);
        }
        if (!umma::mbar_wait(&S.bar[1], parity)) S.timeout = 1;
        umma::fence_after_thread_sync();
        // ---- E3: scatter dX: the level chain (context sources of the coarser level), hyper latents, anchors ---------------
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int c = fifth + 5 * cc;
            if (c < LY::kN1 / 8) {
                uint32_t v[8];
                umma::tmem_ld8(tl + kColDX + 8 * c, v);
                umma::tmem_wait_ld8(v);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = 8 * c + j;
                        const float x = __uint_as_float(v[j]);
                        if (k < 3) {
                            atomicAdd(A.d_anchor + 3 * (size_t)cur.s + k, x);
                        } else if (K1 == 71) {
                            if (k < 3 + kCF) atomicAdd(A.G_feat + (size_t)cur.s * kCF + (k - 3), x);
                            else if (k < 3 + kCF + kCS) atomicAdd(A.G_scaling + (size_t)cur.s * kCS + (k - 3 - kCF), x);
                            else if (k < K1) A.d_hyper_q[(size_t)cur.o * kHyper + (k - 3 - kCF - kCS)] = x;
                        } else {
                            if (k < K1) A.d_hyper_q[(size_t)cur.o * kHyper + (k - 3)] = x;
                        }
                    }
                }
            }
        }
        umma::fence_before_thread_sync();
        __syncthreads();   // dX consumed: the next tile may overwrite columns [0,352)
        umma::fence_after_thread_sync();
    }

    umma::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0 && S.timeout) atomicExch(A.err, 1);
    if (warp == 0) umma::tmem_dealloc(tbase, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------------------
// kernel 2: weight gradients.  Features of one row in the K-major operand block [row / 4][feature][row % 4]:
//     X    kXF  layer-1 input as the forward stages it [anchor | feat_q | scaling_q | hyper_q] (or [level anchor | hyper_q]),
//               then a constant 1 (-> row K1 of dW1^T is the bias gradient), zero padded to a multiple of 8
//     H    112  (forward) with a 1 in column 100 (-> lane 100 of dW2^T is the bias gradient)
//     dOut 176  (kernel 0)
//     dPre 112  (kernel 1)
//   D_w2[hidden, out] += H^T dOut   (M = 128 from feature kXF, N = 176 from feature kXF + 112)     TMEM [0,176)
//   D_w1[in, hidden]  += X^T dPre   (M = 128 from feature 0,   N = 112 from feature kXF + 288)     TMEM [192,304)
template <int K1>
struct WLayout {
    static constexpr int kXF = K1 == 71 ? 72 : 16;
    static constexpr int kFH = kXF, kFO = kFH + kHid, kFP = kFO + kOut, kFeat = kFP + kHid;   // 472 | 416 (multiples of 8)
    static constexpr int kLd = kFeat + 1;                                                    // = 1 (mod 8)
};
constexpr int kConv = 512, kWThreads = kConv + 64, kSlab = 8, kStages = 8;
constexpr int kRawH = kHid + 4, kRawO = kOut + 4, kRawP = kHid + 4;      // padded staged rows (floats; = 20 mod 32)
constexpr uint32_t kColW2 = 0, kColW1 = 192;

template <int K1>
struct WSmem {
    float hi[2][(kSlab / 4) * WLayout<K1>::kLd * 4];
    float lo[2][(kSlab / 4) * WLayout<K1>::kLd * 4];
    float raw_h[kStages][kSlab * kRawH];
    float raw_o[kStages][kSlab * kRawO];
    float raw_p[kStages][kSlab * kRawP];
    uint32_t tmem;
    int timeout;
    alignas(8) uint64_t mma_done[2];
    alignas(8) uint64_t full[kStages];      // the bulk copies into stage s have landed
    alignas(8) uint64_t empty[kStages];     // the 16 converter warps are done reading stage s
};

struct WArgs {
    const int *orig_idx, *ctx_src;
    const float *level_anchor;
    int row0, n_rows;                // level rows [row0, row0 + n_rows)
    const float *anchor, *hyper_q, *feat_q, *scaling_q;
    const float *save_h, *d_out, *d_pre;
    float *d_w;                      // packed layout of pack_grid_weights_bwd: W1[in][101] | b1[100] | W2[100][177] | b2[176]
    int32_t *err;
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

// LITE (rows not chosen for the bit-rate term): only the columns [168,176) of dOut are non-zero -- 64 bytes of the row
// are staged (columns 160..175), 16 features converted and the dW2 product runs with N = 16 into columns [160,176).
template <int K1, bool LITE>
__global__ void __launch_bounds__(kWThreads, 1) context_level_wgrad_umma_kernel(WArgs A)
{
    using LY = WLayout<K1>;
    using SM = WSmem<K1>;
    constexpr int kLd1 = 101, kLd2 = 177;
    constexpr int kW1 = 0, kB1 = K1 * kLd1, kW2 = kB1 + kGH, kB2 = kW2 + kGH * kLd2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SM &S = *reinterpret_cast<SM *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int num_slabs = (A.n_rows + kSlab - 1) / kSlab;
    const int stride = (int)gridDim.x;
    const int n_it = (int)blockIdx.x < num_slabs ? (num_slabs - 1 - (int)blockIdx.x) / stride + 1 : 0;

    if (warp == 0) umma::tmem_alloc(&S.tmem, kTmemCols);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) umma::mbar_init(&S.mma_done[i], 1);
        for (int i = 0; i < kStages; ++i) {
            umma::mbar_init(&S.full[i], 1);
            umma::mbar_init(&S.empty[i], kConv / 32);
        }
        umma::fence_mbar_init();
        S.timeout = 0;
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = S.tmem;
    constexpr int kSyncCount = kConv + 32;     // converters arrive, the MMA warp waits

    const int role_warp = umma::uniform_warp();                  // warp-uniform roles
    if (role_warp == kConv / 32 + 1) {
        // =============================== TMA warp: bulk copies of the staged rows, kStages slabs ahead ===================
        auto stage_rows = [&](int it) {
            const int st = it % kStages, row0 = ((int)blockIdx.x + it * stride) * kSlab;
            const int rows = min(kSlab, A.n_rows - row0);
            constexpr uint32_t kOBytes = LITE ? 64u : (uint32_t)(kOut * 4);
            // ONE elected lane issues every copy of the slab with warp-uniform operands (per-lane copies make the compiler
            // serialise the lanes through a vote / BRA.U.ANY loop around each UBLKCP)
            if (umma::elect_one_sync()) {
                mbar_expect_tx(&S.full[st], (uint32_t)rows * ((uint32_t)(2 * kHid * 4) + kOBytes));
                for (int r = 0; r < rows; ++r) {
                    const size_t g = (size_t)(A.row0 + row0 + r);
                    bulk_g2s(&S.raw_h[st][r * kRawH], A.save_h + g * kHid, kHid * 4, &S.full[st]);
                    if (LITE) bulk_g2s(&S.raw_o[st][r * kRawO + 160], A.d_out + g * kOut + 160, 64, &S.full[st]);
                    else bulk_g2s(&S.raw_o[st][r * kRawO], A.d_out + g * kOut, kOut * 4, &S.full[st]);
                    bulk_g2s(&S.raw_p[st][r * kRawP], A.d_pre + g * kHid, kHid * 4, &S.full[st]);
                }
            }
            __syncwarp();
        };
        for (int i = 0; i < kStages && i < n_it; ++i) stage_rows(i);
        for (int it = 0; it + kStages < n_it; ++it) {
            // stage it % kStages is free once all 16 converter warps have arrived on its `empty` barrier for slab `it`
            if (!umma::mbar_wait(&S.empty[it % kStages], (uint32_t)(it / kStages) & 1u)) S.timeout = 1;
            stage_rows(it + kStages);
        }
    } else if (role_warp == kConv / 32) {
        // =============================== MMA warp ==========================================================================
        const uint32_t idescW2 = umma::idesc_tf32(128, LITE ? 16 : kOut), idescW1 = umma::idesc_tf32(128, kHid);
        constexpr int kOFeat = LITE ? 160 : 0;      // first dOut feature of the dW2 product (and its accumulator column)
        const uint32_t lbo = (uint32_t)LY::kLd * 16u, sbo = 128u;
        for (int it = 0; it < n_it; ++it) {
            const uint32_t b = (uint32_t)it & 1u;
            named_sync(1 + (int)b, kSyncCount);     // the converters have written buffer b
            umma::fence_after_thread_sync();
            if (umma::elect_one_sync()) {
                const uint32_t hi = umma::smem_u32(S.hi[b]), lo = umma::smem_u32(S.lo[b]);
                auto desc = [&](uint32_t base, int feature) {
                    return umma::smem_desc_kmajor(base + (uint32_t)feature * 16u, lbo, sbo);
                };
                const uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t ab = p == 1 ? lo : hi, bb = p == 0 ? lo : hi;
                    umma::mma_tf32_ss(tbase + kColW2 + kOFeat, desc(ab, LY::kFH), desc(bb, LY::kFO + kOFeat), idescW2,
                                      p == 0 ? acc : 1u);
                    umma::mma_tf32_ss(tbase + kColW1, desc(ab, 0), desc(bb, LY::kFP), idescW1, p == 0 ? acc : 1u);
                }
                umma::umma_commit(&S.mma_done[b]);
            }
            __syncwarp();
        }
    } else {
        // =============================== converter warps ===================================================================
        // layer-1 input of a row, feature pair c (as the forward stages it): the row's indices (orig, ctx_src) are fetched
        // TWO slabs ahead and the values ONE slab ahead, so neither of the two dependent loads is waited for
        const bool xthread = tid < (LY::kXF / 2) * kSlab;
        const int xr = tid & (kSlab - 1), xc = tid / kSlab;
        auto load_idx = [&](int it, int &o, int &s2) {
            o = -1; s2 = -1;
            if (!xthread || it >= n_it) return;
            const int gl = ((int)blockIdx.x + it * stride) * kSlab + xr;
            if (gl >= A.n_rows) return;
            o = __ldg(A.orig_idx + A.row0 + gl);
            s2 = K1 == 71 ? __ldg(A.ctx_src + A.row0 + gl) : o;
        };
        auto load_x = [&](int it, int o, int s2, float2 &v) {
            v = make_float2(0.f, 0.f);
            if (o < 0) return;
            const int g = A.row0 + ((int)blockIdx.x + it * stride) * kSlab + xr;
            float e[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int f = 2 * xc + i;
                float x = 0.f;
                if (K1 == 71) {
                    if (f < 3) x = __ldg(A.anchor + 3 * (size_t)s2 + f);
                    else if (f < 3 + kCF) x = A.feat_q[(size_t)s2 * kCF + (f - 3)];
                    else if (f < 3 + kCF + kCS) x = A.scaling_q[(size_t)s2 * kCS + (f - 3 - kCF)];
                    else if (f < K1) x = __ldg(A.hyper_q + (size_t)o * kHyper + (f - 3 - kCF - kCS));
                    else if (f == K1) x = 1.0f;
                } else {
                    if (f < 3) x = __ldg(A.level_anchor + 3 * (size_t)g + f);
                    else if (f < K1) x = __ldg(A.hyper_q + (size_t)o * kHyper + (f - 3));
                    else if (f == K1) x = 1.0f;
                }
                e[i] = x;
            }
            v = make_float2(e[0], e[1]);
        };
        float2 xv;
        int o1, s1, o2, s2;
        load_idx(0, o1, s1);
        load_x(0, o1, s1, xv);
        load_idx(1, o1, s1);
        load_idx(2, o2, s2);
        for (int it = 0; it < n_it; ++it) {
            const uint32_t b = (uint32_t)it & 1u;
            const int st = it % kStages;
            float *Bhi = S.hi[b], *Blo = S.lo[b];
            auto store_item = [&](int r, int f, const float2 &v) {     // features f, f + 1 of row r
                uint32_t h[2], l[2];
                umma::split_tf32(v.x, h[0], l[0]);
                umma::split_tf32(v.y, h[1], l[1]);
                const int dst = ((r >> 2) * LY::kLd + f) * 4 + (r & 3);
                Bhi[dst] = __uint_as_float(h[0]);
                Bhi[dst + 4] = __uint_as_float(h[1]);
                Blo[dst] = __uint_as_float(l[0]);
                Blo[dst + 4] = __uint_as_float(l[1]);
            };
            const int row0 = ((int)blockIdx.x + it * stride) * kSlab;
            if (it >= 2) {   // the MMAs that read buffer b two slabs ago must have completed
                if (!umma::mbar_wait(&S.mma_done[b], (uint32_t)((it >> 1) - 1) & 1u)) S.timeout = 1;
                umma::fence_after_thread_sync();
            }
            if (xthread) store_item(xr, 2 * xc, xv);
            load_x(it + 1, o1, s1, xv);
            o1 = o2; s1 = s2;
            load_idx(it + 3, o2, s2);
            if (tid < kSlab) {   // the skew feature of every chunk: keep it finite
                const int dst = ((tid >> 2) * LY::kLd + LY::kFeat) * 4 + (tid & 3);
                Bhi[dst] = 0.f;
                Blo[dst] = 0.f;
            }
            if (!umma::mbar_wait(&S.full[st], (uint32_t)(it / kStages) & 1u)) S.timeout = 1;
            // H (with the bias one in column 100), dOut, dPre
            for (int i = tid; i < (kHid / 2) * kSlab; i += kConv) {
                const int r = i & (kSlab - 1), c = i / kSlab;
                float2 v = make_float2(0.f, 0.f);
                if (row0 + r < A.n_rows) {
                    v = *reinterpret_cast<const float2 *>(&S.raw_h[st][r * kRawH + 2 * c]);
                    if (c == 50) v.x = 1.0f;
                }
                store_item(r, LY::kFH + 2 * c, v);
            }
            if (LITE) {
                for (int i = tid; i < 8 * kSlab; i += kConv) {          // features 160..175: zeros | the step chunk
                    const int r = i & (kSlab - 1), c = 80 + i / kSlab;
                    float2 v = make_float2(0.f, 0.f);
                    if (row0 + r < A.n_rows && c >= 84) v = *reinterpret_cast<const float2 *>(&S.raw_o[st][r * kRawO + 2 * c]);
                    store_item(r, LY::kFO + 2 * c, v);
                }
            } else {
                for (int i = tid; i < (kOut / 2) * kSlab; i += kConv) {
                    const int r = i & (kSlab - 1), c = i / kSlab;
                    float2 v = make_float2(0.f, 0.f);
                    if (row0 + r < A.n_rows) v = *reinterpret_cast<const float2 *>(&S.raw_o[st][r * kRawO + 2 * c]);
                    store_item(r, LY::kFO + 2 * c, v);
                }
            }
            for (int i = tid; i < (kHid / 2) * kSlab; i += kConv) {
                const int r = i & (kSlab - 1), c = i / kSlab;
                float2 v = make_float2(0.f, 0.f);
                if (row0 + r < A.n_rows) v = *reinterpret_cast<const float2 *>(&S.raw_p[st][r * kRawP + 2 * c]);
                store_item(r, LY::kFP + 2 * c, v);
            }
            __syncwarp();
            if (lane == 0) {   // this warp is done with the staged rows
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(&S.empty[st])) : "memory");
            }
            umma::fence_proxy_async_smem();
            umma::fence_before_thread_sync();
            named_arrive(1 + (int)b, kSyncCount);
        }
    }
    // ---- drain: the last commit of each buffer covers every earlier MMA ----------------------------------------------------
    if (n_it >= 1) {
        const uint32_t last = (uint32_t)n_it - 1;
        if (!umma::mbar_wait(&S.mma_done[last & 1u], (last >> 1) & 1u)) S.timeout = 1;
        if (n_it >= 2) {
            const uint32_t prev = (uint32_t)n_it - 2;
            if (!umma::mbar_wait(&S.mma_done[prev & 1u], (prev >> 1) & 1u)) S.timeout = 1;
        }
    }
    umma::fence_after_thread_sync();
    // ---- flush -----------------------------------------------------------------------------------------------------------------
    if (n_it >= 1 && warp < kConv / 32) {
        const uint32_t tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16);
        const int L = 32 * (warp & 3) + lane;     // accumulator lane: hidden unit (D_w2) / input unit (D_w1)
        const int cq = warp >> 2;                 // four warps share a lane quadrant: split the columns
#pragma unroll 1
        for (int c = (LITE ? 20 : 0) + cq; c < kOut / 8; c += 4) {
            uint32_t v[8];
            umma::tmem_ld8(tl + kColW2 + 8 * c, v);
            umma::tmem_wait_ld8(v);
            if (L <= kGH) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = 8 * c + j;
                    const float x = __uint_as_float(v[j]);
                    if (n < kGO && x != 0.f) atomicAdd(A.d_w + (L < kGH ? kW2 + L * kLd2 + n : kB2 + n), x);
                }
            }
        }
#pragma unroll 1
        for (int c = cq; c < kHid / 8; c += 4) {
            uint32_t v[8];
            umma::tmem_ld8(tl + kColW1 + 8 * c, v);
            umma::tmem_wait_ld8(v);
            if (L <= K1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int hh = 8 * c + j;
                    const float x = __uint_as_float(v[j]);
                    if (hh < kGH && x != 0.f) atomicAdd(A.d_w + (L < K1 ? kW1 + L * kLd1 + hh : kB1 + hh), x);
                }
            }
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0 && S.timeout) atomicExch(A.err, 1);
    if (warp == 0) umma::tmem_dealloc(tbase, kTmemCols);
}

template <int K1>
static int launch_all(ElemArgs e, DArgs d, WArgs w, int n_full, cudaStream_t st)
{
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(context_level_dgrad_umma_kernel<K1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(DSmem<K1>));
        cudaFuncSetAttribute(context_level_dgrad_umma_kernel<K1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(DSmem<K1>));
        cudaFuncSetAttribute(context_level_wgrad_umma_kernel<K1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(WSmem<K1>));
        cudaFuncSetAttribute(context_level_wgrad_umma_kernel<K1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(WSmem<K1>));
        if (sm_count <= 0) sm_count = kNumSMs;
    }
    const int n_all = e.n_rows;
    // rows [0, n_full): chosen for the bit-rate term (full output gradient); rows [n_full, n): step columns only
    for (int part = 0; part < 2; ++part) {
        const int row0 = part == 0 ? 0 : n_full, n = part == 0 ? n_full : n_all - n_full;
        if (n <= 0) continue;
        e.row0 = d.row0 = w.row0 = row0;
        e.n_rows = d.n_rows = w.n_rows = n;
        StageScope sc(ST_CTX_LEVEL_BWD, st, 3);
        const unsigned eb = (unsigned)(((size_t)n * 32 + 255) / 256);
        const int tiles = (n + kRows - 1) / kRows, slabs = (n + kSlab - 1) / kSlab;
        const int gd = tiles < sm_count ? tiles : sm_count, gw = slabs < sm_count ? slabs : sm_count;
        if (part == 0) {
            context_level_bwd_elem_kernel<false><<<eb, 256, 0, st>>>(e);
            context_level_dgrad_umma_kernel<K1, false><<<gd, kDThreads, sizeof(DSmem<K1>), st>>>(d);
            context_level_wgrad_umma_kernel<K1, false><<<gw, kWThreads, sizeof(WSmem<K1>), st>>>(w);
        } else {
            context_level_bwd_elem_kernel<true><<<eb, 256, 0, st>>>(e);
            context_level_dgrad_umma_kernel<K1, true><<<gd, kDThreads, sizeof(DSmem<K1>), st>>>(d);
            context_level_wgrad_umma_kernel<K1, true><<<gw, kWThreads, sizeof(WSmem<K1>), st>>>(w);
        }
    }
    return check_launch("cgs_context_level_backward_umma");
}
}  // namespace cbu
}  // namespace cgs

using namespace cgs;

extern "C" int cgs_context_level_bwd_umma_packed_floats(int in_dim)
{
    if (in_dim == 71) return cbu::DLayout<71>::kPacked;
    if (in_dim == 15) return cbu::DLayout<15>::kPacked;
    return -1;
}

extern "C" int cgs_context_level_backward_umma(int in_dim, const float *packed_bwd, const int32_t *orig_idx,
                                               const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                               const float *anchor, const float *hyper_q, const float *feat_q,
                                               const float *scaling_q, const float *offsets_q, const float *mask,
                                               const uint8_t *choose, const float *noise, float feat_mean,
                                               float scaling_mean, float offset_mean, const float *g_bits_dev,
                                               float bits_factor, const float *params, const float *save_h,
                                               const uint32_t *save_hmask, float *G_feat, float *G_scaling, float *G_offsets,
                                               float *d_mask, float *d_hyper_q, float *d_anchor, float *d_packed_w,
                                               float *scratch_dout, float *scratch_dpre, int32_t *err, int n_full,
                                               void *stream)
{
    if (n_rows <= 0) return 0;
    if (n_full < 0 || n_full > n_rows) {
        set_error("%s: n_full out of range", __func__);
        return -2;
    }
    CGS_CHECK_PTR(packed_bwd); CGS_CHECK_PTR(orig_idx); CGS_CHECK_PTR(anchor); CGS_CHECK_PTR(hyper_q); CGS_CHECK_PTR(feat_q);
    CGS_CHECK_PTR(scaling_q); CGS_CHECK_PTR(offsets_q); CGS_CHECK_PTR(mask); CGS_CHECK_PTR(noise); CGS_CHECK_PTR(params);
    CGS_CHECK_PTR(save_h); CGS_CHECK_PTR(save_hmask); CGS_CHECK_PTR(G_feat); CGS_CHECK_PTR(G_scaling); CGS_CHECK_PTR(G_offsets);
    CGS_CHECK_PTR(d_mask); CGS_CHECK_PTR(d_hyper_q); CGS_CHECK_PTR(d_anchor); CGS_CHECK_PTR(d_packed_w);
    CGS_CHECK_PTR(scratch_dout); CGS_CHECK_PTR(scratch_dpre); CGS_CHECK_PTR(err);
    cbu::ElemArgs e;
    e.orig_idx = orig_idx; e.row0 = 0; e.n_rows = n_rows; e.params = params; e.feat_q = feat_q; e.scaling_q = scaling_q;
    e.offsets_q = offsets_q; e.mask = mask; e.choose = choose; e.noise = noise; e.feat_mean = feat_mean;
    e.scaling_mean = scaling_mean; e.offset_mean = offset_mean; e.g_bits_dev = g_bits_dev; e.bits_factor = bits_factor;
    e.G_feat = G_feat; e.G_scaling = G_scaling; e.G_offsets = G_offsets; e.d_mask = d_mask; e.d_out = scratch_dout;
    cbu::DArgs d;
    d.packed_w = packed_bwd; d.orig_idx = orig_idx; d.ctx_src = ctx_src; d.row0 = 0; d.n_rows = n_rows; d.save_hmask = save_hmask;
    d.d_out = scratch_dout; d.d_pre = scratch_dpre; d.G_feat = G_feat; d.G_scaling = G_scaling; d.d_hyper_q = d_hyper_q;
    d.d_anchor = d_anchor; d.err = err;
    cbu::WArgs w;
    w.orig_idx = orig_idx; w.ctx_src = ctx_src; w.level_anchor = level_anchor; w.row0 = 0; w.n_rows = n_rows; w.anchor = anchor;
    w.hyper_q = hyper_q; w.feat_q = feat_q; w.scaling_q = scaling_q; w.save_h = save_h; w.d_out = scratch_dout;
    w.d_pre = scratch_dpre; w.d_w = d_packed_w; w.err = err;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_dim == 71) {
        CGS_CHECK_PTR(ctx_src);
        return cbu::launch_all<71>(e, d, w, n_full, st);
    }
    if (in_dim == 15) {
        CGS_CHECK_PTR(level_anchor);
        return cbu::launch_all<15>(e, d, w, n_full, st);
    }
    set_error("%s: unsupported context-MLP input width %d", __func__, in_dim);
    return -2;
}

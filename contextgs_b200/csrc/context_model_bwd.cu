// Backward of the anchor-level context / entropy model in TRAINING mode:
// what autograd does in the reference for scene/gaussian_model.py:1556-1707 with training=True,
// predict_bpp=True (x_q = x + U(-1/2,1/2) * Q; bit_per_param over a random 15 % of the anchors).
//
//  * context_level_backward_kernel<K1> : one fused kernel per level, launched fine -> coarse.  For the
//    rows of a level it adds the bit-rate gradient to the gradient arriving on the quantised
//    attributes, pulls it back through x_q = x + n*Q, the adaptive steps Q = Q0 (1 + tanh a), the
//    likelihood (incl. Low_bound, utils/entropy_models.py:141-156) and the context MLP, and
//    scatter-ADDS the input gradient onto the quantised attributes of the context source anchors of
//    the COARSER level (the level chain: one representative feeds ~5 finer anchors).  The arrays
//    G_feat / G_scaling / G_offsets are at once the incoming gradients and, after the last launch,
//    the gradients of the unquantised attributes.
//  * eb_backward_kernel : EntropyBottleneck (hyper prior) likelihood gradient w.r.t. the hyper
//    latents and the packed per-channel parameters.
//
// Forward activations are recomputed per tile; weight gradients live in registers across the tiles of
// a persistent CTA.  One copy of each weight matrix (odd leading dimension) serves W and W^T.
#include "entropy_math.cuh"
#include "mlp_tile.cuh"

namespace cgs {
namespace cmb {
constexpr int kCF = 50, kCS = 6, kCO = 30, kCE = 86, kCtx = 59, kHyper = 12, kGH = 100, kGO = 175;
constexpr int kLd1 = 101, kLd2 = 177;  // odd: conflict-free reads of W and of W^T
__device__ __forceinline__ float q0_of(int g) { return g == 0 ? 1.0f : (g == 1 ? 0.001f : 0.2f); }
constexpr float kClampSteps = 15000.0f;
constexpr int kEbParams = 59;

template <int K1>
struct Smem {
    static constexpr int kW1 = 0, kB1 = K1 * kLd1, kW2 = kB1 + kGH, kB2 = kW2 + kGH * kLd2, kWFloats0 = kB2 + 176;
    static constexpr int kWFloats = (kWFloats0 + 3) / 4 * 4;
    static constexpr int kXRows = (K1 + 3) / 4 * 4;
    float w[kWFloats];
    float x[kXRows * kTMp];
    float h[kGH * kTMp];
    float out[176 * kTMp];
    float zero[kXRows];
    float Q[3 * kTM];
    float dQ[3 * kTM];
    int orig[kTM];
    int src[kTM];
    int lrow[kTM];       // level row of each tile row (identity unless a row list is given)
    uint8_t chosen[kTM];
    int tile;
};

struct Args {
    const float *packed_w;
    const int *orig_idx, *ctx_src;
    const float *level_anchor;
    int n_rows;
    const float *anchor, *hyper_q;
    const float *feat_q, *scaling_q, *offsets_q;   // forward outputs [N,*]
    const float *mask;                             // [N,10]
    const uint8_t *choose;                         // [N]
    const float *noise;                            // [n_rows,86]
    float feat_mean, scaling_mean, offset_mean;
    const float *g_bits_dev;                       // device scalar: dL / d bit_per_param
    float bits_factor;                             // rate / (n_chosen * 86)
    float *G_feat, *G_scaling, *G_offsets;         // [N,*] in: grad of the quantised values; out: grad of x
    float *d_mask, *d_hyper_q, *d_anchor;          // [N,10] (+=), [N,12] (=), [N,3] (+=)
    float *d_w;                                    // packed layout, +=
    uint32_t *ticket;
    const int *row_list;                           // optional: the level rows this launch processes (n_rows = its length)
};

// derivatives of bits = -log2(max(|Phi_hi - Phi_lo|, 1e-6)) * keep   (Low_bound: zero below the bound)
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

// LITE = true: every row of the launch is NOT chosen for the bit-rate term (85 % of the anchors in training,
// scene/gaussian_model.py:1658-1659), so its only path into the context MLP is through the three adaptive
// quantisation steps (x_q = x + n Q): 3 of the 175 outputs.  The second-layer product, the W2 gradient and the
// W2^T back-projection shrink from 175 to 3 columns and the likelihood derivative disappears.
template <int K1, bool LITE>
__global__ void __launch_bounds__(kMlpThreads, 1) context_level_backward_kernel(Args A)
{
    using SM = Smem<K1>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM &S = *reinterpret_cast<SM *>(smem_raw);
    const int tid = threadIdx.x;
    const int num_tiles = (A.n_rows + kTM - 1) / kTM;
    copy_to_smem(S.w, A.packed_w, SM::kWFloats);
    if (tid < SM::kXRows) S.zero[tid] = 0.f;
    if (tid == 0) S.tile = (int)atomicAdd(A.ticket, 1u);
    __syncthreads();
    const float wbits = A.g_bits_dev ? __ldg(A.g_bits_dev) * A.bits_factor : 0.f;

    // register-resident weight-gradient blocks: W2 (7 outputs x 10 hidden), W1 (8 inputs x 4 hidden)
    float gw2[7][10], gw1[8][4];
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 10; ++j) gw2[i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) gw1[i][j] = 0.f;
    float gb2 = 0.f, gb1 = 0.f;
    const int w2_nb = tid % 25, w2_hb = tid / 25;   // valid when tid < 250
    const int w1_ib = tid % 9, w1_hb = tid / 9;     // valid when tid < 225

    for (int tile = S.tile; tile < num_tiles; tile = S.tile) {
        const int row0 = tile * kTM;
        // ---- stage inputs exactly as the forward ------------------------------------------------
        {
            const int r = tid >> 2, q = tid & 3;
            const int lrow_i = row0 + r;
            if (lrow_i < A.n_rows) {
                const int row = A.row_list ? A.row_list[lrow_i] : lrow_i;
                const int o = A.orig_idx[row];
                if (q == 0) {
                    S.orig[r] = o;
                    S.lrow[r] = row;
                    S.chosen[r] = LITE ? 0 : (A.choose ? A.choose[o] : 1);
                }
                if (K1 == kCtx + kHyper) {
                    const int s = A.ctx_src[row];
                    if (q == 0) {
                        S.src[r] = s;
                        for (int k = 0; k < 3; ++k) S.x[k * kTMp + r] = A.anchor[3 * (size_t)s + k];
                    }
                    const float *fq = A.feat_q + (size_t)s * kCF;
                    for (int k = q; k < kCF; k += 4) S.x[(3 + k) * kTMp + r] = fq[k];
                    if (q == 1)
                        for (int k = 0; k < kCS; ++k) S.x[(3 + kCF + k) * kTMp + r] = A.scaling_q[(size_t)s * kCS + k];
                } else {
                    if (q == 0) {
                        S.src[r] = o;
                        for (int k = 0; k < 3; ++k) S.x[k * kTMp + r] = A.level_anchor[3 * (size_t)row + k];
                    }
                }
                const float *hq = A.hyper_q + (size_t)o * kHyper;
                for (int k = q; k < kHyper; k += 4) S.x[(K1 - kHyper + k) * kTMp + r] = hq[k];
            } else {
                if (q == 0) {
                    S.orig[r] = -1;
                    S.src[r] = -1;
                    S.lrow[r] = 0;
                    S.chosen[r] = 0;
                }
                for (int k = q; k < K1; k += 4) S.x[k * kTMp + r] = 0.f;
            }
            if (tid < 3 * kTM) S.dQ[tid] = 0.f;
        }
        __syncthreads();
        // ---- recompute the forward activations ----------------------------------------------------
        tile_gemm<4, ACT_RELU>(S.x, K1, S.w + SM::kW1, kLd1, S.w + SM::kB1, kGH, S.h);
        __syncthreads();
        if (!LITE) {
            tile_gemm<6, ACT_NONE>(S.h, kGH, S.w + SM::kW2, kLd2, S.w + SM::kB2, kGO, S.out);
        } else {
            // only the three step outputs (columns 172..174); columns 168..171 share their weight-gradient block
            // and must read as zero
            if (tid < 3 * kTM) {
                const int g = tid / kTM, r = tid - g * kTM;
                const int n = 2 * kCE + g;
                float acc = S.w[SM::kB2 + n];
                for (int hh = 0; hh < kGH; ++hh) acc = fmaf(S.h[hh * kTMp + r], S.w[SM::kW2 + hh * kLd2 + n], acc);
                S.out[n * kTMp + r] = acc;
            } else if (tid < 4 * kTM) {
                const int r = tid - 3 * kTM;
#pragma unroll
                for (int n = 168; n < 172; ++n) S.out[n * kTMp + r] = 0.f;
                S.out[175 * kTMp + r] = 0.f;
            }
        }
        __syncthreads();
        if (tid < 3 * kTM) {
            const int g = tid / kTM, r = tid - g * kTM;
            S.Q[g * kTM + r] = fmaxf(q0_of(g) * (1.0f + tanhf(S.out[(2 * kCE + g) * kTMp + r])), 1e-9f);
        }
        __syncthreads();

        // ---- per coded value: gradient of x_q, mean, scale, Q ----------------------------------------
        if (LITE) {
            // no bit-rate term on these rows: d x = d x_q (G stays as it is) and dQ_g = sum_j noise_j * G_j is a pure
            // reduction -- four threads per row, no store between the loads, so all of them are in flight at once
            const int r = tid >> 2, q = tid & 3;
            const int o = S.orig[r];
            float dq0 = 0.f, dq1 = 0.f, dq2 = 0.f;
            if (o >= 0) {
                const float *nz = A.noise + (size_t)S.lrow[r] * kCE;
                const float *gf = A.G_feat + (size_t)o * kCF, *gs = A.G_scaling + (size_t)o * kCS;
                const float *go = A.G_offsets + (size_t)o * kCO;
                for (int j = q; j < kCF; j += 4) dq0 = fmaf(nz[j], gf[j], dq0);
                for (int j = q; j < kCS; j += 4) dq1 = fmaf(nz[kCF + j], gs[j], dq1);
                for (int j = q; j < kCO; j += 4) dq2 = fmaf(nz[kCF + kCS + j], go[j], dq2);
            }
#pragma unroll
            for (int sft = 1; sft < 4; sft <<= 1) {
                dq0 += __shfl_xor_sync(0xffffffffu, dq0, sft);
                dq1 += __shfl_xor_sync(0xffffffffu, dq1, sft);
                dq2 += __shfl_xor_sync(0xffffffffu, dq2, sft);
            }
            if (q == 0) {
                S.dQ[0 * kTM + r] = dq0;
                S.dQ[1 * kTM + r] = dq1;
                S.dQ[2 * kTM + r] = dq2;
            }
        } else
        for (int e = tid; e < kTM * kCE; e += kMlpThreads) {
            const int r = e / kCE, j = e - r * kCE;
            const int o = S.orig[r];
            int grp, mrow, srow;
            float *G;
            const float *XQ;
            float x_mean, keep = 1.f;
            if (j < kCF) {
                grp = 0; mrow = j; srow = kCF + j; x_mean = A.feat_mean;
                G = A.G_feat + (size_t)o * kCF + j; XQ = A.feat_q + (size_t)o * kCF + j;
            } else if (j < kCF + kCS) {
                const int s = j - kCF;
                grp = 1; mrow = 2 * kCF + s; srow = 2 * kCF + kCS + s; x_mean = A.scaling_mean;
                G = A.G_scaling + (size_t)o * kCS + s; XQ = A.scaling_q + (size_t)o * kCS + s;
            } else {
                const int t = j - kCF - kCS;
                grp = 2; mrow = 2 * kCF + 2 * kCS + t; srow = 2 * kCF + 2 * kCS + kCO + t; x_mean = A.offset_mean;
                G = A.G_offsets + (size_t)o * kCO + t; XQ = A.offsets_q + (size_t)o * kCO + t;
            }
            float d_mean = 0.f, d_scale = 0.f;
            if (o >= 0) {
                float gx_total = *G;
                float dq = 0.f;
                if (S.chosen[r] && wbits != 0.f) {
                    if (grp == 2) keep = A.mask[(size_t)o * 10 + (j - kCF - kCS) / 3];
                    float bits, gx, gm, gs, gq;
                    bits_grad(*XQ, S.out[mrow * kTMp + r], S.out[srow * kTMp + r], S.Q[grp * kTM + r], x_mean, bits, gx,
                              gm, gs, gq);
                    const float wk = wbits * keep;
                    gx_total += wk * gx;
                    d_mean = wk * gm;
                    d_scale = wk * gs;
                    dq = wk * gq;
                    if (grp == 2) atomicAdd(A.d_mask + (size_t)o * 10 + (j - kCF - kCS) / 3, wbits * bits);
                }
                *G = gx_total;  // x_q = x + n Q  ->  d x = d x_q
                dq += A.noise[(size_t)S.lrow[r] * kCE + j] * gx_total;
                if (dq != 0.f) atomicAdd(&S.dQ[grp * kTM + r], dq);
            }
            if (!LITE) {
                S.out[mrow * kTMp + r] = d_mean;
                S.out[srow * kTMp + r] = d_scale;
            }
        }
        __syncthreads();
        if (tid < 3 * kTM) {
            const int g = tid / kTM, r = tid - g * kTM;
            const float t = tanhf(S.out[(2 * kCE + g) * kTMp + r]);
            const bool pass = q0_of(g) * (1.0f + t) >= 1e-9f;  // .clamp(1e-9)
            S.out[(2 * kCE + g) * kTMp + r] = pass ? S.dQ[g * kTM + r] * q0_of(g) * (1.0f - t * t) : 0.f;
        } else if (tid < 4 * kTM) {
            S.out[175 * kTMp + (tid - 3 * kTM)] = 0.f;
        }
        __syncthreads();

        // ---- dW2 += h (x) d_out, db2 -------------------------------------------------------------------
        if (tid < 250 && (!LITE || w2_nb == 24)) outer_accumulate<7, 10>(S.out, 7 * w2_nb, kGO, S.h, 10 * w2_hb, kGH, gw2);
        if (tid < kGO && (!LITE || tid >= 2 * kCE)) {
            float s = 0.f;
            for (int r = 0; r < kTM; ++r) s += S.out[tid * kTMp + r];
            gb2 += s;
        }
        __syncthreads();
        // ---- d_h = (W2^T d_out) * relu'(h), in place ------------------------------------------------------
        if (!LITE) {
            tile_gemm_relu_mask<4, true>(S.out, kGO, S.w + SM::kW2, kLd2, kGH, S.h);
        } else {
            for (int e = tid; e < kGH * kTM; e += kMlpThreads) {
                const int hh = e / kTM, r = e - hh * kTM;
                float acc = 0.f;
#pragma unroll
                for (int g = 0; g < 3; ++g) acc = fmaf(S.out[(2 * kCE + g) * kTMp + r], S.w[SM::kW2 + hh * kLd2 + 2 * kCE + g], acc);
                S.h[hh * kTMp + r] = S.h[hh * kTMp + r] > 0.f ? acc : 0.f;
            }
        }
        __syncthreads();
        // ---- dW1 += x (x) d_h, db1 ---------------------------------------------------------------------------
        if (tid < 225 && 8 * w1_ib < K1) outer_accumulate<8, 4>(S.x, 8 * w1_ib, K1, S.h, 4 * w1_hb, kGH, gw1);
        if (tid < kGH) {
            float s = 0.f;
            for (int r = 0; r < kTM; ++r) s += S.h[tid * kTMp + r];
            gb1 += s;
        }
        __syncthreads();
        // ---- d_x = W1^T d_h, in place over x ------------------------------------------------------------------
        tile_gemm<3, ACT_NONE, true>(S.h, kGH, S.w + SM::kW1, kLd1, S.zero, K1, S.x);
        __syncthreads();
        // ---- scatter-add onto the context sources (coarser level) ---------------------------------------------
        {
            const int r = tid >> 2, q = tid & 3;
            const int o = S.orig[r], s = S.src[r];
            if (o >= 0) {
                if (q == 0)
                    for (int k = 0; k < 3; ++k) atomicAdd(A.d_anchor + 3 * (size_t)s + k, S.x[k * kTMp + r]);
                if (K1 == kCtx + kHyper) {
                    for (int k = q; k < kCF; k += 4) atomicAdd(A.G_feat + (size_t)s * kCF + k, S.x[(3 + k) * kTMp + r]);
                    if (q == 1)
                        for (int k = 0; k < kCS; ++k)
                            atomicAdd(A.G_scaling + (size_t)s * kCS + k, S.x[(3 + kCF + k) * kTMp + r]);
                }
                for (int k = q; k < kHyper; k += 4) A.d_hyper_q[(size_t)o * kHyper + k] = S.x[(K1 - kHyper + k) * kTMp + r];
            }
        }
        if (tid == 0) S.tile = (int)atomicAdd(A.ticket, 1u);
        __syncthreads();
    }

    // ---- one atomic per weight per CTA -------------------------------------------------------------------------
    if (tid < 250) {
#pragma unroll
        for (int i = 0; i < 7; ++i)
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const int n = 7 * w2_nb + i, hh = 10 * w2_hb + j;
                if (n < kGO && gw2[i][j] != 0.f) atomicAdd(A.d_w + SM::kW2 + hh * kLd2 + n, gw2[i][j]);
            }
    }
    if (tid < 225) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 8 * w1_ib + i, hh = 4 * w1_hb + j;
                if (k < K1 && gw1[i][j] != 0.f) atomicAdd(A.d_w + SM::kW1 + k * kLd1 + hh, gw1[i][j]);
            }
    }
    if (tid < kGO && gb2 != 0.f) atomicAdd(A.d_w + SM::kB2 + tid, gb2);
    if (tid < kGH && gb1 != 0.f) atomicAdd(A.d_w + SM::kB1 + tid, gb1);
}

// ------------------------------------------------------------------------------------ EntropyBottleneck
// Reverse mode through the per-channel cumulative network of eb_logits (context_model.cu): returns
// d logits / d v and accumulates g * d logits / d params into dp[59].
__device__ __forceinline__ float eb_logits_backward(const float *__restrict__ p, float v, float g, float *dp)
{
    // forward with the intermediates kept:  s = pre-activation sum, t = tanh(s)
    float l0[3], t0[3], s[3][3], t[3][3], a[4][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        l0[j] = p[j] * v + p[3 + j];
        t0[j] = eb_tanh(l0[j]);
        a[0][j] = l0[j] + p[6 + j] * t0[j];
    }
    const float *q = p + 9;
#pragma unroll
    for (int layer = 0; layer < 3; ++layer) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            s[layer][i] = q[3 * i] * a[layer][0] + q[3 * i + 1] * a[layer][1] + q[3 * i + 2] * a[layer][2] + q[9 + i];
            t[layer][i] = eb_tanh(s[layer][i]);
            a[layer + 1][i] = s[layer][i] + q[12 + i] * t[layer][i];
        }
        q += 15;
    }
    // last layer: logit = q[0..2] . a[3] + q[3]
    float ga[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        dp[54 + i] += g * a[3][i];
        ga[i] = g * q[i];
    }
    dp[57] += g;
#pragma unroll
    for (int layer = 2; layer >= 0; --layer) {
        q -= 15;
        float gprev[3] = {0.f, 0.f, 0.f};
        const int base = 9 + 15 * layer;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            dp[base + 12 + i] += ga[i] * t[layer][i];                                        // factor
            const float gs = ga[i] * (1.0f + q[12 + i] * (1.0f - t[layer][i] * t[layer][i]));  // d / d s
            dp[base + 9 + i] += gs;                                                          // bias
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dp[base + 3 * i + k] += gs * a[layer][k];                                    // matrix
                gprev[k] += gs * q[3 * i + k];
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) ga[k] = gprev[k];
    }
    float gv = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        dp[6 + j] += ga[j] * t0[j];
        const float gl = ga[j] * (1.0f + p[6 + j] * (1.0f - t0[j] * t0[j]));
        dp[3 + j] += gl;
        dp[j] += gl * v;
        gv += gl * p[j];
    }
    return gv;
}

// d hyper[e] (+)= w * d(-log2 lik)/d hyper_q for chosen anchors; d_params[C,59] += ...
// Every thread stays on ONE channel (the grid stride is a multiple of C), so the 58 parameter gradients of the channel
// accumulate in REGISTERS over all the elements the thread visits and reach shared memory once per thread (the first
// version did 58 shared-memory atomics per element: 0.87 ms for 1.5 M anchors, most of it serialised atomics).
__global__ void __launch_bounds__(192)
eb_backward_kernel(const float *__restrict__ params, int C, const float *__restrict__ hyper_q, int N,
                   const uint8_t *__restrict__ choose, const float *__restrict__ g_bits_dev, float bits_factor,
                   float *__restrict__ d_hyper, float *__restrict__ d_params)
{
    extern __shared__ float sp[];           // params [C*59] | gradient accumulators [C*59]
    float *sg = sp + C * kEbParams;
    for (int i = threadIdx.x; i < C * kEbParams; i += blockDim.x) {
        sp[i] = params[i];
        sg[i] = 0.f;
    }
    __syncthreads();
    const float w = __ldg(g_bits_dev) * bits_factor;
    const size_t total = (size_t)N * C;
    const size_t stride = (size_t)gridDim.x * blockDim.x;       // a multiple of C (host): e % C is constant per thread
    const size_t e0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(e0 % C);
    const float *p = sp + c * kEbParams;
    float dp[kEbParams];
#pragma unroll
    for (int i = 0; i < kEbParams; ++i) dp[i] = 0.f;
    bool any = false;
    for (size_t e = e0; e < total; e += stride) {
        if (w == 0.f || (choose && !choose[e / C])) continue;
        const float out = hyper_q[e];
        float lower, upper;
        {
            float l[3], m[3];
            for (int side = 0; side < 2; ++side) {
                const float v = out + (side ? 0.5f : -0.5f);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    l[j] = p[j] * v + p[3 + j];
                    l[j] += p[6 + j] * eb_tanh(l[j]);
                }
                const float *q = p + 9;
#pragma unroll
                for (int layer = 0; layer < 3; ++layer) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float s = q[3 * i] * l[0] + q[3 * i + 1] * l[1] + q[3 * i + 2] * l[2] + q[9 + i];
                        m[i] = s + q[12 + i] * eb_tanh(s);
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i) l[i] = m[i];
                    q += 15;
                }
                const float lg = q[0] * l[0] + q[1] * l[1] + q[2] * l[2] + q[3];
                if (side) upper = lg; else lower = lg;
            }
        }
        const float sum = lower + upper;
        const float sign = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
        const float a = 1.0f / (1.0f + expf(-sign * upper));
        const float b = 1.0f / (1.0f + expf(-sign * lower));
        const float diff = a - b;
        const float lk = fmaxf(fabsf(diff), 1e-9f);
        // LowerBound passes the gradient when lik >= bound or the incoming gradient is negative;
        // d L / d lik = w * (-1 / (lik ln2)) is negative for w > 0
        const float g_lik = w * (-1.0f / (lk * 0.69314718055994531f));
        if (!(fabsf(diff) >= 1e-9f || g_lik < 0.f)) continue;
        const float sd = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const float g_upper = g_lik * sd * a * (1.0f - a) * sign;
        const float g_lower = -g_lik * sd * b * (1.0f - b) * sign;
        float gv = eb_logits_backward(p, out + 0.5f, g_upper, dp);
        gv += eb_logits_backward(p, out - 0.5f, g_lower, dp);
        d_hyper[e] += gv;
        any = true;
    }
    if (any) {
#pragma unroll
        for (int i = 0; i < kEbParams - 1; ++i)
            if (dp[i] != 0.f) atomicAdd(&sg[c * kEbParams + i], dp[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * kEbParams; i += blockDim.x)
        if (sg[i] != 0.f) atomicAdd(d_params + i, sg[i]);
}

template <int K1>
static int launch_level_backward(const cmb::Args &a, bool lite, cudaStream_t st)
{
    using SM = cmb::Smem<K1>;
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(cmb::context_level_backward_kernel<K1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(SM));
        cudaFuncSetAttribute(cmb::context_level_backward_kernel<K1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(SM));
        if (sm_count <= 0) sm_count = kNumSMs;
    }
    const int tiles = (a.n_rows + kTM - 1) / kTM;
    StageScope sc(ST_CTX_LEVEL_BWD, st, 1);
    if (lite)
        cmb::context_level_backward_kernel<K1, true><<<tiles < sm_count ? tiles : sm_count, kMlpThreads, sizeof(SM), st>>>(a);
    else
        cmb::context_level_backward_kernel<K1, false><<<tiles < sm_count ? tiles : sm_count, kMlpThreads, sizeof(SM), st>>>(a);
    return check_launch("cgs_context_level_backward");
}

}  // namespace cmb
}  // namespace cgs

using namespace cgs;

extern "C" int cgs_context_level_backward_packed_floats(int in_dim)
{
    if (in_dim == 71) return cmb::Smem<71>::kWFloats;
    if (in_dim == 15) return cmb::Smem<15>::kWFloats;
    return -1;
}

extern "C" int cgs_context_level_backward(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                          const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                          const float *anchor, const float *hyper_q, const float *feat_q,
                                          const float *scaling_q, const float *offsets_q, const float *mask,
                                          const uint8_t *choose, const float *noise, float feat_mean, float scaling_mean,
                                          float offset_mean, const float *g_bits_dev, float bits_factor, float *G_feat,
                                          float *G_scaling, float *G_offsets, float *d_mask, float *d_hyper_q,
                                          float *d_anchor, float *d_packed_w, uint32_t *ticket_dev, void *stream)
{
    return cgs_context_level_backward_rows(in_dim, packed_w, orig_idx, ctx_src, level_anchor, nullptr, n_rows, 0, anchor,
                                           hyper_q, feat_q, scaling_q, offsets_q, mask, choose, noise, feat_mean,
                                           scaling_mean, offset_mean, g_bits_dev, bits_factor, G_feat, G_scaling, G_offsets,
                                           d_mask, d_hyper_q, d_anchor, d_packed_w, ticket_dev, stream);
}

extern "C" int cgs_context_level_backward_rows(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                          const int32_t *ctx_src, const float *level_anchor, const int32_t *row_list,
                                          int n_rows, int lite,
                                          const float *anchor, const float *hyper_q, const float *feat_q,
                                          const float *scaling_q, const float *offsets_q, const float *mask,
                                          const uint8_t *choose, const float *noise, float feat_mean, float scaling_mean,
                                          float offset_mean, const float *g_bits_dev, float bits_factor, float *G_feat,
                                          float *G_scaling, float *G_offsets, float *d_mask, float *d_hyper_q,
                                          float *d_anchor, float *d_packed_w, uint32_t *ticket_dev, void *stream)
{
    if (n_rows <= 0) return 0;
    CGS_CHECK_PTR(packed_w); CGS_CHECK_PTR(orig_idx); CGS_CHECK_PTR(anchor); CGS_CHECK_PTR(hyper_q);
    CGS_CHECK_PTR(feat_q); CGS_CHECK_PTR(scaling_q); CGS_CHECK_PTR(offsets_q); CGS_CHECK_PTR(mask);
    CGS_CHECK_PTR(noise); CGS_CHECK_PTR(G_feat); CGS_CHECK_PTR(G_scaling); CGS_CHECK_PTR(G_offsets);
    CGS_CHECK_PTR(d_mask); CGS_CHECK_PTR(d_hyper_q); CGS_CHECK_PTR(d_anchor); CGS_CHECK_PTR(d_packed_w);
    CGS_CHECK_PTR(ticket_dev);
    cmb::Args a;
    a.packed_w = packed_w; a.orig_idx = orig_idx; a.ctx_src = ctx_src; a.level_anchor = level_anchor; a.n_rows = n_rows;
    a.anchor = anchor; a.hyper_q = hyper_q; a.feat_q = feat_q; a.scaling_q = scaling_q; a.offsets_q = offsets_q;
    a.mask = mask; a.choose = choose; a.noise = noise; a.feat_mean = feat_mean; a.scaling_mean = scaling_mean;
    a.offset_mean = offset_mean; a.g_bits_dev = g_bits_dev; a.bits_factor = bits_factor; a.G_feat = G_feat;
    a.G_scaling = G_scaling; a.G_offsets = G_offsets; a.d_mask = d_mask; a.d_hyper_q = d_hyper_q; a.d_anchor = d_anchor;
    a.d_w = d_packed_w; a.ticket = ticket_dev; a.row_list = row_list;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(ticket_dev, 0, sizeof(uint32_t), st);
    if (in_dim == 71) {
        CGS_CHECK_PTR(ctx_src);
        return cmb::launch_level_backward<71>(a, lite != 0, st);
    }
    if (in_dim == 15) {
        CGS_CHECK_PTR(level_anchor);
        return cmb::launch_level_backward<15>(a, lite != 0, st);
    }
    set_error("%s: unsupported context-MLP input width %d", __func__, in_dim);
    return -2;
}

extern "C" int cgs_eb_backward(const float *packed_params, int C, const float *hyper_q, int N, const uint8_t *choose,
                               const float *g_bits_dev, float bits_factor, float *d_hyper, float *d_packed_params,
                               void *stream)
{
    if (N <= 0) return 0;
    CGS_CHECK_PTR(packed_params); CGS_CHECK_PTR(hyper_q); CGS_CHECK_PTR(g_bits_dev); CGS_CHECK_PTR(d_hyper);
    CGS_CHECK_PTR(d_packed_params);
    if (C <= 0 || C > 64) {
        set_error("%s: unsupported channel count %d", __func__, C);
        return -2;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t total = (size_t)N * C;
    // block size = a multiple of C, so that the grid stride keeps every thread on one channel
    const int block = C <= 192 ? (192 / C) * C : C;
    const int grid = (int)min((size_t)kNumSMs * 8, (total + block - 1) / block);
    StageScope sc(ST_EB, st, 1);
    cmb::eb_backward_kernel<<<grid, block, 2 * C * cmb::kEbParams * sizeof(float), st>>>(
        packed_params, C, hyper_q, N, choose, g_bits_dev, bits_factor, d_hyper, d_packed_params);
    return check_launch(__func__);
}

// Anchor-level context / entropy model kernels.
//
//  * eb_forward_kernel         : EntropyBottleneck.forward on the hyper latents (E4)
//  * context_level_kernel<K1>  : ONE fused kernel per level of the 3-level autoregression (E5+E6+E7):
//        gather coarser-level context -> context MLP (71|15 -> 100 ReLU -> 175) -> adaptive
//        quantisation steps -> quantise (STE round, or add uniform noise) -> scatter the
//        quantised attributes -> discretised-Gaussian likelihood -> per-level bit sums.
//    The reference runs ~45 PyTorch kernels per level plus three more whole-array
//    Entropy_gaussian passes (scene/gaussian_model.py:1562-1670); here (mu, sigma, Q) never
//    leave shared memory.
//  * gaussian_bits_{forward,backward}, ste_multistep, quantize_anchor : stand-alone elementwise
//    kernels behind the drop-in utils.entropy_models / utils.encodings surface.
#include <algorithm>

#include "entropy_math.cuh"
#include "mlp_tile.cuh"

namespace cgs {

// ------------------------------------------------------------------------------------ E4
__device__ __forceinline__ float eb_logits(const float *__restrict__ p, float v)
{
    float l[3], m[3];
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
            float s = q[3 * i] * l[0] + q[3 * i + 1] * l[1] + q[3 * i + 2] * l[2] + q[9 + i];
            m[i] = s + q[12 + i] * eb_tanh(s);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) l[i] = m[i];
        q += 15;
    }
    return q[0] * l[0] + q[1] * l[1] + q[2] * l[2] + q[3];
}

// hyper [N,C] -> hyper_q [N,C], likelihood [N,C].  noise (training) is [N,C] or null (eval: round
// about the per-channel median).  The block size is a multiple of C and the grid stride a multiple of the block, so a
// thread stays on ONE channel: its 59 parameters live in registers (from shared memory they cost 116 LDS per latent and
// bound the kernel) and consecutive threads still touch consecutive addresses.
__global__ void __launch_bounds__(256)
eb_forward_kernel(const float *__restrict__ params, int C, const float *__restrict__ hyper,
                  const float *__restrict__ noise, int N, float *__restrict__ hyper_q, float *__restrict__ lik,
                  const uint8_t *__restrict__ choose, double *bit_sum)
{
    float local_bits = 0.f;
    const int c = threadIdx.x % C;
    float p[kEbParams];
#pragma unroll
    for (int i = 0; i < kEbParams; ++i) p[i] = params[c * kEbParams + i];
    const size_t total = (size_t)N * C;
    const uint32_t used = blockDim.x / C * C;                // threads of the block that work (the rest only join the reduction)
    uint32_t row = (blockIdx.x * used + threadIdx.x) / C;   // used % C == 0: the row advances uniformly
    const uint32_t row_step = gridDim.x * used / C;
    for (size_t e = threadIdx.x < used ? (size_t)blockIdx.x * used + threadIdx.x : total; e < total;
         e += (size_t)gridDim.x * used, row += row_step) {
        const float x = hyper[e];
        float out;
        if (noise) {
            out = x + noise[e];
        } else {
            const float med = p[kEbParams - 1];
            out = rintf(x - med) + med;
        }
        const float lower = eb_logits(p, out - 0.5f);
        const float upper = eb_logits(p, out + 0.5f);
        const float sum = lower + upper;
        const float sign = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
        const float a = 1.0f / (1.0f + expf(-sign * upper));
        const float b = 1.0f / (1.0f + expf(-sign * lower));
        hyper_q[e] = out;
        const float lk = fmaxf(fabsf(a - b), 1e-9f);
        lik[e] = lk;
        if (bit_sum && (!choose || choose[row])) local_bits += -log2f(lk);
    }
    if (bit_sum) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_bits += __shfl_xor_sync(0xffffffffu, local_bits, o);
        if ((threadIdx.x & 31) == 0 && local_bits != 0.f) atomicAdd(bit_sum, (double)local_bits);
    }
}

// ------------------------------------------------------------------------------------ E5-E7
// packed context-MLP weights (floats): W1[K1][100] | b1[100] | W2[100][176] | b2[176]
template <int K1>
struct LevelSmem {
    static constexpr int kW1 = 0, kB1 = K1 * kGH, kW2 = kB1 + kGH, kB2 = kW2 + kGH * kLdG2, kWFloats = kB2 + kLdG2;
    float w[kWFloats];
    float x[K1 * kTMp];
    float h[kGH * kTMp];
    float out[kGO * kTMp];
    float Q[3 * kTM];
    int orig[kTM];
    uint8_t chosen[kTM];
    float red[3 * (kMlpThreads / 32)];
};

struct LevelArgs {
    const float *packed_w;
    const int *orig_idx;      // [n_rows] original anchor index of each coded row
    const int *ctx_src;       // [n_rows] representative anchor of each ROW's context (K1 == 71)
    const float *level_anchor; // [n_rows,3] hybrid anchors (K1 == 15)
    int n_rows;
    const float *anchor;      // [N,3] (context source positions)
    const float *hyper_q;     // [N,12]
    const float *feat, *scaling, *offsets;  // [N,50] [N,6] [N,30] unquantised attributes
    const float *mask;        // [N,10] binary offset masks (bits of masked offsets are dropped)
    const uint8_t *choose;    // [N] anchors whose bits are accumulated (null = all)
    const float *noise;       // [n_rows,86] uniform(-.5,.5) noise in level-row order, or null = STE rounding
    float feat_mean, scaling_mean, offset_mean;
    float *feat_q, *scaling_q, *offsets_q;  // [N,*] quantised attributes, scattered by original index
    float *bits_out;          // optional [N,86] per-element bits (0 where not chosen)
    double *bit_sums;         // [4]: feat, scaling, offsets bit sums and chosen-row count of this level
};

template <int K1>
__global__ void __launch_bounds__(kMlpThreads, 1) context_level_kernel(LevelArgs A)
{
    using SM = LevelSmem<K1>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM &S = *reinterpret_cast<SM *>(smem_raw);
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * kTM;
    copy_to_smem(S.w, A.packed_w, SM::kWFloats);

    // ---- stage inputs (4 threads per row) --------------------------------------------------
    {
        const int r = tid >> 2, q = tid & 3;
        const int row = row0 + r;
        if (row < A.n_rows) {
            const int o = A.orig_idx[row];
            if (q == 0) {
                S.orig[r] = o;
                S.chosen[r] = A.choose ? A.choose[o] : 1;
            }
            if (K1 == kCtx + kHyper) {
                const int s = A.ctx_src[row];
                if (q == 0)
                    for (int k = 0; k < 3; ++k) S.x[k * kTMp + r] = A.anchor[3 * (size_t)s + k];
                const float *fq = A.feat_q + (size_t)s * kCF;
                for (int k = q; k < kCF; k += 4) S.x[(3 + k) * kTMp + r] = fq[k];
                if (q == 1)
                    for (int k = 0; k < kCS; ++k) S.x[(3 + kCF + k) * kTMp + r] = A.scaling_q[(size_t)s * kCS + k];
            } else {
                if (q == 0)
                    for (int k = 0; k < 3; ++k) S.x[k * kTMp + r] = A.level_anchor[3 * (size_t)row + k];
            }
            const float *hq = A.hyper_q + (size_t)o * kHyper;
            for (int k = q; k < kHyper; k += 4) S.x[(K1 - kHyper + k) * kTMp + r] = hq[k];
        } else {
            if (q == 0) {
                S.orig[r] = -1;
                S.chosen[r] = 0;
            }
            for (int k = q; k < K1; k += 4) S.x[k * kTMp + r] = 0.f;
        }
    }
    __syncthreads();

    tile_gemm<4, ACT_RELU>(S.x, K1, S.w + SM::kW1, kGH, S.w + SM::kB1, kGH, S.h);
    __syncthreads();
    tile_gemm<6, ACT_NONE>(S.h, kGH, S.w + SM::kW2, kLdG2, S.w + SM::kB2, kGO, S.out);
    __syncthreads();

    // ---- adaptive quantisation steps (scene/gaussian_model.py:1606-1608) --------------------
    if (tid < 3 * kTM) {
        const int g = tid / kTM, r = tid - g * kTM;
        const float q0 = g == 0 ? kQf0 : (g == 1 ? kQs0 : kQo0);
        S.Q[g * kTM + r] = fmaxf(q0 * (1.0f + tanhf(S.out[(2 * kCE + g) * kTMp + r])), 1e-9f);
    }
    __syncthreads();

    // ---- quantise + scatter + likelihood -----------------------------------------------------
    float sum_f = 0.f, sum_s = 0.f, sum_o = 0.f;
    for (int e = tid; e < kTM * kCE; e += kMlpThreads) {
        const int r = e / kCE, j = e - r * kCE;
        const int o = S.orig[r];
        if (o < 0) continue;
        float x, mean, scale, Q, x_mean, keep = 1.f;
        float *dst;
        int grp;
        if (j < kCF) {
            x = A.feat[(size_t)o * kCF + j];
            mean = S.out[j * kTMp + r];
            scale = S.out[(kCF + j) * kTMp + r];
            Q = S.Q[r];
            x_mean = A.feat_mean;
            dst = A.feat_q + (size_t)o * kCF + j;
            grp = 0;
        } else if (j < kCF + kCS) {
            const int s = j - kCF;
            x = A.scaling[(size_t)o * kCS + s];
            mean = S.out[(2 * kCF + s) * kTMp + r];
            scale = S.out[(2 * kCF + kCS + s) * kTMp + r];
            Q = S.Q[kTM + r];
            x_mean = A.scaling_mean;
            dst = A.scaling_q + (size_t)o * kCS + s;
            grp = 1;
        } else {
            const int t = j - kCF - kCS;
            x = A.offsets[(size_t)o * kCO + t];
            mean = S.out[(2 * kCF + 2 * kCS + t) * kTMp + r];
            scale = S.out[(2 * kCF + 2 * kCS + kCO + t) * kTMp + r];
            Q = S.Q[2 * kTM + r];
            x_mean = A.offset_mean;
            dst = A.offsets_q + (size_t)o * kCO + t;
            keep = A.mask[(size_t)o * 10 + t / 3];
            grp = 2;
        }
        float xq;
        if (A.noise)
            xq = x + A.noise[(size_t)(row0 + r) * kCE + j] * Q;
        else
            xq = ste_round(x, Q);
        *dst = xq;
        float bits = 0.f;
        if (S.chosen[r]) {
            bits = gaussian_bits_one(xq, mean, scale, Q, x_mean) * keep;
            if (grp == 0) sum_f += bits;
            else if (grp == 1) sum_s += bits;
            else sum_o += bits;
        }
        if (A.bits_out) A.bits_out[(size_t)o * kCE + j] = bits;
    }
    // block reduction -> one fp64 atomic per sum per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum_f += __shfl_xor_sync(0xffffffffu, sum_f, o);
        sum_s += __shfl_xor_sync(0xffffffffu, sum_s, o);
        sum_o += __shfl_xor_sync(0xffffffffu, sum_o, o);
    }
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) {
        S.red[warp] = sum_f;
        S.red[8 + warp] = sum_s;
        S.red[16 + warp] = sum_o;
    }
    __syncthreads();
    if (tid < 4) {
        double v = 0.0;
        if (tid < 3)
            for (int w = 0; w < kMlpThreads / 32; ++w) v += (double)S.red[8 * tid + w];
        else
            for (int r = 0; r < kTM; ++r) v += S.chosen[r] ? 1.0 : 0.0;
        if (v != 0.0) atomicAdd(A.bit_sums + tid, v);
    }
}

// ------------------------------------------------------------------------------------ stand-alone elementwise
// x, mean, scale [n, D]; Q [n] (per row) ; bits [n, D]
__global__ void __launch_bounds__(256)
gaussian_bits_forward_kernel(const float *__restrict__ x, const float *__restrict__ mean,
                             const float *__restrict__ scale, const float *__restrict__ Q, int q_per_elem,
                             float x_mean, size_t n, int D, float *__restrict__ bits)
{
    const size_t total = n * D;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const float q = q_per_elem ? Q[e] : Q[e / D];
        bits[e] = gaussian_bits_one(x[e], mean[e], scale[e], q, x_mean);
    }
}

// Backward of bits = -log2(max(|Phi_hi - Phi_lo|, 1e-6)).  Low_bound.backward is g * (lik >= 1e-6)
// (utils/entropy_models.py:149-156, quirk Q2).  dQ is reduced per row when Q is per row.
__global__ void __launch_bounds__(256)
gaussian_bits_backward_kernel(const float *__restrict__ x, const float *__restrict__ mean,
                              const float *__restrict__ scale, const float *__restrict__ Q, int q_per_elem,
                              float x_mean, size_t n, int D, const float *__restrict__ g, float *__restrict__ dx,
                              float *__restrict__ dmean, float *__restrict__ dscale, float *__restrict__ dQ)
{
    const size_t total = n * D;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const float q = q_per_elem ? Q[e] : Q[e / D];
        const float x0 = x[e], mu = mean[e], s0 = scale[e];
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
        float gx = 0.f, gm = 0.f, gs = 0.f, gq = 0.f;
        if (lk >= 1e-6f) {
            const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
            const float dbits_dlk = -1.0f / (lk * 0.69314718055994531f);
            const float c = 0.3989422804014327f * inv;  // 1/(sigma sqrt(2 pi))
            const float ph = c * expf(-zh * zh), pl = c * expf(-zl * zl);
            const float gl = g[e] * dbits_dlk * sg;
            gx = x_pass ? gl * (ph - pl) : 0.f;
            gm = -gl * (ph - pl);
            gq = gl * 0.5f * (ph + pl);
            gs = s_pass ? gl * (-(ph * dh - pl * dl) * inv) : 0.f;
        }
        dx[e] = gx;
        dmean[e] = gm;
        dscale[e] = gs;
        if (q_per_elem) dQ[e] = gq;
        else if (gq != 0.f) atomicAdd(dQ + e / D, gq);
    }
}

__global__ void __launch_bounds__(256)
ste_multistep_kernel(const float *__restrict__ x, const float *__restrict__ Q, size_t n, int D,
                     float *__restrict__ out)
{
    const size_t total = n * D;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
        out[e] = ste_round(x[e], Q[e / D]);
}

// utils/encodings.py:219-227 with torch's floor-division algorithm (c10::div_floor_floating).
__global__ void __launch_bounds__(256)
quantize_anchor_kernel(const float *__restrict__ a, float3 mn, float3 mx, size_t n, float *__restrict__ out,
                       float *__restrict__ qv)
{
    const float qa = 1.0f / 65535.0f;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * 3; e += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % 3);
        const float lo = c == 0 ? mn.x : (c == 1 ? mn.y : mn.z), hi = c == 0 ? mx.x : (c == 1 ? mx.y : mx.z);
        const float interval = (hi - lo) * qa + 1e-6f;
        const float num = a[e] - lo;
        const float mod = fmodf(num, interval);
        float div = __fdiv_rn(num - mod, interval);
        if (mod != 0.f && ((interval < 0.f) != (mod < 0.f))) div -= 1.f;
        float fl;
        if (div != 0.f) {
            fl = floorf(div);
            if (div - fl > 0.5f) fl += 1.f;
        } else {
            fl = copysignf(0.f, __fdiv_rn(num, interval));
        }
        fl = fminf(fmaxf(fl, 0.f), 65535.f);
        qv[e] = fl;
        out[e] = fl * interval + lo;
    }
}

static int ew_grid(size_t total) { return (int)min((size_t)kNumSMs * 16, (total + 255) / 256); }

template <int K1>
static int launch_level(const LevelArgs &a, cudaStream_t st)
{
    using SM = LevelSmem<K1>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(context_level_kernel<K1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
        attr_set = true;
    }
    StageScope sc(ST_CTX_LEVEL, st, 1);
    context_level_kernel<K1><<<(a.n_rows + kTM - 1) / kTM, kMlpThreads, sizeof(SM), st>>>(a);
    return check_launch("cgs_context_level_forward");
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_eb_param_floats(void) { return kEbParams; }

extern "C" int cgs_eb_forward(const float *packed_params, int C, const float *hyper, const float *noise, int N,
                              float *hyper_q, float *likelihood, const uint8_t *choose, double *bit_sum, void *stream)
{
    if (N <= 0) return 0;
    CGS_CHECK_PTR(packed_params);
    CGS_CHECK_PTR(hyper);
    CGS_CHECK_PTR(hyper_q);
    CGS_CHECK_PTR(likelihood);
    if (C <= 0 || C > 64) {
        set_error("%s: unsupported channel count %d", __func__, C);
        return -2;
    }
    StageScope sc(ST_EB, static_cast<cudaStream_t>(stream), 1);
    const int used = 256 / C * C;   // working threads per block, a multiple of C: every thread keeps one channel
    const size_t want = ((size_t)N * C + used - 1) / used;
    eb_forward_kernel<<<(unsigned)std::min<size_t>(want, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        packed_params, C, hyper, noise, N, hyper_q, likelihood, choose, bit_sum);
    return check_launch(__func__);
}

extern "C" int cgs_context_level_packed_floats(int in_dim)
{
    if (in_dim == kCtx + kHyper) return LevelSmem<kCtx + kHyper>::kWFloats;
    if (in_dim == 3 + kHyper) return LevelSmem<3 + kHyper>::kWFloats;
    return -1;
}

extern "C" int cgs_context_level_forward(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                         const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                         const float *anchor, const float *hyper_q, const float *feat,
                                         const float *scaling, const float *offsets, const float *mask,
                                         const uint8_t *choose, const float *noise, float feat_mean,
                                         float scaling_mean, float offset_mean, float *feat_q, float *scaling_q,
                                         float *offsets_q, float *bits_out, double *bit_sums, void *stream)
{
    if (n_rows <= 0) return 0;
    CGS_CHECK_PTR(packed_w);
    CGS_CHECK_PTR(orig_idx);
    CGS_CHECK_PTR(anchor);
    CGS_CHECK_PTR(hyper_q);
    CGS_CHECK_PTR(feat);
    CGS_CHECK_PTR(scaling);
    CGS_CHECK_PTR(offsets);
    CGS_CHECK_PTR(mask);
    CGS_CHECK_PTR(feat_q);
    CGS_CHECK_PTR(scaling_q);
    CGS_CHECK_PTR(offsets_q);
    CGS_CHECK_PTR(bit_sums);
    LevelArgs a;
    a.packed_w = packed_w; a.orig_idx = orig_idx; a.ctx_src = ctx_src; a.level_anchor = level_anchor;
    a.n_rows = n_rows; a.anchor = anchor; a.hyper_q = hyper_q; a.feat = feat; a.scaling = scaling;
    a.offsets = offsets; a.mask = mask; a.choose = choose; a.noise = noise; a.feat_mean = feat_mean;
    a.scaling_mean = scaling_mean; a.offset_mean = offset_mean; a.feat_q = feat_q; a.scaling_q = scaling_q;
    a.offsets_q = offsets_q; a.bits_out = bits_out; a.bit_sums = bit_sums;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_dim == kCtx + kHyper) {
        CGS_CHECK_PTR(ctx_src);
        return launch_level<kCtx + kHyper>(a, st);
    }
    if (in_dim == 3 + kHyper) {
        CGS_CHECK_PTR(level_anchor);
        return launch_level<3 + kHyper>(a, st);
    }
    set_error("%s: unsupported context-MLP input width %d", __func__, in_dim);
    return -2;
}

extern "C" int cgs_gaussian_bits_forward(const float *x, const float *mean, const float *scale, const float *Q,
                                         int q_per_elem, float x_mean, int64_t n, int D, float *bits, void *stream)
{
    if (n <= 0 || D <= 0) return 0;
    CGS_CHECK_PTR(x); CGS_CHECK_PTR(mean); CGS_CHECK_PTR(scale); CGS_CHECK_PTR(Q); CGS_CHECK_PTR(bits);
    StageScope sc(ST_BITS, static_cast<cudaStream_t>(stream), 1);
    gaussian_bits_forward_kernel<<<ew_grid((size_t)n * D), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, mean, scale, Q, q_per_elem, x_mean, (size_t)n, D, bits);
    return check_launch(__func__);
}

extern "C" int cgs_gaussian_bits_backward(const float *x, const float *mean, const float *scale, const float *Q,
                                          int q_per_elem, float x_mean, int64_t n, int D, const float *grad_bits,
                                          float *dx, float *dmean, float *dscale, float *dQ, void *stream)
{
    if (n <= 0 || D <= 0) return 0;
    CGS_CHECK_PTR(x); CGS_CHECK_PTR(mean); CGS_CHECK_PTR(scale); CGS_CHECK_PTR(Q); CGS_CHECK_PTR(grad_bits);
    CGS_CHECK_PTR(dx); CGS_CHECK_PTR(dmean); CGS_CHECK_PTR(dscale); CGS_CHECK_PTR(dQ);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!q_per_elem) cudaMemsetAsync(dQ, 0, (size_t)n * sizeof(float), st);
    StageScope sc(ST_BITS, st, 1);
    gaussian_bits_backward_kernel<<<ew_grid((size_t)n * D), 256, 0, st>>>(x, mean, scale, Q, q_per_elem, x_mean,
                                                                          (size_t)n, D, grad_bits, dx, dmean, dscale,
                                                                          dQ);
    return check_launch(__func__);
}

extern "C" int cgs_ste_multistep(const float *x, const float *Q, int64_t n, int D, float *out, void *stream)
{
    if (n <= 0 || D <= 0) return 0;
    CGS_CHECK_PTR(x); CGS_CHECK_PTR(Q); CGS_CHECK_PTR(out);
    StageScope sc(ST_ELEMWISE, static_cast<cudaStream_t>(stream), 1);
    ste_multistep_kernel<<<ew_grid((size_t)n * D), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, Q, (size_t)n, D,
                                                                                              out);
    return check_launch(__func__);
}

extern "C" int cgs_quantize_anchor(const float *anchors, const float *min_host, const float *max_host, int64_t n,
                                   float *anchors_q, float *quantized_v, void *stream)
{
    if (n <= 0) return 0;
    CGS_CHECK_PTR(anchors); CGS_CHECK_PTR(min_host); CGS_CHECK_PTR(max_host); CGS_CHECK_PTR(anchors_q);
    CGS_CHECK_PTR(quantized_v);
    StageScope sc(ST_ELEMWISE, static_cast<cudaStream_t>(stream), 1);
    quantize_anchor_kernel<<<ew_grid((size_t)n * 3), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        anchors, make_float3(min_host[0], min_host[1], min_host[2]), make_float3(max_host[0], max_host[1], max_host[2]),
        (size_t)n, anchors_q, quantized_v);
    return check_launch(__func__);
}

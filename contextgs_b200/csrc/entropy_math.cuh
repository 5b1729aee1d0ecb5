// Scalar building blocks of the entropy model shared by the SIMT and the tcgen05 level kernels.
// Files including this header are compiled with -fmad=false: every operation rounds once, like the
// PyTorch expressions of utils/entropy_models.py:34-50 and utils/encodings.py:203-216 they restate.
#pragma once
#include "common.cuh"

namespace cgs {

constexpr int kCF = 50, kCS = 6, kCO = 30, kCE = kCF + kCS + kCO;  // 86 coded values per anchor
constexpr int kCtx = 3 + kCF + kCS;                               // 59
constexpr int kHyper = 12;
constexpr int kGH = 100, kGO = 175, kLdG2 = 176;
constexpr float kQf0 = 1.0f, kQs0 = 0.001f, kQo0 = 0.2f;
constexpr float kClampSteps = 15000.0f;
constexpr int kEbParams = 59;

// ------------------------------------------------------------------------------------ E6 core
__device__ __forceinline__ float normal_cdf(float v, float mean, float inv_scale)
{
    // torch.distributions.Normal.cdf: 0.5 * (1 + erf((v - loc) * scale.reciprocal() / sqrt(2)))
    return 0.5f * (1.0f + erff(__fdiv_rn((v - mean) * inv_scale, 1.41421356237309515f)));
}

__device__ __forceinline__ float gaussian_bits_one(float x, float mean, float scale, float Q, float x_mean)
{
    x = fminf(fmaxf(x, x_mean - kClampSteps * Q), x_mean + kClampSteps * Q);
    scale = fmaxf(scale, 1e-9f);
    const float inv = __frcp_rn(scale);
    const float upper = normal_cdf(x + 0.5f * Q, mean, inv);
    const float lower = normal_cdf(x - 0.5f * Q, mean, inv);
    const float lk = fmaxf(fabsf(upper - lower), 1e-6f);
    return -log2f(lk);
}

__device__ __forceinline__ float ste_round(float x, float Q)
{
    x = fminf(fmaxf(x, -kClampSteps * Q), kClampSteps * Q);
    return rintf(__fdiv_rn(x, Q)) * Q;
}

// same value; `sym` receives the integer step count (the symbol the bitstream codec codes: rint(value / Q))
__device__ __forceinline__ float ste_round_sym(float x, float Q, float &sym)
{
    x = fminf(fmaxf(x, -kClampSteps * Q), kClampSteps * Q);
    sym = rintf(__fdiv_rn(x, Q));
    return sym * Q;
}

// tanh through one ex2 and one reciprocal: absolute error < 5e-7 (tanhf: ~20 instructions and a branch per call, and the
// factorised prior evaluates 24 of them per latent); exact limits +-1 for |x| > 9
__device__ __forceinline__ float eb_tanh(float x)
{
    return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f);
}

}  // namespace cgs

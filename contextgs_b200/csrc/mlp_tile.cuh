// Register-tiled fp32 tile GEMM used by the fused anchor-MLP kernels (neural_gaussians.cu,
// context_model.cu).  A CTA of 256 threads owns TM = 64 rows; activations live in shared memory
// TRANSPOSED (k-major, rows contiguous, row stride kTMp) so that a thread's 8 rows are two
// broadcast LDS.128; weights live k-major with the N outputs contiguous so that a warp's 32
// column groups read 32 consecutive floats (conflict free).  Each thread accumulates an
// 8 x CJ register tile (CJ = ceil(N/32)).
//
// fp32 FMA keeps the 1e-4 relative-L2 parity budget with a wide margin; the tensor-core
// (tcgen05, 3xTF32) variant of these tiles is the planned next step.
#pragma once
#include "common.cuh"

namespace cgs {

constexpr int kTM = 64;        // rows (anchors) per CTA tile
constexpr int kTMp = 68;       // padded row stride of transposed activations (16 B aligned)
constexpr int kMlpThreads = 256;
constexpr int kRT = 8;         // rows per thread

enum { ACT_NONE = 0, ACT_RELU = 1 };

// out[n][r] (stride kTMp) = act(sum_k A[k][r] * W[k][n] + bias[n]),   r < 64, n < N
// A: smem [K][kTMp]; W: smem [K][ldw] ; out: smem [N][kTMp].  All 256 threads must call.
// WT = true reads the weight as W[n][k] (row n, leading dimension ldw) instead of W[k][n]: with an ODD
// ldw the 32 lanes (consecutive n) hit 32 different banks, so one copy of a weight matrix serves both
// the forward product and its transpose in the backward pass.
template <int CJ, int ACT, bool WT = false>
__device__ __forceinline__ void tile_gemm(const float *__restrict__ A, int K, const float *__restrict__ W, int ldw,
                                          const float *__restrict__ bias, int N, float *__restrict__ out)
{
    const int cg = threadIdx.x & 31;         // column group
    const int r0 = (threadIdx.x >> 5) * kRT; // first row of this thread
    float acc[kRT][CJ];
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
        const int n = cg + 32 * j;
        const float b = n < N ? bias[n] : 0.f;
#pragma unroll
        for (int i = 0; i < kRT; ++i) acc[i][j] = b;
    }
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float4 x0 = *reinterpret_cast<const float4 *>(A + k * kTMp + r0);
        const float4 x1 = *reinterpret_cast<const float4 *>(A + k * kTMp + r0 + 4);
        const float x[kRT] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        float w[CJ];
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int n = cg + 32 * j;
            w[j] = n < N ? (WT ? W[n * ldw + k] : W[k * ldw + n]) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kRT; ++i)
#pragma unroll
            for (int j = 0; j < CJ; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
        const int n = cg + 32 * j;
        if (n < N) {
            float v[kRT];
#pragma unroll
            for (int i = 0; i < kRT; ++i) v[i] = ACT == ACT_RELU ? fmaxf(acc[i][j], 0.f) : acc[i][j];
            *reinterpret_cast<float4 *>(out + n * kTMp + r0) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4 *>(out + n * kTMp + r0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

// out[n][r] = (mask_src[n][r] > 0) ? sum_k A[k][r] * W[k][n] : 0 -- GEMM with the ReLU derivative as
// epilogue; `out` may alias `mask_src` (every thread reads exactly the elements it overwrites).
template <int CJ, bool WT = false>
__device__ __forceinline__ void tile_gemm_relu_mask(const float *__restrict__ A, int K, const float *__restrict__ W,
                                                    int ldw, int N, float *out)
{
    const int cg = threadIdx.x & 31;
    const int r0 = (threadIdx.x >> 5) * kRT;
    float acc[kRT][CJ];
#pragma unroll
    for (int j = 0; j < CJ; ++j)
#pragma unroll
        for (int i = 0; i < kRT; ++i) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float4 x0 = *reinterpret_cast<const float4 *>(A + k * kTMp + r0);
        const float4 x1 = *reinterpret_cast<const float4 *>(A + k * kTMp + r0 + 4);
        const float x[kRT] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        float w[CJ];
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
            const int n = cg + 32 * j;
            w[j] = n < N ? (WT ? W[n * ldw + k] : W[k * ldw + n]) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kRT; ++i)
#pragma unroll
            for (int j = 0; j < CJ; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
        const int n = cg + 32 * j;
        if (n < N) {
            float *o = out + n * kTMp + r0;
            const float4 m0 = *reinterpret_cast<const float4 *>(o), m1 = *reinterpret_cast<const float4 *>(o + 4);
            *reinterpret_cast<float4 *>(o) = make_float4(m0.x > 0.f ? acc[0][j] : 0.f, m0.y > 0.f ? acc[1][j] : 0.f,
                                                         m0.z > 0.f ? acc[2][j] : 0.f, m0.w > 0.f ? acc[3][j] : 0.f);
            *reinterpret_cast<float4 *>(o + 4) = make_float4(m1.x > 0.f ? acc[4][j] : 0.f, m1.y > 0.f ? acc[5][j] : 0.f,
                                                             m1.z > 0.f ? acc[6][j] : 0.f, m1.w > 0.f ? acc[7][j] : 0.f);
        }
    }
}

// acc[i][j] += sum_r P[(p0+i)][r] * Q[(q0+j)][r] over the 64 tile rows (outer-product accumulation of a
// weight-gradient block; rows of P / Q are kTMp apart, r contiguous)
template <int NI, int NJ>
__device__ __forceinline__ void outer_accumulate(const float *__restrict__ P, int p0, int pmax,
                                                 const float *__restrict__ Q, int q0, int qmax, float (&acc)[NI][NJ])
{
#pragma unroll 1
    for (int r = 0; r < kTM; r += 4) {
        float4 pv[NI], qv[NJ];
#pragma unroll
        for (int i = 0; i < NI; ++i)
            pv[i] = p0 + i < pmax ? *reinterpret_cast<const float4 *>(P + (p0 + i) * kTMp + r) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            qv[j] = q0 + j < qmax ? *reinterpret_cast<const float4 *>(Q + (q0 + j) * kTMp + r) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                acc[i][j] = fmaf(pv[i].x, qv[j].x, acc[i][j]);
                acc[i][j] = fmaf(pv[i].y, qv[j].y, acc[i][j]);
                acc[i][j] = fmaf(pv[i].z, qv[j].z, acc[i][j]);
                acc[i][j] = fmaf(pv[i].w, qv[j].w, acc[i][j]);
            }
    }
}

// cooperative flat copy global -> shared (n floats, n % 4 == 0, both 16 B aligned)
__device__ __forceinline__ void copy_to_smem(float *dst, const float *__restrict__ src, int n)
{
    const float4 *s = reinterpret_cast<const float4 *>(src);
    float4 *d = reinterpret_cast<float4 *>(dst);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d[i] = __ldg(s + i);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

}  // namespace cgs

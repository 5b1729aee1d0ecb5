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
template <int CJ, int ACT>
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
            w[j] = n < N ? W[k * ldw + n] : 0.f;
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

// cooperative flat copy global -> shared (n floats, n % 4 == 0, both 16 B aligned)
__device__ __forceinline__ void copy_to_smem(float *dst, const float *__restrict__ src, int n)
{
    const float4 *s = reinterpret_cast<const float4 *>(src);
    float4 *d = reinterpret_cast<float4 *>(dst);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d[i] = __ldg(s + i);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

}  // namespace cgs

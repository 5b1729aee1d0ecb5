// Stand-alone check of the tcgen05 building blocks in umma.cuh: one CTA computes
// D[128 x N] = A[128 x K] * W[N x K]^T with the A operand written to TMEM by tcgen05.st, the B
// operand staged in shared memory in the K-major no-swizzle core-matrix layout, tcgen05.mma
// (kind::tf32) issued by one thread, completion through tcgen05.commit -> mbarrier, and the result
// read back with tcgen05.ld.  tests/test_umma_gpu.py compares it with an fp64 product; the fused
// MLP kernels (neural_gaussians_umma.cu) are built from exactly these pieces.
#include "umma.cuh"

namespace cgs {

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float *__restrict__ A, const float *__restrict__ W, int N, int K, int mode,
                     float *__restrict__ D, int32_t *__restrict__ err)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *w_hi = reinterpret_cast<float *>(smem_raw);
    float *w_lo = w_hi + (size_t)N * K;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        umma::mbar_init(&s_bar, 1);
        umma::fence_mbar_init();
    }
    // B operand: element (n, k) -> [(k/4)][n][k%4]
    for (int i = tid; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i - n * K;
        uint32_t hi, lo;
        umma::split_tf32(W[i], hi, lo);
        const int dst = (k >> 2) * (N * 4) + n * 4 + (k & 3);
        w_hi[dst] = __uint_as_float(hi);
        w_lo[dst] = __uint_as_float(lo);
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = s_tmem;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t A_HI = 0, A_LO = 64, D_COL = 128;  // K <= 64, N <= 256

    // A operand: thread `tid` owns row tid
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) umma::split_tf32(A[(size_t)tid * K + k0 + j], hi[j], lo[j]);
        umma::tmem_st8(lane_base + A_HI + k0, hi);
        umma::tmem_st8(lane_base + A_LO + k0, lo);
    }
    umma::tmem_wait_st();
    umma::fence_before_thread_sync();
    __syncthreads();

    if (tid == 0) {
        umma::fence_after_thread_sync();
        if (mode == 0) {  // plain TF32
            const uint32_t idesc = umma::idesc_tf32(128, N);
            const uint32_t lbo = (uint32_t)N * 16u;
            for (int s = 0; s < K / 8; ++s)
                umma::mma_tf32_ts(tbase + D_COL, tbase + A_HI + 8 * s,
                                  umma::smem_desc_kmajor(umma::smem_u32(w_hi) + (uint32_t)s * 2u * lbo, lbo, 128u), idesc,
                                  s > 0 ? 1u : 0u);
        } else {          // modes 1, 2: 3xTF32
            umma::gemm_3xtf32(tbase + D_COL, tbase + A_HI, tbase + A_LO, w_hi, w_lo, N, K, true);
        }
        umma::umma_commit(&s_bar);
    }
    const bool ok = umma::mbar_wait(&s_bar, 0);
    umma::fence_after_thread_sync();
    if (!ok && lane == 0) atomicExch(err, 1);

    // mode 2 reads the accumulator with UNALIGNED column starts (3, 11, 19, ...): the fused kernels read
    // per-anchor column groups that do not start on a multiple of the load width
    const int start = mode == 2 ? 3 : 0;
    if (mode == 2) {
        uint32_t v[8];
        umma::tmem_ld8(lane_base + D_COL, v);
        umma::tmem_wait_ld();
        for (int j = 0; j < 3; ++j) D[(size_t)tid * N + j] = __uint_as_float(v[j]);
    }
    for (int n0 = start; n0 + 8 <= N; n0 += 8) {
        uint32_t v[8];
        umma::tmem_ld8(lane_base + D_COL + n0, v);
        umma::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) D[(size_t)tid * N + n0 + j] = __uint_as_float(v[j]);
    }
    if (mode == 2) {  // tail: the last 5 columns
        uint32_t v[8];
        umma::tmem_ld8(lane_base + D_COL + N - 8, v);
        umma::tmem_wait_ld();
        for (int j = 0; j < 8; ++j) D[(size_t)tid * N + N - 8 + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

// Probe for the weight-gradient GEMMs of the backward kernels: D[M x N] = P^T Q with
// P[128 rows x M], Q[128 rows x N], i.e. the contraction runs over the tile's ROWS.  Both operands go through
// shared memory (SS form) in the K-major core-matrix layout with the row index as K: thread `row` writes its own
// values at ((row / 4) * ld + m) * 4 + row % 4 -- no transposition pass, no MN-major descriptor.  `skew` adds
// 16-byte units to the leading-dimension byte offset (bank spreading of the 8 row groups).  3xTF32.
__global__ void __launch_bounds__(128, 1)
umma_selftest_ss_kernel(const float *__restrict__ P, const float *__restrict__ Q, int M, int N, int skew,
                        float *__restrict__ D, int32_t *__restrict__ err)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lda = 128 + skew, ldb = N + skew;                  // rows of 16 bytes per K chunk
    float *a_hi = reinterpret_cast<float *>(smem_raw);
    float *a_lo = a_hi + (size_t)32 * lda * 4;
    float *b_hi = a_lo + (size_t)32 * lda * 4;
    float *b_lo = b_hi + (size_t)32 * ldb * 4;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        umma::mbar_init(&s_bar, 1);
        umma::fence_mbar_init();
    }
    const int row = tid;                                          // K index of this thread
    for (int m = 0; m < 128; ++m) {
        uint32_t hi = 0, lo = 0;
        if (m < M) umma::split_tf32(P[(size_t)row * M + m], hi, lo);
        const int dst = ((row >> 2) * lda + m) * 4 + (row & 3);
        a_hi[dst] = __uint_as_float(hi);
        a_lo[dst] = __uint_as_float(lo);
    }
    for (int n = 0; n < N; ++n) {
        uint32_t hi, lo;
        umma::split_tf32(Q[(size_t)row * N + n], hi, lo);
        const int dst = ((row >> 2) * ldb + n) * 4 + (row & 3);
        b_hi[dst] = __uint_as_float(hi);
        b_lo[dst] = __uint_as_float(lo);
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = s_tmem;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    if (tid == 0) {
        const uint32_t idesc = umma::idesc_tf32(128, N);
        const uint32_t lbo_a = (uint32_t)lda * 16u, lbo_b = (uint32_t)ldb * 16u;
        for (int s = 0; s < 16; ++s) {                            // K = 128 rows, 8 per instruction
            const uint64_t ahi = umma::smem_desc_kmajor(umma::smem_u32(a_hi) + (uint32_t)s * 2u * lbo_a, lbo_a, 128u);
            const uint64_t alo = umma::smem_desc_kmajor(umma::smem_u32(a_lo) + (uint32_t)s * 2u * lbo_a, lbo_a, 128u);
            const uint64_t bhi = umma::smem_desc_kmajor(umma::smem_u32(b_hi) + (uint32_t)s * 2u * lbo_b, lbo_b, 128u);
            const uint64_t blo = umma::smem_desc_kmajor(umma::smem_u32(b_lo) + (uint32_t)s * 2u * lbo_b, lbo_b, 128u);
            umma::mma_tf32_ss(tbase, ahi, blo, idesc, s > 0 ? 1u : 0u);
            umma::mma_tf32_ss(tbase, alo, bhi, idesc, 1u);
            umma::mma_tf32_ss(tbase, ahi, bhi, idesc, 1u);
        }
        umma::umma_commit(&s_bar);
    }
    const bool ok = umma::mbar_wait(&s_bar, 0);
    umma::fence_after_thread_sync();
    if (!ok && lane == 0) atomicExch(err, 1);
    for (int n0 = 0; n0 + 8 <= N; n0 += 8) {                      // lane `tid` of the accumulator = output row m
        uint32_t v[8];
        umma::tmem_ld8(lane_base + n0, v);
        umma::tmem_wait_ld();
        if (tid < M)
#pragma unroll
            for (int j = 0; j < 8; ++j) D[(size_t)tid * N + n0 + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}


// Same contraction over the tile's ROWS, but with both operands in the MN-MAJOR no-swizzle layout: a 16-byte unit
// holds FOUR CONSECUTIVE FEATURES of ONE row, 8 rows x 16 B form a 128-byte core matrix, so thread `row` stores its
// features as float4 (conflict-free, 4x fewer stores than the K-major form) at [(feature / 4)][row][4].
// variant 0: SBO = stride between 4-feature groups (rows * 16 B), LBO = stride between 8-row groups (128 B)
//            (CUTLASS make_umma_desc<Major::MN>, SWIZZLE_NONE);  variant 1: the two offsets swapped.
__global__ void __launch_bounds__(128, 1)
umma_selftest_ss_mn_kernel(const float *__restrict__ P, const float *__restrict__ Q, int M, int N, int variant,
                           float *__restrict__ D, int32_t *__restrict__ err)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *a_hi = reinterpret_cast<float4 *>(smem_raw);          // [32 groups][128 rows]
    float4 *a_lo = a_hi + 32 * 128;
    float4 *b_hi = a_lo + 32 * 128;                               // [N / 4 groups][128 rows]
    float4 *b_lo = b_hi + (N / 4) * 128;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        umma::mbar_init(&s_bar, 1);
        umma::fence_mbar_init();
    }
    const int row = tid;
    for (int g = 0; g < 32; ++g) {
        float v[4], h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = 4 * g + j;
            v[j] = m < M ? P[(size_t)row * M + m] : 0.f;
            uint32_t hi, lo;
            umma::split_tf32(v[j], hi, lo);
            h[j] = __uint_as_float(hi);
            l[j] = __uint_as_float(lo);
        }
        a_hi[g * 128 + row] = make_float4(h[0], h[1], h[2], h[3]);
        a_lo[g * 128 + row] = make_float4(l[0], l[1], l[2], l[3]);
    }
    for (int g = 0; g < N / 4; ++g) {
        float h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t hi, lo;
            umma::split_tf32(Q[(size_t)row * N + 4 * g + j], hi, lo);
            h[j] = __uint_as_float(hi);
            l[j] = __uint_as_float(lo);
        }
        b_hi[g * 128 + row] = make_float4(h[0], h[1], h[2], h[3]);
        b_lo[g * 128 + row] = make_float4(l[0], l[1], l[2], l[3]);
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = s_tmem;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    if (tid == 0) {
        const uint32_t idesc = umma::idesc_tf32(128, N) | (1u << 15) | (1u << 16);   // A and B MN-major
        const uint32_t grp = 128u * 16u, k8 = 128u;
        const uint32_t lbo = variant == 0 ? k8 : grp, sbo = variant == 0 ? grp : k8;
        for (int s = 0; s < 16; ++s) {                            // K = 128 rows, 8 per instruction
            const uint64_t ahi = umma::smem_desc_kmajor(umma::smem_u32(a_hi) + (uint32_t)s * k8, lbo, sbo);
            const uint64_t alo = umma::smem_desc_kmajor(umma::smem_u32(a_lo) + (uint32_t)s * k8, lbo, sbo);
            const uint64_t bhi = umma::smem_desc_kmajor(umma::smem_u32(b_hi) + (uint32_t)s * k8, lbo, sbo);
            const uint64_t blo = umma::smem_desc_kmajor(umma::smem_u32(b_lo) + (uint32_t)s * k8, lbo, sbo);
            umma::mma_tf32_ss(tbase, ahi, blo, idesc, s > 0 ? 1u : 0u);
            umma::mma_tf32_ss(tbase, alo, bhi, idesc, 1u);
            umma::mma_tf32_ss(tbase, ahi, bhi, idesc, 1u);
        }
        umma::umma_commit(&s_bar);
    }
    const bool ok = umma::mbar_wait(&s_bar, 0);
    umma::fence_after_thread_sync();
    if (!ok && lane == 0) atomicExch(err, 1);
    for (int n0 = 0; n0 + 8 <= N; n0 += 8) {
        uint32_t v[8];
        umma::tmem_ld8(lane_base + n0, v);
        umma::tmem_wait_ld();
        if (tid < M)
#pragma unroll
            for (int j = 0; j < 8; ++j) D[(size_t)tid * N + n0 + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}


// Micro-benchmark: cycles per tcgen05.mma.kind::tf32 (M = 128, K = 8) for a chain of `iters` instructions issued by one
// thread.  form 0: A in TMEM (TS), 1: A in shared memory (SS).  n_acc independent accumulators are used round-robin
// (n_acc = 1: every MMA accumulates into the columns the previous one wrote).  Operand contents are irrelevant (zeros).
__global__ void __launch_bounds__(128, 1)
umma_mma_rate_kernel(int form, int N, int n_acc, int iters, long long *__restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sm = reinterpret_cast<float *>(smem_raw);
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&s_tmem, 512);
    if (tid == 0) {
        umma::mbar_init(&s_bar, form >= 6 ? 4 : (form >= 4 ? 2 : 1));
        umma::fence_mbar_init();
    }
    for (int i = tid; i < 2 * (128 + 256) * 8; i += blockDim.x) sm[i] = 0.f;     // A [2][128][4] + B [2][256][4]
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = s_tmem;
    uint32_t tbase_off = 0;
    // forms 2, 3: canonical issue (warp-uniform branch + elect); forms 4..7: the same from TWO / FOUR warps at once (4, 5: two
    // issuing warps TS / SS; 6, 7: four), each with its own accumulator -- does the ~200-cycle cost per instruction overlap?
    const int n_issuers = form >= 6 ? 4 : (form >= 4 ? 2 : 1);
    const int uw = umma::uniform_warp();
    const bool one = form >= 2 ? (uw < n_issuers && umma::elect_one_sync()) : tid == 0;
    form &= 1;
    if (one) {
        n_acc = 1;
        tbase_off = (uint32_t)uw * 64u;
        const uint32_t idesc = umma::idesc_tf32(128, N);
        const uint32_t a_s = umma::smem_u32(sm), b_s = a_s + 2 * 128 * 16;
        const uint64_t adesc = umma::smem_desc_kmajor(a_s, 128 * 16, 128), bdesc = umma::smem_desc_kmajor(b_s, 256 * 16, 128);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tbase + 256 + tbase_off + (uint32_t)(i % n_acc) * (uint32_t)N;     // accumulators in columns [256,512)
            if (form == 0) umma::mma_tf32_ts(d, tbase, bdesc, idesc, 1u);
            else umma::mma_tf32_ss(d, adesc, bdesc, idesc, 1u);
        }
        umma::umma_commit(&s_bar);
        const long long t1 = clock64();
        umma::mbar_wait(&s_bar, 0);
        const long long t2 = clock64();
        if (uw == 0) {
            out[0] = t1 - t0;      // issue time
            out[1] = t2 - t0;      // issue + completion
        }
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}
}  // namespace cgs

using namespace cgs;

extern "C" int cgs_umma_selftest_ss_mn(const float *P, const float *Q, int M, int N, int variant, float *D, int32_t *err,
                                       void *stream)
{
    CGS_CHECK_PTR(P); CGS_CHECK_PTR(Q); CGS_CHECK_PTR(D); CGS_CHECK_PTR(err);
    if (M < 1 || M > 128 || N < 16 || N > 48 || (N % 16) || variant < 0 || variant > 1) {
        set_error("%s: need 1 <= M <= 128, 16 <= N <= 48 (N %% 16 == 0), variant 0 / 1", __func__);
        return -2;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)2 * 32 * 128 * 16 + (size_t)2 * (N / 4) * 128 * 16;
    cudaFuncSetAttribute(umma_selftest_ss_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemsetAsync(err, 0, sizeof(int32_t), st);
    StageScope sc(ST_ELEMWISE, st, 1);
    umma_selftest_ss_mn_kernel<<<1, 128, smem, st>>>(P, Q, M, N, variant, D, err);
    return check_launch(__func__);
}

extern "C" int cgs_umma_selftest_ss(const float *P, const float *Q, int M, int N, int skew, float *D, int32_t *err,
                                    void *stream)
{
    CGS_CHECK_PTR(P); CGS_CHECK_PTR(Q); CGS_CHECK_PTR(D); CGS_CHECK_PTR(err);
    if (M < 1 || M > 128 || N < 16 || N > 64 || (N % 16) || skew < 0 || skew > 4) {
        set_error("%s: need 1 <= M <= 128, 16 <= N <= 64 (N %% 16 == 0), 0 <= skew <= 4", __func__);
        return -2;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)2 * 32 * (128 + skew) * 16 + (size_t)2 * 32 * (N + skew) * 16;
    cudaFuncSetAttribute(umma_selftest_ss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemsetAsync(err, 0, sizeof(int32_t), st);
    StageScope sc(ST_ELEMWISE, st, 1);
    umma_selftest_ss_kernel<<<1, 128, smem, st>>>(P, Q, M, N, skew, D, err);
    return check_launch(__func__);
}

extern "C" int cgs_umma_selftest(const float *A, const float *W, int N, int K, int mode, float *D, int32_t *err,
                                 void *stream)
{
    CGS_CHECK_PTR(A);
    CGS_CHECK_PTR(W);
    CGS_CHECK_PTR(D);
    CGS_CHECK_PTR(err);
    if (N < 16 || N > 256 || (N % 16) || K < 8 || K > 64 || (K % 8)) {
        set_error("%s: need 16 <= N <= 256 (N %% 16 == 0) and 8 <= K <= 64 (K %% 8 == 0)", __func__);
        return -2;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)2 * N * K * sizeof(float);
    cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemsetAsync(err, 0, sizeof(int32_t), st);
    StageScope sc(ST_ELEMWISE, st, 1);
    umma_selftest_kernel<<<1, 128, smem, st>>>(A, W, N, K, mode, D, err);
    return check_launch(__func__);
}

extern "C" int cgs_umma_mma_rate(int form, int N, int n_acc, int iters, long long *out_cycles, void *stream)
{
    CGS_CHECK_PTR(out_cycles);
    if (form < 0 || form > 7 || (form >= 4 && N > 64) || N < 16 || N > 256 || (N % 16) || n_acc < 1 || n_acc * N > 256 || iters < 1) {
        set_error("%s: invalid arguments", __func__);
        return -2;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = (size_t)2 * (128 + 256) * 8 * sizeof(float);
    cudaFuncSetAttribute(umma_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    StageScope sc(ST_ELEMWISE, st, 1);
    umma_mma_rate_kernel<<<1, 128, smem, st>>>(form, N, n_acc, iters, out_cycles);
    return check_launch(__func__);
}

// Thin inline-PTX layer over Blackwell's 5th-generation tensor cores (tcgen05) for the fused
// anchor-MLP kernels: TMEM allocation, tcgen05.ld/st, single-thread tcgen05.mma (kind::tf32,
// A operand in TMEM, B operand in shared memory through a K-major no-swizzle descriptor),
// tcgen05.commit -> mbarrier, and the 3xTF32 split that keeps fp32-grade accuracy
// (x = hi + lo, x*w ~= hi*w_hi + lo*w_hi + hi*w_lo, relative error ~2^-21).
//
// sm_100a only.  Layout conventions used by every caller:
//   * accumulators and A operands live in TMEM as [128 lanes = tile rows] x [32-bit columns];
//     warp w may touch lanes 32*(w%4) .. +31 only (hardware rule), thread `lane` <-> one row;
//   * B operands (weights W[n][k], k contiguous = "K-major") live in shared memory as
//     [K/4 chunks][N rows][4 floats]: 8 rows x 16 B form one contiguous 128-byte core matrix,
//     SBO (8-row group stride) = 128 B, LBO (stride between the two 16-byte K chunks of one
//     K = 8 instruction) = N * 16 B.
#pragma once
#include "common.cuh"

namespace cgs {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- single-thread issue from a warp-uniform context ------------------------------------------------------
// tcgen05.mma / tcgen05.commit / bulk copies are issued by ONE thread.  Guarding them with `threadIdx.x == 0` makes the
// compiler wrap EVERY such instruction in a vote / ELECT / BRA.U.ANY loop (it cannot prove that one lane is active and that
// the uniform-register operands are warp-uniform): measured 199 cycles per tcgen05.mma regardless of its shape
// (scripts/umma_rate_probe.py).  The canonical form -- a warp-uniform branch on `uniform_warp()` followed by
// `if (elect_one_sync())` -- compiles to back-to-back UTCHMMA.
__device__ __forceinline__ int uniform_warp() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one_sync()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- TMEM allocation: one full warp calls these (.sync.aligned) ---------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor-core operand fetch)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a tensor-core completion that never arrives must not hang the GPU.
// Returns false after ~2^22 polls (seconds); callers raise a device-side error flag.
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t i = 0; i < (1u << 22); ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when they complete
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- descriptors ---------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE, Blackwell version field = 1.
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
// Instruction descriptor: D = fp32, A = B = tf32, both K-major, dense, M x N (M = 128, N % 16 == 0).
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]^T for one K = 8 step; issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T for one K = 8 step (both operands through shared-memory descriptors).
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- TMEM <-> registers (32 lanes x 32 bit, N consecutive columns; whole warp, .sync.aligned) ----
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3])
                 : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t r0, uint32_t r1)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}
// wait::ld that names the destination registers of an earlier tcgen05.ld as in/out operands: the compiler cannot
// move a read of them above the wait (needed when a load is issued ahead of the code that consumes the previous one)
__device__ __forceinline__ void tmem_wait_ld8(uint32_t (&r)[8])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld4(uint32_t (&r)[4])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) : : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- 3xTF32 split ----------------------------------------------------------------------------
// hi = x with the 13 low mantissa bits cleared (what the tensor core would read anyway), lo = the
// same truncation of the exact remainder x - hi: hi + lo carries >= 21 significant bits of x.
// (cvt.rna.tf32.f32 lowers to a ~6-instruction integer sequence on sm_100a; truncation is 1 LOP3.)
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}

// Issue the three TF32 products of one logical fp32 GEMM  D[128 x N] (+)= A[128 x K] * W[N x K]^T :
// A hi / lo in TMEM (K columns each), W hi / lo in shared memory ([K/4][N][4] floats each).
// Called by ONE thread.  `first` = true overwrites D.
__device__ __forceinline__ void gemm_3xtf32(uint32_t d_tmem, uint32_t a_hi_tmem, uint32_t a_lo_tmem,
                                            const float *w_hi, const float *w_lo, int N, int K, bool first)
{
    const uint32_t idesc = idesc_tf32(128, N);
    const uint32_t lbo = (uint32_t)N * 16u, sbo = 128u;
    const uint32_t whi = smem_u32(w_hi), wlo = smem_u32(w_lo);
    for (int s = 0; s < K / 8; ++s) {
        const uint64_t b_hi = smem_desc_kmajor(whi + (uint32_t)s * 2u * lbo, lbo, sbo);
        const uint64_t b_lo = smem_desc_kmajor(wlo + (uint32_t)s * 2u * lbo, lbo, sbo);
        mma_tf32_ts(d_tmem, a_hi_tmem + 8 * s, b_lo, idesc, (first && s == 0) ? 0u : 1u);
        mma_tf32_ts(d_tmem, a_lo_tmem + 8 * s, b_hi, idesc, 1u);
        mma_tf32_ts(d_tmem, a_hi_tmem + 8 * s, b_hi, idesc, 1u);
    }
}

// Several INDEPENDENT 3xTF32 GEMMs (different accumulators) issued round-robin, one TF32 product of each in turn.
// A tcgen05.mma that accumulates into the columns the previous one wrote has to wait for it; with K = 8 per
// instruction and narrow N the instructions are far shorter than that dependency latency, so back-to-back MMAs of ONE
// chain run at the latency, not at the throughput, of the tensor pipe.  Interleaving independent chains hides it.
struct Gemm3x {
    uint32_t d_tmem, a_hi_tmem, a_lo_tmem;
    const float *w_hi, *w_lo;
    int N, K;
};
template <int kCount>
__device__ __forceinline__ void gemm_3xtf32_interleaved(const Gemm3x (&g)[kCount])
{
    int max_steps = 0;
#pragma unroll
    for (int i = 0; i < kCount; ++i) max_steps = g[i].K / 8 > max_steps ? g[i].K / 8 : max_steps;
    for (int s = 0; s < max_steps; ++s) {
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int i = 0; i < kCount; ++i) {
                if (s < g[i].K / 8) {
                    const uint32_t lbo = (uint32_t)g[i].N * 16u;
                    const uint32_t w = smem_u32(p == 0 ? g[i].w_lo : g[i].w_hi) + (uint32_t)s * 2u * lbo;
                    const uint32_t a = (p == 1 ? g[i].a_lo_tmem : g[i].a_hi_tmem) + 8 * s;
                    mma_tf32_ts(g[i].d_tmem, a, smem_desc_kmajor(w, lbo, 128u), idesc_tf32(128, g[i].N),
                                (s == 0 && p == 0) ? 0u : 1u);
                }
            }
        }
    }
}

}  // namespace umma
}  // namespace cgs

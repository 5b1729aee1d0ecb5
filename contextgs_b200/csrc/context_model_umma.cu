// One level of the anchor-level context model on the tcgen05 tensor cores.
//
// Same contract as context_level_kernel in context_model.cu (reference loop body
// scene/gaussian_model.py:1562-1652 + Entropy_gaussian :1666-1670 + the sums :1685-1693): gather the
// coarser-level context -> context MLP (71|15 -> 100 ReLU -> 175) -> adaptive steps -> quantise ->
// scatter -> discretised-Gaussian bits -> fp64 bit sums.  The two MLP layers run as 3xTF32
// tcgen05.mma with every activation resident in TENSOR MEMORY (see umma.cuh and
// neural_gaussians_umma.cu for the scheme):
//
//   persistent CTA per SM, 256 threads, tile = 128 level rows = 128 TMEM lanes;
//   thread (row = 32*(warp%4) + lane, half = warp/4): two threads share a row.
//   TMEM columns (432 of 512):
//     [  0,144)  input x_hi [0,72) | x_lo [72,144)            -> later hidden_lo [0,112)
//     [144,256)  layer-1 accumulator (100 units + 12 zero pads) -> hidden_hi in place
//     [256,432)  layer-2 accumulator: mu_f 50 | sigma_f 50 | mu_s 6 | sigma_s 6 | mu_o 30 | sigma_o 30 |
//                dQ_f dQ_s dQ_o | pad
//   shared memory: W1 hi/lo [K1p/4][112][4], W2 hi/lo [26][176][4], biases (212 KB for K1 = 71).
//   The next tile's gathered rows are prefetched into registers while the MMAs and epilogues run.
#include "entropy_math.cuh"
#include "umma.cuh"

namespace cgs {
namespace cmu {
constexpr int kRows = 128, kFront = 256, kBack = 512, kThreads = kFront + kBack;   // FRONT: 8 warps, BACK: 16 warps (see the kernel)
constexpr int kN1 = 112, kK2 = 104, kN2 = 176;
constexpr uint32_t kColXHi = 0, kColXLo = 72, kColHLo = 0, kColD1 = 144, kColD2 = 256, kTmemCols = 512;

template <int K1>
struct Layout {
    static constexpr int kK1p = (K1 + 7) / 8 * 8;                 // 72 | 16
    static constexpr int kW1 = (kK1p / 4) * kN1 * 4;              // floats per hi / lo part
    static constexpr int kW2 = (kK2 / 4) * kN2 * 4;
    static constexpr int kOffW1Hi = 0, kOffW1Lo = kW1, kOffW2Hi = 2 * kW1, kOffW2Lo = 2 * kW1 + kW2;
    static constexpr int kOffB1 = 2 * kW1 + 2 * kW2, kOffB2 = kOffB1 + kN1, kPacked = kOffB2 + kN2;
    static constexpr int kHalf0 = K1 == 71 ? 40 : 8;              // layer-1 inputs staged by half 0
    static constexpr int kXRegs = K1 == 71 ? 40 : 8;              // registers per thread for its share
};

template <int K1>
struct Smem {
    float w[Layout<K1>::kPacked];
    uint32_t tmem;
    int timeout;
    int sym_minmax[6];            // min / max coded symbol of the feat, scaling, offsets streams seen by this CTA
    alignas(8) uint64_t bar[3];   // layer-1 done | layer-2 done | layer-2 accumulator released by BACK
};

template <int kCount>
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kCount) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

struct Args {
    const float *packed_w;
    const int *orig_idx, *ctx_src;
    const float *level_anchor;
    int n_rows;
    const float *anchor, *hyper_q, *feat, *scaling, *offsets, *mask;
    const uint8_t *choose;
    const float *noise;
    float feat_mean, scaling_mean, offset_mean;
    float *feat_q, *scaling_q, *offsets_q, *bits_out;
    double *bit_sums;
    int32_t *err_flag;
    float *params_out;   // [n_rows][176]: mean[86] | scale[86] | Q_feat Q_scaling Q_offsets | 0 (codec / backward; optional)
    int32_t *symbol_minmax;   // optional [6]: min / max coded symbol of the feat, scaling, offsets streams of this level
    int predict_only;    // 1: only params_out is produced (decoder: the attributes are not known yet)
    float *save_h;       // [n_rows][112] hidden activations relu(D1 + b1) (training: consumed by the tcgen05 backward; optional)
    uint32_t *save_hmask;   // [n_rows][4] their sign bits: bit (8c + j) of word pair `half` <-> hidden unit 56 half + 8c + j
};

template <int K1>
struct RowInputs {
    float x[Layout<K1>::kXRegs];
    int o;   // original anchor index of the row (-1: padding)
};

// half 0 stages inputs k in [0, kHalf0), half 1 the rest (K1 = 71: [anchor 3 | feat_q 50 | scaling_q 6 |
// hyper_q 12]; K1 = 15: [level anchor 3 | hyper_q 12]).
template <int K1>
__device__ __forceinline__ void load_row(RowInputs<K1> &r, const Args &A, int row, int half)
{
    constexpr int XR = Layout<K1>::kXRegs;
#pragma unroll
    for (int j = 0; j < XR; ++j) r.x[j] = 0.f;
    r.o = -1;
    if (row >= A.n_rows) return;
    const int o = __ldg(A.orig_idx + row);
    r.o = o;
    const float *hq = A.hyper_q + (size_t)o * kHyper;
    if (K1 == 71) {
        const int s = __ldg(A.ctx_src + row);
        const float *fq = A.feat_q + (size_t)s * kCF;
        if (half == 0) {  // k 0..39 = anchor[s] (3) | feat_q[s][0..36]
#pragma unroll
            for (int j = 0; j < 3; ++j) r.x[j] = __ldg(A.anchor + 3 * (size_t)s + j);
#pragma unroll
            for (int j = 0; j < 37; ++j) r.x[3 + j] = fq[j];
        } else {          // k 40..70 = feat_q[s][37..49] | scaling_q[s] (6) | hyper_q[o] (12)
#pragma unroll
            for (int j = 0; j < 13; ++j) r.x[j] = fq[37 + j];
#pragma unroll
            for (int j = 0; j < 6; ++j) r.x[13 + j] = A.scaling_q[(size_t)s * kCS + j];
#pragma unroll
            for (int j = 0; j < 12; ++j) r.x[19 + j] = __ldg(hq + j);
        }
    } else {
        if (half == 0) {  // k 0..7 = level anchor (3) | hyper_q[0..4]
#pragma unroll
            for (int j = 0; j < 3; ++j) r.x[j] = __ldg(A.level_anchor + 3 * (size_t)row + j);
#pragma unroll
            for (int j = 0; j < 5; ++j) r.x[3 + j] = __ldg(hq + j);
        } else {          // k 8..14 = hyper_q[5..11]
#pragma unroll
            for (int j = 0; j < 7; ++j) r.x[j] = __ldg(hq + 5 + j);
        }
    }
}

struct ChunkDesc {
    int j0;              // index among the 86 coded values
    int cnt;             // values in the group
    uint32_t mu_col, sg_col;  // accumulator columns of mean / scale
    int grp, dim, k0;    // attribute (0 feat, 1 scaling, 2 offsets), its row width, first index inside it
};

// BACK: four threads share a row; the 86 coded values are split into quarters of three groups of <= 8 values each:
//   q0: feat 0..23;  q1: feat 24..47;  q2: feat 48..49 | scaling 0..5 | offsets 0..7;  q3: offsets 8..29
__device__ __forceinline__ ChunkDesc chunk_desc_q(int c, int q)
{
    ChunkDesc d;
    if (q < 2) {
        const int k0 = 24 * q + 8 * c;
        d.j0 = k0; d.cnt = 8; d.mu_col = k0; d.sg_col = kCF + k0; d.grp = 0; d.dim = kCF; d.k0 = k0;
    } else if (q == 2) {
        if (c == 0) {
            d.j0 = 48; d.cnt = 2; d.mu_col = 48; d.sg_col = kCF + 48; d.grp = 0; d.dim = kCF; d.k0 = 48;
        } else if (c == 1) {
            d.j0 = kCF; d.cnt = 6; d.mu_col = 100; d.sg_col = 106; d.grp = 1; d.dim = kCS; d.k0 = 0;
        } else {
            d.j0 = kCF + kCS; d.cnt = 8; d.mu_col = 112; d.sg_col = 142; d.grp = 2; d.dim = kCO; d.k0 = 0;
        }
    } else {
        const int k0 = 8 + 8 * c;
        d.j0 = kCF + kCS + k0; d.cnt = k0 + 8 <= 30 ? 8 : 30 - k0; d.mu_col = 112 + k0; d.sg_col = 142 + k0;
        d.grp = 2; d.dim = kCO; d.k0 = k0;
    }
    return d;
}

// kSym: also reduce the min / max coded symbol of each stream (bitstream encoder); kNoise: training (A.noise is given)
template <int K1, bool kSym, bool kNoise>
__global__ void __launch_bounds__(kThreads, 1) context_level_umma_kernel(Args A)
{
    using LY = Layout<K1>;
    using SM = Smem<K1>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SM &S = *reinterpret_cast<SM *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Two warp groups work on DIFFERENT tiles (as in neural_gaussians_umma.cu): FRONT (warps 0-7, two threads per row)
    // gathers and stages the context rows of tile t+1, runs layer 1 and the ReLU epilogue and issues layer 2 as soon as
    // BACK has released the layer-2 accumulator of tile t; BACK (warps 8-23, FOUR threads per row: the likelihood
    // epilogue is the long pole) quantises and scores tile t.  TMEM columns [0,256) belong to FRONT (free once the
    // layer-2 MMAs have read the hidden activations), [256,432) change hands through an mbarrier.
    const bool front = tid < kFront;
    const int uwarp = umma::uniform_warp();                   // warp index the compiler knows to be warp-uniform
    const int gtid = front ? tid : tid - kFront, gwarp = gtid >> 5;
    const int half = gwarp >> 2;      // FRONT: half of the row's inputs / hidden units; BACK: quarter of its coded values
    const int row = 32 * (warp & 3) + lane;
    const int num_tiles = (A.n_rows + kRows - 1) / kRows;
    const int stride = (int)gridDim.x;

    if (warp == 0) umma::tmem_alloc(&S.tmem, kTmemCols);
    if (tid == 0) {
        umma::mbar_init(&S.bar[0], 1);
        umma::mbar_init(&S.bar[1], 1);
        umma::mbar_init(&S.bar[2], 1);
        umma::fence_mbar_init();
        S.timeout = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) S.sym_minmax[i] = (i & 1) ? -0x7fffffff : 0x7fffffff;
    }
    {
        const float4 *s4 = reinterpret_cast<const float4 *>(A.packed_w);
        float4 *d4 = reinterpret_cast<float4 *>(S.w);
        for (int i = tid; i < LY::kPacked / 4; i += kThreads) d4[i] = __ldg(s4 + i);
    }
    umma::fence_proxy_async_smem();
    umma::fence_before_thread_sync();
    __syncthreads();
    umma::fence_after_thread_sync();
    const uint32_t tbase = S.tmem;
    const uint32_t tl = tbase + ((uint32_t)(32 * (warp & 3)) << 16);

    double tot_f = 0.0, tot_s = 0.0, tot_o = 0.0;   // fp64 across tiles: the sums are exact to ~1e-7
    float n_chosen = 0.f;
    const bool pred = A.predict_only != 0;

    if (front) {
        // =============================== FRONT: gather + stage, layer 1, ReLU epilogue, MMA issue ===============
        int tile = blockIdx.x;
        for (uint32_t it = 0; tile < num_tiles; ++it, tile += stride) {
            const uint32_t parity = it & 1u;
            // the gathered context rows of this tile travel while FRONT waits for its TMEM columns (FRONT has slack:
            // no register double buffer, which keeps the 768-thread CTA inside 85 registers per thread)
            RowInputs<K1> cur;
            load_row<K1>(cur, A, tile * kRows + row, half);
            if (it > 0) {   // the layer-2 MMAs of the previous tile have read the hidden activations: columns [0,256) are free
                if (!umma::mbar_wait(&S.bar[1], parity ^ 1u)) S.timeout = 1;
                umma::fence_after_thread_sync();
            }
            // ---- stage the layer-1 input --------------------------------------------------------------
            {
                const uint32_t k0 = half == 0 ? 0u : (uint32_t)LY::kHalf0;
                constexpr int kChunks0 = LY::kHalf0 / 8, kChunks1 = (LY::kK1p - LY::kHalf0) / 8;
#pragma unroll
                for (int c = 0; c < (kChunks0 > kChunks1 ? kChunks0 : kChunks1); ++c) {
                    if (c < (half == 0 ? kChunks0 : kChunks1)) {
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
            group_sync<kFront>(1);
            if (uwarp == 0) {      // warp-uniform branch + elect: back-to-back tcgen05.mma (see umma.cuh)
                if (umma::elect_one_sync()) {
                    umma::fence_after_thread_sync();
                    umma::gemm_3xtf32(tbase + kColD1, tbase + kColXHi, tbase + kColXLo, S.w + LY::kOffW1Hi, S.w + LY::kOffW1Lo,
                                      kN1, LY::kK1p, true);
                    umma::umma_commit(&S.bar[0]);
                }
                __syncwarp();
            }
            if (!umma::mbar_wait(&S.bar[0], parity)) S.timeout = 1;
            umma::fence_after_thread_sync();
            // ---- epilogue 1: hidden = relu(D1 + b1) -> hi in place, lo to region 0 (cols 56*half .. +56) ----
            const int grow_f = tile * kRows + row;
            float *save_row = (A.save_h && grow_f < A.n_rows) ? A.save_h + (size_t)grow_f * kN1 : nullptr;
            uint32_t hm0 = 0u, hm1 = 0u;
#pragma unroll 1
            for (int c = 0; c < 7; ++c) {
                const uint32_t col = (uint32_t)(56 * half + 8 * c);
                uint32_t v[8], hi[8], lo[8];
                float h[8];
                uint32_t bits = 0u;
                umma::tmem_ld8(tl + kColD1 + col, v);
                umma::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    h[j] = fmaxf(__uint_as_float(v[j]) + S.w[LY::kOffB1 + col + j], 0.f);
                    umma::split_tf32(h[j], hi[j], lo[j]);
                    bits |= h[j] > 0.f ? (1u << j) : 0u;
                }
                umma::tmem_st8(tl + kColD1 + col, hi);
                umma::tmem_st8(tl + kColHLo + col, lo);
                if (save_row) {
                    reinterpret_cast<float4 *>(save_row + col)[0] = make_float4(h[0], h[1], h[2], h[3]);
                    reinterpret_cast<float4 *>(save_row + col)[1] = make_float4(h[4], h[5], h[6], h[7]);
                    if (c < 4) hm0 |= bits << (8 * c);
                    else hm1 |= bits << (8 * (c - 4));
                }
            }
            if (save_row) {
                A.save_hmask[(size_t)grow_f * 4 + 2 * half] = hm0;
                A.save_hmask[(size_t)grow_f * 4 + 2 * half + 1] = hm1;
            }
            umma::tmem_wait_st();
            umma::fence_before_thread_sync();
            group_sync<kFront>(1);
            if (uwarp == 0 && umma::elect_one_sync()) {
                umma::fence_after_thread_sync();
                if (it > 0) {   // BACK has finished reading the layer-2 accumulator of the previous tile
                    if (!umma::mbar_wait(&S.bar[2], parity ^ 1u)) S.timeout = 1;
                    umma::fence_after_thread_sync();
                }
                umma::gemm_3xtf32(tbase + kColD2, tbase + kColD1, tbase + kColHLo, S.w + LY::kOffW2Hi, S.w + LY::kOffW2Lo,
                                  kN2, kK2, true);
                umma::umma_commit(&S.bar[1]);
            }
        }
    } else {
        // =============================== BACK: steps, quantise, scatter, likelihood ===============================
        int tile = blockIdx.x;
        auto orig_of = [&](int t) -> int {
            const int g = t * kRows + row;
            return (t < num_tiles && g < A.n_rows) ? __ldg(A.orig_idx + g) : -1;
        };
        int o_next = orig_of(tile);
        for (uint32_t it = 0; tile < num_tiles; ++it, tile += stride) {
            const uint32_t parity = it & 1u;
            const int grow = tile * kRows + row;
            const int o = o_next;
            o_next = orig_of(tile + stride);
            float sum_f = 0.f, sum_s = 0.f, sum_o = 0.f;
            const bool chosen = !pred && o >= 0 && (A.choose ? A.choose[o] != 0 : true);
            // offset masks of the row as bits (values are exactly 0 / 1: utils/entropy_models / gaussian_model.py:1670)
            uint32_t mkbits = 0x3ffu;
            if (o >= 0 && half >= 2 && (chosen || (kSym && !pred))) {   // kSym: the alphabets only cover coded offsets
                mkbits = 0;
#pragma unroll
                for (int k = 0; k < 10; ++k) mkbits |= __ldg(A.mask + o * 10 + k) != 0.f ? (1u << k) : 0u;
            }
            // The six groups run as a ROLLED loop: fully unrolled (with the attributes prefetched into 48 registers) the
            // kernel was 24.9 k SASS instructions and spent its time waiting for the instruction cache ("no instruction"
            // was the top stall reason, profiles/r01_ctx2_*).  The attributes of group c+1 are fetched while group c is
            // evaluated; those of group 0 while the layer-2 MMAs run.
            auto fetch_x = [&](int c, float (&x8)[8]) {
                const ChunkDesc cd = chunk_desc_q(c, half);
                const float *src = (cd.grp == 0 ? A.feat : (cd.grp == 1 ? A.scaling : A.offsets)) + (size_t)(o < 0 ? 0 : o) * cd.dim + cd.k0;
                // every group starts on an even index of an 8-byte aligned row and holds an even number of values
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float2 v = (!pred && o >= 0 && j < cd.cnt) ? __ldg(reinterpret_cast<const float2 *>(src + j))
                                                                     : make_float2(0.f, 0.f);
                    x8[j] = v.x; x8[j + 1] = v.y;
                }
            };
            float xn[8];
            fetch_x(0, xn);
            if (!umma::mbar_wait(&S.bar[1], parity)) S.timeout = 1;
            umma::fence_after_thread_sync();

            // ---- epilogue 2: steps, quantise, scatter, bits ----------------------------------------------
            float Qf, Qs, Qo;
            {
                uint32_t v[4];
                umma::tmem_ld4(tl + kColD2 + 172, v);
                umma::tmem_wait_ld();
                Qf = fmaxf(kQf0 * (1.0f + tanhf(__uint_as_float(v[0]) + S.w[LY::kOffB2 + 172])), 1e-9f);
                Qs = fmaxf(kQs0 * (1.0f + tanhf(__uint_as_float(v[1]) + S.w[LY::kOffB2 + 173])), 1e-9f);
                Qo = fmaxf(kQo0 * (1.0f + tanhf(__uint_as_float(v[2]) + S.w[LY::kOffB2 + 174])), 1e-9f);
            }
            if (half == 0 && chosen) n_chosen += 1.f;
            float *prow = (A.params_out && o >= 0) ? A.params_out + (size_t)grow * kLdG2 : nullptr;
            if (prow && half == 0) *reinterpret_cast<float4 *>(prow + 172) = make_float4(Qf, Qs, Qo, 0.f);
            const float *nz = kNoise ? A.noise + (size_t)grow * kCE : nullptr;
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {
                const ChunkDesc cd = chunk_desc_q(c, half);
                uint32_t vm[8], vs[8];
                float xc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) xc[j] = xn[j];
                umma::tmem_ld8(tl + kColD2 + cd.mu_col, vm);
                umma::tmem_ld8(tl + kColD2 + cd.sg_col, vs);
                float nzv[kNoise ? 8 : 1];   // training noise of the chunk: in flight while the TMEM loads and the stores complete
                if (kNoise) {
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const float2 v = (o >= 0 && j < cd.cnt) ? __ldg(reinterpret_cast<const float2 *>(nz + cd.j0 + j))
                                                                : make_float2(0.f, 0.f);   // (rows past the level have no noise)
                        nzv[kNoise ? j : 0] = v.x; nzv[kNoise ? j + 1 : 0] = v.y;
                    }
                }
                if (c + 1 < 3) fetch_x(c + 1, xn);
                umma::tmem_wait_ld();
                // alphabet bounds of the level's three streams (bitstream codec): the symbols exist here anyway
                float sym_lo = 3.0e38f, sym_hi = -3.0e38f;
                if (o >= 0) {
                const float Q = cd.grp == 0 ? Qf : (cd.grp == 1 ? Qs : Qo);
                const float x_mean = cd.grp == 0 ? A.feat_mean : (cd.grp == 1 ? A.scaling_mean : A.offset_mean);
                float *dst = (cd.grp == 0 ? A.feat_q : (cd.grp == 1 ? A.scaling_q : A.offsets_q)) + o * cd.dim + cd.k0;
                float acc = 0.f, xq_even = 0.f;
                if (prow && (chosen || !A.save_h)) {   // training: the backward reads (mean, scale) of the chosen rows only
                    // (mean, scale) of the group as 8-byte stores (every group starts on an even index and holds an even
                    // number of values; scalar stores cost one 32-byte sector transaction per value and doubled the
                    // kernel's time when the training path started to save them)
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        if (j < cd.cnt) {
                            *reinterpret_cast<float2 *>(prow + cd.j0 + j) =
                                make_float2(__uint_as_float(vm[j]) + S.w[LY::kOffB2 + cd.mu_col + j],
                                            __uint_as_float(vm[j + 1]) + S.w[LY::kOffB2 + cd.mu_col + j + 1]);
                            *reinterpret_cast<float2 *>(prow + kCE + cd.j0 + j) =
                                make_float2(__uint_as_float(vs[j]) + S.w[LY::kOffB2 + cd.sg_col + j],
                                            __uint_as_float(vs[j + 1]) + S.w[LY::kOffB2 + cd.sg_col + j + 1]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < cd.cnt) {
                        const float mean = __uint_as_float(vm[j]) + S.w[LY::kOffB2 + cd.mu_col + j];
                        const float scale = __uint_as_float(vs[j]) + S.w[LY::kOffB2 + cd.sg_col + j];
                        if (pred) continue;
                        const float x = xc[j];
                        float sym;
                        const float xq = kNoise ? x + nzv[kNoise ? j : 0] * Q : ste_round_sym(x, Q, sym);
                        if (j & 1) *reinterpret_cast<float2 *>(dst + j - 1) = make_float2(xq_even, xq);   // 8-byte stores
                        else xq_even = xq;
                        if (kSym && !kNoise && (cd.grp != 2 || ((mkbits >> ((cd.k0 + j) / 3)) & 1u))) {
                            sym_lo = fminf(sym_lo, sym);
                            sym_hi = fmaxf(sym_hi, sym);
                        }
                        float bits = 0.f;
                        if (chosen) {
                            bits = gaussian_bits_one(xq, mean, scale, Q, x_mean);
                            if (cd.grp == 2 && !((mkbits >> ((cd.k0 + j) / 3)) & 1u)) bits = 0.f;
                            acc += bits;
                        }
                        if (A.bits_out) A.bits_out[(size_t)o * kCE + cd.j0 + j] = bits;
                    }
                }
                if (cd.grp == 0) sum_f += acc;
                else if (cd.grp == 1) sum_s += acc;
                else sum_o += acc;
                }
                if (kSym) {   // whole warp, same group: one shared atomic pair per warp and chunk
                    const int lo = __reduce_min_sync(0xffffffffu, (int)fminf(sym_lo, 2.0e9f));
                    const int hi = __reduce_max_sync(0xffffffffu, (int)fmaxf(sym_hi, -2.0e9f));
                    if (lane == 0 && lo <= hi) {
                        atomicMin(&S.sym_minmax[2 * cd.grp], lo);
                        atomicMax(&S.sym_minmax[2 * cd.grp + 1], hi);
                    }
                }
            }
            tot_f += (double)sum_f; tot_s += (double)sum_s; tot_o += (double)sum_o;
            // all TMEM reads of this tile are complete: FRONT may issue the next tile's layer-2 MMAs
            umma::fence_before_thread_sync();
            group_sync<kBack>(2);
            if (gtid == 0) mbar_arrive(&S.bar[2]);
        }
    }

    // ---- per-CTA reduction of the bit sums -> one fp64 atomic per sum ----------------------------------
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        tot_f += __shfl_xor_sync(0xffffffffu, tot_f, off);
        tot_s += __shfl_xor_sync(0xffffffffu, tot_s, off);
        tot_o += __shfl_xor_sync(0xffffffffu, tot_o, off);
        n_chosen += __shfl_xor_sync(0xffffffffu, n_chosen, off);
    }
    __shared__ double s_red[4][kThreads / 32];
    if (lane == 0) {
        s_red[0][warp] = tot_f; s_red[1][warp] = tot_s; s_red[2][warp] = tot_o; s_red[3][warp] = (double)n_chosen;
    }
    umma::fence_before_thread_sync();
    __syncthreads();
    if (tid < 4) {
        double v = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) v += s_red[tid][w];
        if (v != 0.0) atomicAdd(A.bit_sums + tid, v);
    }
    if (kSym && tid < 6 && S.sym_minmax[tid & ~1] <= S.sym_minmax[tid | 1]) {   // published by the barrier above
        if (tid & 1) atomicMax(A.symbol_minmax + tid, S.sym_minmax[tid]);
        else atomicMin(A.symbol_minmax + tid, S.sym_minmax[tid]);
    }
    if (tid == 0 && S.timeout) atomicExch(A.err_flag, 1);
    if (warp == 0) umma::tmem_dealloc(tbase, kTmemCols);
}

__global__ void symbol_minmax_init_kernel(int32_t *minmax)
{
    if (threadIdx.x < 6) minmax[threadIdx.x] = (threadIdx.x & 1) ? -2139062144 : 2139062143;   // empty stream: max < min
}

template <int K1>
static int launch(const Args &a, cudaStream_t st)
{
    using SM = Smem<K1>;
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(context_level_umma_kernel<K1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
        cudaFuncSetAttribute(context_level_umma_kernel<K1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
        cudaFuncSetAttribute(context_level_umma_kernel<K1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM));
        if (sm_count <= 0) sm_count = kNumSMs;
    }
    // the per-anchor attributes are read and written two floats at a time
    for (const void *q : {(const void *)a.feat, (const void *)a.scaling, (const void *)a.offsets, (const void *)a.feat_q,
                          (const void *)a.scaling_q, (const void *)a.offsets_q}) {
        if (reinterpret_cast<uintptr_t>(q) & 7u) {
            set_error("cgs_context_level_umma_forward: the attribute arrays must be 8-byte aligned");
            return -2;
        }
    }
    const int tiles = (a.n_rows + kRows - 1) / kRows;
    const int grid = tiles < sm_count ? tiles : sm_count;
    StageScope sc(ST_CTX_LEVEL, st, 1);
    if (a.noise) context_level_umma_kernel<K1, false, true><<<grid, kThreads, sizeof(SM), st>>>(a);   // (training has no alphabets)
    else if (a.symbol_minmax) context_level_umma_kernel<K1, true, false><<<grid, kThreads, sizeof(SM), st>>>(a);
    else context_level_umma_kernel<K1, false, false><<<grid, kThreads, sizeof(SM), st>>>(a);
    return check_launch("cgs_context_level_umma_forward");
}
}  // namespace cmu
}  // namespace cgs

using namespace cgs;

extern "C" int cgs_context_level_umma_packed_floats(int in_dim)
{
    if (in_dim == 71) return cmu::Layout<71>::kPacked;
    if (in_dim == 15) return cmu::Layout<15>::kPacked;
    return -1;
}

extern "C" int cgs_context_level_umma_forward(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                              const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                              const float *anchor, const float *hyper_q, const float *feat,
                                              const float *scaling, const float *offsets, const float *mask,
                                              const uint8_t *choose, const float *noise, float feat_mean,
                                              float scaling_mean, float offset_mean, float *feat_q, float *scaling_q,
                                              float *offsets_q, float *bits_out, double *bit_sums, int32_t *err_flag,
                                              void *stream)
{
    return cgs_context_level_umma_forward_ex(in_dim, packed_w, orig_idx, ctx_src, level_anchor, n_rows, anchor, hyper_q, feat,
                                             scaling, offsets, mask, choose, noise, feat_mean, scaling_mean, offset_mean,
                                             feat_q, scaling_q, offsets_q, bits_out, bit_sums, err_flag, nullptr, 0, nullptr, stream);
}

extern "C" int cgs_context_level_umma_forward_train(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                                    const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                                    const float *anchor, const float *hyper_q, const float *feat,
                                                    const float *scaling, const float *offsets, const float *mask,
                                                    const uint8_t *choose, const float *noise, float feat_mean,
                                                    float scaling_mean, float offset_mean, float *feat_q, float *scaling_q,
                                                    float *offsets_q, float *bits_out, double *bit_sums, int32_t *err_flag,
                                                    float *params_out, float *save_h, uint32_t *save_hmask, void *stream)
{
    if (n_rows <= 0) return 0;
    CGS_CHECK_PTR(packed_w); CGS_CHECK_PTR(orig_idx); CGS_CHECK_PTR(anchor); CGS_CHECK_PTR(hyper_q);
    CGS_CHECK_PTR(feat_q); CGS_CHECK_PTR(scaling_q); CGS_CHECK_PTR(bit_sums); CGS_CHECK_PTR(err_flag);
    CGS_CHECK_PTR(feat); CGS_CHECK_PTR(scaling); CGS_CHECK_PTR(offsets); CGS_CHECK_PTR(mask); CGS_CHECK_PTR(offsets_q);
    CGS_CHECK_PTR(params_out); CGS_CHECK_PTR(save_h); CGS_CHECK_PTR(save_hmask);
    cmu::Args a;
    a.params_out = params_out; a.predict_only = 0; a.save_h = save_h; a.save_hmask = save_hmask; a.symbol_minmax = nullptr;
    a.packed_w = packed_w; a.orig_idx = orig_idx; a.ctx_src = ctx_src; a.level_anchor = level_anchor; a.n_rows = n_rows;
    a.anchor = anchor; a.hyper_q = hyper_q; a.feat = feat; a.scaling = scaling; a.offsets = offsets; a.mask = mask;
    a.choose = choose; a.noise = noise; a.feat_mean = feat_mean; a.scaling_mean = scaling_mean;
    a.offset_mean = offset_mean; a.feat_q = feat_q; a.scaling_q = scaling_q; a.offsets_q = offsets_q;
    a.bits_out = bits_out; a.bit_sums = bit_sums; a.err_flag = err_flag;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_dim == 71) {
        CGS_CHECK_PTR(ctx_src);
        return cmu::launch<71>(a, st);
    }
    if (in_dim == 15) {
        CGS_CHECK_PTR(level_anchor);
        return cmu::launch<15>(a, st);
    }
    set_error("%s: unsupported context-MLP input width %d", __func__, in_dim);
    return -2;
}

extern "C" int cgs_context_level_umma_forward_ex(int in_dim, const float *packed_w, const int32_t *orig_idx,
                                                 const int32_t *ctx_src, const float *level_anchor, int n_rows,
                                                 const float *anchor, const float *hyper_q, const float *feat,
                                                 const float *scaling, const float *offsets, const float *mask,
                                                 const uint8_t *choose, const float *noise, float feat_mean,
                                                 float scaling_mean, float offset_mean, float *feat_q, float *scaling_q,
                                                 float *offsets_q, float *bits_out, double *bit_sums, int32_t *err_flag,
                                                 float *params_out, int predict_only, int32_t *symbol_minmax, void *stream)
{
    if (symbol_minmax) cmu::symbol_minmax_init_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(symbol_minmax);
    if (n_rows <= 0) return symbol_minmax ? check_launch(__func__) : 0;
    CGS_CHECK_PTR(packed_w); CGS_CHECK_PTR(orig_idx); CGS_CHECK_PTR(anchor); CGS_CHECK_PTR(hyper_q);
    CGS_CHECK_PTR(feat_q); CGS_CHECK_PTR(scaling_q); CGS_CHECK_PTR(bit_sums); CGS_CHECK_PTR(err_flag);
    if (predict_only) {
        CGS_CHECK_PTR(params_out);
    } else {
        CGS_CHECK_PTR(feat); CGS_CHECK_PTR(scaling); CGS_CHECK_PTR(offsets); CGS_CHECK_PTR(mask); CGS_CHECK_PTR(offsets_q);
    }
    cmu::Args a;
    a.params_out = params_out; a.predict_only = predict_only; a.save_h = nullptr; a.save_hmask = nullptr;
    a.symbol_minmax = predict_only ? nullptr : symbol_minmax;
    a.packed_w = packed_w; a.orig_idx = orig_idx; a.ctx_src = ctx_src; a.level_anchor = level_anchor; a.n_rows = n_rows;
    a.anchor = anchor; a.hyper_q = hyper_q; a.feat = feat; a.scaling = scaling; a.offsets = offsets; a.mask = mask;
    a.choose = choose; a.noise = noise; a.feat_mean = feat_mean; a.scaling_mean = scaling_mean;
    a.offset_mean = offset_mean; a.feat_q = feat_q; a.scaling_q = scaling_q; a.offsets_q = offsets_q;
    a.bits_out = bits_out; a.bit_sums = bit_sums; a.err_flag = err_flag;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_dim == 71) {
        CGS_CHECK_PTR(ctx_src);
        return cmu::launch<71>(a, st);
    }
    if (in_dim == 15) {
        CGS_CHECK_PTR(level_anchor);
        return cmu::launch<15>(a, st);
    }
    set_error("%s: unsupported context-MLP input width %d", __func__, in_dim);
    return -2;
}

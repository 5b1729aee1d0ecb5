// Backward of the fused anchor -> neural-Gaussian generation.
//
// Replaces what autograd does for gaussian_renderer/__init__.py:106-145 in the reference: the
// gradients of the emitted Gaussians (xyz, color, opacity, scaling, rot -- in emission order) are
// pulled back through the post-processing, the boolean compaction, the three decoder MLPs
// (scene/gaussian_model.py:153-174) and the view-direction computation, to
//   d anchor[N,3], d feat[N,50], d offsets[N,10,3], d scaling[N,6], d mask[N,10]
// (rows of non-visible anchors are left untouched: the caller zero-fills) and to the MLP weights.
//
// One persistent CTA per SM, 64 anchors per tile.  The forward activations are RECOMPUTED per tile
// (cheaper than a 440 B/anchor round trip through HBM); per-tile buffers are reused in place:
//   out  : pre-activations  -> their gradients          h : hidden -> hidden gradient (ReLU masked)
//   x    : layer-1 input    -> its gradient
// Weight gradients are accumulated in REGISTERS across all tiles of the CTA (each thread owns a
// fixed 5x5 block of W2 and a 7x5 block of W1) and added to HBM once per CTA at the end.
// fp32 FMA tiles; the tcgen05 variant of these GEMMs is the next step for this kernel.
#include "mlp_tile.cuh"

namespace cgs {
namespace ngb {
constexpr int kFeat = 50, kK = 10, kIn = 54, kHid = 150, kOut = 110, kPairs = kTM * kK;
// forward-layout block (same as neural_gaussians.cu)
constexpr int kLd1 = 152, kLdO = 12, kLdC = 32, kLdV = 72;
constexpr int kOffW1 = 0, kOffB1 = kOffW1 + kIn * kLd1, kOffW2o = kOffB1 + kLd1, kOffB2o = kOffW2o + kFeat * kLdO;
constexpr int kOffW2c = kOffB2o + kLdO, kOffB2c = kOffW2c + kFeat * kLdC, kOffW2v = kOffB2c + kLdC;
constexpr int kOffB2v = kOffW2v + kFeat * kLdV, kFwdFloats = kOffB2v + kLdV;  // 14292
// transposed block for the backward GEMMs: W1T[150][56] (hid-major), W2T per head [n_out][52] (out-major)
constexpr int kLdT1 = 56, kLdT2 = 52;
constexpr int kOffW1T = 0, kOffW2oT = kOffW1T + kHid * kLdT1, kOffW2cT = kOffW2oT + 10 * kLdT2;
constexpr int kOffW2vT = kOffW2cT + 30 * kLdT2, kBwdFloats = kOffW2vT + 70 * kLdT2;  // 14120

struct Smem {
    float wf[kFwdFloats];
    float wt[kBwdFloats];
    float x[56 * kTMp];          // 54 inputs (+2 pad rows used by the in-place d_x GEMM epilogue)
    float h[152 * kTMp];
    float out[112 * kTMp];
    float d_anchor[kTM * 3];
    float d_sc[kTM * 6];
    float anchor[kTM * 3];
    float scaling[kTM * 6];
    float dist[kTM];
    int src[kTM];
    uint32_t warp_cnt[kMlpThreads / 32];
    uint32_t tile_base;
    int tile;
};

// index of output unit n (0..109: opacity 10 | color 30 | cov 70) inside the forward-layout block, for
// hidden unit hh of the owning head
__device__ __forceinline__ int w2_index(int hh, int n)
{
    if (n < 10) return kOffW2o + hh * kLdO + n;
    if (n < 40) return kOffW2c + hh * kLdC + (n - 10);
    return kOffW2v + hh * kLdV + (n - 40);
}
__device__ __forceinline__ int b2_index(int n)
{
    if (n < 10) return kOffB2o + n;
    if (n < 40) return kOffB2c + (n - 10);
    return kOffB2v + (n - 40);
}
__device__ __forceinline__ int head_of(int n) { return n < 10 ? 0 : (n < 40 ? 1 : 2); }
}  // namespace ngb

__global__ void __launch_bounds__(kMlpThreads, 1)
neural_gaussians_backward_kernel(const float *__restrict__ w_fwd, const float *__restrict__ w_bwd,
                                 const int *__restrict__ vis_idx, int Nv, const float *__restrict__ anchor,
                                 const float *__restrict__ feat, const float *__restrict__ offsets,
                                 const float *__restrict__ scaling, const float *__restrict__ mask, float cx, float cy,
                                 float cz, const uint8_t *__restrict__ keep_mask, const float *__restrict__ g_xyz,
                                 const float *__restrict__ g_color, const float *__restrict__ g_opacity,
                                 const float *__restrict__ g_scaling, const float *__restrict__ g_rot,
                                 float *__restrict__ d_anchor, float *__restrict__ d_feat, float *__restrict__ d_offsets,
                                 float *__restrict__ d_scaling, float *__restrict__ d_mask, float *__restrict__ d_w,
                                 unsigned long long *scan_state, uint32_t *ticket)
{
    using namespace ngb;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int num_tiles = (Nv + kTM - 1) / kTM;

    copy_to_smem(S.wf, w_fwd, kFwdFloats);
    copy_to_smem(S.wt, w_bwd, kBwdFloats);
    for (int i = tid; i < 2 * kTMp; i += kMlpThreads) S.x[54 * kTMp + i] = 0.f;
    if (tid == 0) S.tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();

    // register-resident weight-gradient blocks
    //   W2: thread -> (n-group of 5 outputs, h-group of 5 hidden units): 22 x 10 = 220 threads
    //   W1: thread -> (hid-group of 5, in-group of 7): 30 x 8 = 240 threads
    float gw2[5][5], gw1[7][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) gw2[i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) gw1[i][j] = 0.f;
    float gb2 = 0.f, gb1 = 0.f;  // thread t < 110: bias of output t; thread t < 150: bias of hidden t
    const int w2_ng = tid % 22, w2_hg = tid / 22;          // valid when tid < 220
    const int w1_hg = tid % 30, w1_ig = tid / 30;          // valid when tid < 240

    for (int tile = S.tile; tile < num_tiles; tile = S.tile) {
        const int row0 = tile * kTM;
        // ---- stage inputs (as the forward) ---------------------------------------------------------
        {
            const int r = tid >> 2, q = tid & 3;
            const int row = row0 + r;
            int a = -1;
            if (row < Nv) a = vis_idx ? vis_idx[row] : row;
            if (q == 0) S.src[r] = a;
            if (a >= 0) {
                const float *f = feat + (size_t)a * kFeat;
                for (int k = q; k < kFeat; k += 4) S.x[k * kTMp + r] = f[k];
                if (q == 0) {
                    const float ax = anchor[3 * (size_t)a], ay = anchor[3 * (size_t)a + 1], az = anchor[3 * (size_t)a + 2];
                    const float vx = ax - cx, vy = ay - cy, vz = az - cz;
                    const float d = sqrtf(vx * vx + vy * vy + vz * vz);
                    S.x[(kFeat + 0) * kTMp + r] = vx / d;
                    S.x[(kFeat + 1) * kTMp + r] = vy / d;
                    S.x[(kFeat + 2) * kTMp + r] = vz / d;
                    S.x[(kFeat + 3) * kTMp + r] = d;
                    S.dist[r] = d;
                    S.anchor[3 * r] = ax; S.anchor[3 * r + 1] = ay; S.anchor[3 * r + 2] = az;
                }
                if (q == 1)
                    for (int k = 0; k < 6; ++k) S.scaling[6 * r + k] = scaling[(size_t)a * 6 + k];
            } else {
                for (int k = q; k < kIn; k += 4) S.x[k * kTMp + r] = 0.f;
                if (q == 0) S.dist[r] = 1.f;
            }
            if (tid < kTM * 3) S.d_anchor[tid] = 0.f;
            for (int i = tid; i < kTM * 6; i += kMlpThreads) S.d_sc[i] = 0.f;
        }
        __syncthreads();

        // ---- recompute the forward activations ---------------------------------------------------
        tile_gemm<5, ACT_RELU>(S.x, kIn, S.wf + kOffW1, kLd1, S.wf + kOffB1, kHid, S.h);
        __syncthreads();
        tile_gemm<1, ACT_NONE>(S.h, kFeat, S.wf + kOffW2o, kLdO, S.wf + kOffB2o, kK, S.out);
        tile_gemm<1, ACT_NONE>(S.h + kFeat * kTMp, kFeat, S.wf + kOffW2c, kLdC, S.wf + kOffB2c, 3 * kK,
                               S.out + kK * kTMp);
        tile_gemm<3, ACT_NONE>(S.h + 2 * kFeat * kTMp, kFeat, S.wf + kOffW2v, kLdV, S.wf + kOffB2v, 7 * kK,
                               S.out + 4 * kK * kTMp);

        // ---- ranks of the kept pairs in emission order (saved selection mask of the forward) ------
        bool keep[3];
        uint32_t seg_rank[3], seg_total[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int p = tid + i * kMlpThreads;
            keep[i] = false;
            if (p < kPairs) {
                const int r = p / kK;
                if (S.src[r] >= 0) keep[i] = keep_mask[(size_t)(row0 + r) * kK + (p - r * kK)] != 0;
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const uint32_t b = __ballot_sync(0xffffffffu, keep[i]);
            const uint32_t within = __popc(b & ((1u << lane) - 1));
            if (lane == 0) S.warp_cnt[warp] = __popc(b);
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kMlpThreads / 32; ++w) {
                const uint32_t c = S.warp_cnt[w];
                before += w < warp ? c : 0u;
                total += c;
            }
            seg_rank[i] = before + within;
            seg_total[i] = total;
            __syncthreads();
        }
        const uint32_t tile_total = seg_total[0] + seg_total[1] + seg_total[2];
        if (warp == 0) {
            const uint64_t excl = lookback_exclusive(scan_state, tile, tile_total);
            if (lane == 0) S.tile_base = (uint32_t)excl;
        }
        __syncthreads();  // also: all pre-activations are in S.out
        const uint32_t base = S.tile_base;

        // ---- per pair: gradient of the post-processing, written over the pre-activations -----------
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int p = tid + i * kMlpThreads;
            if (p >= kPairs) continue;
            const int r = p / kK, k = p - r * kK;
            float *po = S.out + k * kTMp + r;                      // opacity pre-activation
            float *pc = S.out + (kK + 3 * k) * kTMp + r;           // 3 colour pre-activations
            float *pv = S.out + (4 * kK + 7 * k) * kTMp + r;       // 7 covariance pre-activations
            if (!keep[i]) {
                po[0] = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) pc[c * kTMp] = 0.f;
#pragma unroll
                for (int c = 0; c < 7; ++c) pv[c * kTMp] = 0.f;
                continue;
            }
            const int a = S.src[r];
            uint32_t pos = base + seg_rank[i];
            if (i >= 1) pos += seg_total[0];
            if (i >= 2) pos += seg_total[1];
            const size_t P3 = 3 * (size_t)pos;
            const float gx = g_xyz[P3], gy = g_xyz[P3 + 1], gz = g_xyz[P3 + 2];
            const float *of = offsets + ((size_t)a * kK + k) * 3;
            const float *sc = S.scaling + 6 * r;
            float *dof = d_offsets + ((size_t)a * kK + k) * 3;
            dof[0] = gx * sc[0]; dof[1] = gy * sc[1]; dof[2] = gz * sc[2];
            atomicAdd(&S.d_anchor[3 * r + 0], gx);
            atomicAdd(&S.d_anchor[3 * r + 1], gy);
            atomicAdd(&S.d_anchor[3 * r + 2], gz);
            atomicAdd(&S.d_sc[6 * r + 0], gx * of[0]);
            atomicAdd(&S.d_sc[6 * r + 1], gy * of[1]);
            atomicAdd(&S.d_sc[6 * r + 2], gz * of[2]);
            // opacity = tanh(pre) * mask
            const float t = tanhf(po[0]);
            const float go = g_opacity[pos];
            const float mk = mask[(size_t)a * kK + k];
            d_mask[(size_t)a * kK + k] = go * t;
            po[0] = go * mk * (1.0f - t * t);
            // colour = sigmoid(pre)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float s = 1.0f / (1.0f + expf(-pc[c * kTMp]));
                pc[c * kTMp] = g_color[P3 + c] * s * (1.0f - s);
            }
            // scaling = sc[3:6] * sigmoid(pre[0:3]) ; rot = normalize(pre[3:7])
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float s = 1.0f / (1.0f + expf(-pv[c * kTMp]));
                const float g = g_scaling[P3 + c];
                atomicAdd(&S.d_sc[6 * r + 3 + c], g * s);
                pv[c * kTMp] = g * sc[3 + c] * s * (1.0f - s);
            }
            const float q0 = pv[3 * kTMp], q1 = pv[4 * kTMp], q2 = pv[5 * kTMp], q3 = pv[6 * kTMp];
            const float nrm = fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);
            const float4 gr = reinterpret_cast<const float4 *>(g_rot)[pos];
            const float r0 = q0 / nrm, r1 = q1 / nrm, r2 = q2 / nrm, r3 = q3 / nrm;
            const float dot = r0 * gr.x + r1 * gr.y + r2 * gr.z + r3 * gr.w;
            pv[3 * kTMp] = (gr.x - r0 * dot) / nrm;
            pv[4 * kTMp] = (gr.y - r1 * dot) / nrm;
            pv[5 * kTMp] = (gr.z - r2 * dot) / nrm;
            pv[6 * kTMp] = (gr.w - r3 * dot) / nrm;
        }
        __syncthreads();

        // ---- dW2 += h (x) d_out, db2 -----------------------------------------------------------------
        if (tid < 220) {
            const int n0 = 5 * w2_ng, head = head_of(n0);  // groups of 5 never straddle a head (10, 30, 70)
            outer_accumulate<5, 5>(S.out, n0, kOut, S.h + head * kFeat * kTMp, 5 * w2_hg, kFeat, gw2);
        }
        if (tid < kOut) {
            float s = 0.f;
            for (int r = 0; r < kTM; ++r) s += S.out[tid * kTMp + r];
            gb2 += s;
        }
        __syncthreads();
        // ---- d_h = (W2^T d_out) * relu'(h), in place over h --------------------------------------------
        tile_gemm_relu_mask<2>(S.out, kK, S.wt + kOffW2oT, kLdT2, kFeat, S.h);
        tile_gemm_relu_mask<2>(S.out + kK * kTMp, 3 * kK, S.wt + kOffW2cT, kLdT2, kFeat, S.h + kFeat * kTMp);
        tile_gemm_relu_mask<2>(S.out + 4 * kK * kTMp, 7 * kK, S.wt + kOffW2vT, kLdT2, kFeat, S.h + 2 * kFeat * kTMp);
        __syncthreads();
        // ---- dW1 += x (x) d_h, db1 ----------------------------------------------------------------------
        if (tid < 240) outer_accumulate<7, 5>(S.x, 7 * w1_ig, kIn, S.h, 5 * w1_hg, kHid, gw1);
        if (tid < kHid) {
            float s = 0.f;
            for (int r = 0; r < kTM; ++r) s += S.h[tid * kTMp + r];
            gb1 += s;
        }
        __syncthreads();
        // ---- d_x = W1^T d_h, in place over x --------------------------------------------------------------
        tile_gemm<2, ACT_NONE>(S.h, kHid, S.wt + kOffW1T, kLdT1, S.x + 54 * kTMp /* zero bias rows */, kIn, S.x);
        __syncthreads();
        // ---- scatter the per-anchor gradients ---------------------------------------------------------------
        {
            const int r = tid >> 2, q = tid & 3;
            const int a = S.src[r];
            if (a >= 0) {
                float *df = d_feat + (size_t)a * kFeat;
                for (int k = q; k < kFeat; k += 4) df[k] = S.x[k * kTMp + r];
                if (q == 0) {
                    // view = u / |u|, dist = |u|, u = anchor - cam
                    const float d = S.dist[r];
                    const float vx = (S.anchor[3 * r] - cx) / d, vy = (S.anchor[3 * r + 1] - cy) / d,
                                vz = (S.anchor[3 * r + 2] - cz) / d;
                    const float gvx = S.x[(kFeat + 0) * kTMp + r], gvy = S.x[(kFeat + 1) * kTMp + r],
                                gvz = S.x[(kFeat + 2) * kTMp + r], gd = S.x[(kFeat + 3) * kTMp + r];
                    const float dot = vx * gvx + vy * gvy + vz * gvz;
                    d_anchor[3 * (size_t)a + 0] = S.d_anchor[3 * r + 0] + (gvx - vx * dot) / d + gd * vx;
                    d_anchor[3 * (size_t)a + 1] = S.d_anchor[3 * r + 1] + (gvy - vy * dot) / d + gd * vy;
                    d_anchor[3 * (size_t)a + 2] = S.d_anchor[3 * r + 2] + (gvz - vz * dot) / d + gd * vz;
                }
                if (q == 1)
                    for (int k = 0; k < 6; ++k) d_scaling[(size_t)a * 6 + k] = S.d_sc[6 * r + k];
            }
        }
        if (tid == 0) S.tile = (int)atomicAdd(ticket, 1u);
        __syncthreads();
    }

    // ---- one atomic per weight per CTA ------------------------------------------------------------------
    if (tid < 220) {
        const int n0 = 5 * w2_ng, head = head_of(n0);
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int n = n0 + i, hh = 5 * w2_hg + j;
                if (n < kOut && hh < kFeat) atomicAdd(d_w + w2_index(hh, n), gw2[i][j]);
            }
        (void)head;
    }
    if (tid < 240) {
#pragma unroll
        for (int i = 0; i < 7; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int k = 7 * w1_ig + i, hh = 5 * w1_hg + j;
                if (k < kIn && hh < kHid) atomicAdd(d_w + kOffW1 + k * kLd1 + hh, gw1[i][j]);
            }
    }
    if (tid < kOut) atomicAdd(d_w + b2_index(tid), gb2);
    if (tid < kHid) atomicAdd(d_w + kOffB1 + tid, gb1);
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_neural_gaussians_backward_packed_floats(void) { return ngb::kBwdFloats; }

extern "C" size_t cgs_neural_gaussians_backward_workspace_bytes(int Nv)
{
    const size_t tiles = (size_t)(Nv > 0 ? (Nv + kTM - 1) / kTM : 1);
    return align_up(tiles * 8) + align_up(16);
}

extern "C" int cgs_neural_gaussians_backward(const float *packed_fwd, const float *packed_bwd, const int32_t *vis_idx,
                                             int Nv, const float *anchor, const float *feat, const float *offsets,
                                             const float *scaling, const float *mask, const float *campos_host,
                                             const uint8_t *keep_mask, const float *g_xyz, const float *g_color,
                                             const float *g_opacity, const float *g_scaling, const float *g_rot,
                                             float *d_anchor, float *d_feat, float *d_offsets, float *d_scaling,
                                             float *d_mask, float *d_packed_fwd, void *workspace, size_t workspace_bytes,
                                             void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (Nv <= 0) return 0;
    CGS_CHECK_PTR(packed_fwd); CGS_CHECK_PTR(packed_bwd); CGS_CHECK_PTR(anchor); CGS_CHECK_PTR(feat);
    CGS_CHECK_PTR(offsets); CGS_CHECK_PTR(scaling); CGS_CHECK_PTR(mask); CGS_CHECK_PTR(campos_host);
    CGS_CHECK_PTR(keep_mask); CGS_CHECK_PTR(g_xyz); CGS_CHECK_PTR(g_color); CGS_CHECK_PTR(g_opacity);
    CGS_CHECK_PTR(g_scaling); CGS_CHECK_PTR(g_rot); CGS_CHECK_PTR(d_anchor); CGS_CHECK_PTR(d_feat);
    CGS_CHECK_PTR(d_offsets); CGS_CHECK_PTR(d_scaling); CGS_CHECK_PTR(d_mask); CGS_CHECK_PTR(d_packed_fwd);
    CGS_CHECK_PTR(workspace);
    if (workspace_bytes < cgs_neural_gaussians_backward_workspace_bytes(Nv)) {
        set_error("%s: workspace too small", __func__);
        return -3;
    }
    const int tiles = (Nv + kTM - 1) / kTM;
    char *ws = static_cast<char *>(workspace);
    cudaMemsetAsync(ws, 0, cgs_neural_gaussians_backward_workspace_bytes(Nv), st);
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(neural_gaussians_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(ngb::Smem));
        if (sm_count <= 0) sm_count = kNumSMs;
    }
    const int grid = tiles < sm_count ? tiles : sm_count;
    StageScope sc(ST_G1_BWD, st, 1);
    neural_gaussians_backward_kernel<<<grid, kMlpThreads, sizeof(ngb::Smem), st>>>(
        packed_fwd, packed_bwd, vis_idx, Nv, anchor, feat, offsets, scaling, mask, campos_host[0], campos_host[1],
        campos_host[2], keep_mask, g_xyz, g_color, g_opacity, g_scaling, g_rot, d_anchor, d_feat, d_offsets, d_scaling,
        d_mask, d_packed_fwd, reinterpret_cast<unsigned long long *>(ws),
        reinterpret_cast<uint32_t *>(ws + align_up((size_t)tiles * 8)));
    return check_launch(__func__);
}

// Fused anchor -> neural-Gaussian generation.
//
// One kernel replaces the reference's chain at gaussian_renderer/__init__.py:106-145:
// view direction/distance, the three decoder MLPs (scene/gaussian_model.py:153-174:
// 54->50 ReLU ->10 tanh | ->30 sigmoid | ->70), `neural_opacity * mask > 0` selection,
// the [Nv*K, 22] concat + boolean compaction, and the scale / rotation / position
// post-processing.  Nothing but the final per-Gaussian attributes is written to HBM
// (56 B per emitted Gaussian + 50 B per anchor for neural_opacity / selection mask).
//
// The boolean compaction is ORDER PRESERVING (Gaussians come out in (anchor, offset) order, as
// `tensor[mask]` does) through a chained scan over CTA tiles: downstream code indexes by
// position (training_statis, scene/gaussian_model.py:704-713; rasterizer tie-break).
#include "mlp_tile.cuh"

namespace cgs {

constexpr int kFeat = 50;
constexpr int kK = 10;            // n_offsets
constexpr int kIn = kFeat + 4;    // 54
constexpr int kHid = 3 * kFeat;   // 150 (opacity | color | cov hidden units)
constexpr int kOut = kK + 3 * kK + 7 * kK;  // 110
constexpr int kPairs = kTM * kK;  // 640 (anchor, offset) pairs per tile

// packed decoder weights (floats), produced once per weight version by the host shim:
//   W1[54][152] (k-major; cols 0-49 opacity, 50-99 color, 100-149 cov, 150-151 zero pad)
//   b1[152]
//   W2o[50][12] b2o[12] | W2c[50][32] b2c[32] | W2v[50][72] b2v[72]   (h-major, zero padded)
constexpr int kLd1 = 152, kLdO = 12, kLdC = 32, kLdV = 72;
constexpr int kOffW1 = 0;
constexpr int kOffB1 = kOffW1 + kIn * kLd1;
constexpr int kOffW2o = kOffB1 + kLd1;
constexpr int kOffB2o = kOffW2o + kFeat * kLdO;
constexpr int kOffW2c = kOffB2o + kLdO;
constexpr int kOffB2c = kOffW2c + kFeat * kLdC;
constexpr int kOffW2v = kOffB2c + kLdC;
constexpr int kOffB2v = kOffW2v + kFeat * kLdV;
constexpr int kPackedFloats = kOffB2v + kLdV;  // 14292 floats = 57 KB

struct NgSmem {
    float w[kPackedFloats];
    float x[kIn * kTMp];        // transposed inputs; later reused for the outputs' head
    float h[kHid * kTMp];       // hidden activations
    float out[kOut * kTMp];     // MLP outputs [o][r]
    float anchor[kTM * 3];
    float scaling[kTM * 6];
    int src[kTM];               // source anchor index of each tile row (-1 = padding)
    uint32_t warp_cnt[kMlpThreads / 32];
    uint32_t tile_base;
    uint32_t tile_id;
};


__global__ void __launch_bounds__(kMlpThreads, 1)
neural_gaussians_forward_kernel(const float *__restrict__ packed_w, const int *__restrict__ vis_idx, int Nv,
                                const float *__restrict__ anchor, const float *__restrict__ feat,
                                const float *__restrict__ offsets, const float *__restrict__ scaling,
                                const float *__restrict__ mask, float cx, float cy, float cz,
                                float *__restrict__ o_xyz, float *__restrict__ o_color, float *__restrict__ o_opacity,
                                float *__restrict__ o_scaling, float *__restrict__ o_rot,
                                float *__restrict__ o_neural_opacity, uint8_t *__restrict__ o_mask,
                                uint32_t *__restrict__ tile_prefix, unsigned long long *scan_state, uint32_t *ticket,
                                int32_t *__restrict__ count_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NgSmem &S = *reinterpret_cast<NgSmem *>(smem_raw);
    const int tid = threadIdx.x;
    if (tid == 0) S.tile_id = atomicAdd(ticket, 1u);
    copy_to_smem(S.w, packed_w, kPackedFloats);
    __syncthreads();
    const int tile = (int)S.tile_id;
    const int row0 = tile * kTM;
    const int num_tiles = (Nv + kTM - 1) / kTM;

    // ---- stage inputs: 4 threads per row -------------------------------------------------
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
                const float ax = anchor[3 * a], ay = anchor[3 * a + 1], az = anchor[3 * a + 2];
                const float vx = ax - cx, vy = ay - cy, vz = az - cz;
                const float d = sqrtf(vx * vx + vy * vy + vz * vz);
                S.x[(kFeat + 0) * kTMp + r] = vx / d;
                S.x[(kFeat + 1) * kTMp + r] = vy / d;
                S.x[(kFeat + 2) * kTMp + r] = vz / d;
                S.x[(kFeat + 3) * kTMp + r] = d;
                S.anchor[3 * r] = ax; S.anchor[3 * r + 1] = ay; S.anchor[3 * r + 2] = az;
            }
            if (q == 1)
                for (int k = 0; k < 6; ++k) S.scaling[6 * r + k] = scaling[(size_t)a * 6 + k];
        } else {
            for (int k = q; k < kIn; k += 4) S.x[k * kTMp + r] = 0.f;
        }
    }
    __syncthreads();

    // ---- layer 1 (54 -> 150, ReLU) and the three output heads ------------------------------
    tile_gemm<5, ACT_RELU>(S.x, kIn, S.w + kOffW1, kLd1, S.w + kOffB1, kHid, S.h);
    __syncthreads();
    tile_gemm<1, ACT_NONE>(S.h, kFeat, S.w + kOffW2o, kLdO, S.w + kOffB2o, kK, S.out);
    tile_gemm<1, ACT_NONE>(S.h + kFeat * kTMp, kFeat, S.w + kOffW2c, kLdC, S.w + kOffB2c, 3 * kK, S.out + kK * kTMp);
    tile_gemm<3, ACT_NONE>(S.h + 2 * kFeat * kTMp, kFeat, S.w + kOffW2v, kLdV, S.w + kOffB2v, 7 * kK,
                           S.out + 4 * kK * kTMp);
    __syncthreads();

    // ---- selection: pair p = r*K + k, thread handles pairs tid, tid+256, tid+512 -------------
    float nop[3];
    bool keep[3];
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int p = tid + i * kMlpThreads;
        nop[i] = 0.f;
        keep[i] = false;
        if (p < kPairs) {
            const int r = p / kK, k = p - r * kK;
            const int a = S.src[r];
            if (a >= 0) {
                nop[i] = tanhf(S.out[k * kTMp + r]) * mask[(size_t)a * kK + k];
                keep[i] = nop[i] > 0.0f;
                const size_t gp = (size_t)(row0 + r) * kK + k;
                o_neural_opacity[gp] = nop[i];
                o_mask[gp] = keep[i] ? 1 : 0;
            }
        }
        cnt += keep[i] ? 1u : 0u;
    }
    // Ordered ranks: pairs are interleaved over threads (p = tid + 256 i), so rank by segments:
    // segment i covers pairs [256 i, 256 i + 256); rank = (#kept in earlier segments) + (#kept by
    // lower threads in this segment).
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t seg_rank[3], seg_total[3];
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
    (void)cnt;

    // ---- chained scan over tiles -----------------------------------------------------------
    if (warp == 0) {
        const uint64_t excl = lookback_exclusive(scan_state, tile, tile_total);
        if (lane == 0) {
            S.tile_base = (uint32_t)excl;
            tile_prefix[tile] = (uint32_t)excl;
            if (tile == num_tiles - 1) *count_out = (int32_t)(excl + tile_total);
        }
    }
    __syncthreads();
    const uint32_t base = S.tile_base;

    // ---- emit ------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (!keep[i]) continue;
        const int p = tid + i * kMlpThreads;
        const int r = p / kK, k = p - r * kK;
        const int a = S.src[r];
        uint32_t pos = base + seg_rank[i];
        if (i >= 1) pos += seg_total[0];
        if (i >= 2) pos += seg_total[1];
        const float *of = offsets + ((size_t)a * kK + k) * 3;
        const float *sc = S.scaling + 6 * r;
        o_xyz[3 * (size_t)pos + 0] = S.anchor[3 * r + 0] + of[0] * sc[0];
        o_xyz[3 * (size_t)pos + 1] = S.anchor[3 * r + 1] + of[1] * sc[1];
        o_xyz[3 * (size_t)pos + 2] = S.anchor[3 * r + 2] + of[2] * sc[2];
        const float *oc = S.out + (kK + 3 * k) * kTMp + r;
        o_color[3 * (size_t)pos + 0] = 1.0f / (1.0f + expf(-oc[0]));
        o_color[3 * (size_t)pos + 1] = 1.0f / (1.0f + expf(-oc[kTMp]));
        o_color[3 * (size_t)pos + 2] = 1.0f / (1.0f + expf(-oc[2 * kTMp]));
        o_opacity[pos] = nop[i];
        const float *ov = S.out + (4 * kK + 7 * k) * kTMp + r;
        o_scaling[3 * (size_t)pos + 0] = sc[3] * (1.0f / (1.0f + expf(-ov[0])));
        o_scaling[3 * (size_t)pos + 1] = sc[4] * (1.0f / (1.0f + expf(-ov[kTMp])));
        o_scaling[3 * (size_t)pos + 2] = sc[5] * (1.0f / (1.0f + expf(-ov[2 * kTMp])));
        const float q0 = ov[3 * kTMp], q1 = ov[4 * kTMp], q2 = ov[5 * kTMp], q3 = ov[6 * kTMp];
        const float nrm = fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);  // F.normalize eps
        reinterpret_cast<float4 *>(o_rot)[pos] = make_float4(q0 / nrm, q1 / nrm, q2 / nrm, q3 / nrm);
    }
}

// Ordered stream compaction of a boolean mask into an index list (decoupled look-back scan).
constexpr int kCompactItems = 8, kCompactThreads = 256, kCompactTile = kCompactItems * kCompactThreads;

// MaskT = uint8_t: select mask[i] != 0;  MaskT = int32_t: select mask[i] > 0 (screen radii of visible_filter)
template <typename MaskT>
__global__ void __launch_bounds__(kCompactThreads)
compact_indices_kernel(const MaskT *__restrict__ mask, int N, int *__restrict__ out_idx,
                       unsigned long long *scan_state, uint32_t *ticket, int32_t *__restrict__ count_out)
{
    __shared__ uint32_t s_tile, s_warp[kCompactThreads / 32], s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base_i = tile * kCompactTile + threadIdx.x * kCompactItems;
    uint32_t bits = 0;
    if (sizeof(MaskT) == 4) {
        if (base_i + kCompactItems <= N) {
            const int4 v0 = *reinterpret_cast<const int4 *>(mask + base_i);
            const int4 v1 = *reinterpret_cast<const int4 *>(mask + base_i + 4);
            const int v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) bits |= v[i] > 0 ? (1u << i) : 0u;
        } else {
            for (int i = 0; i < kCompactItems; ++i)
                if (base_i + i < N && (int)mask[base_i + i] > 0) bits |= 1u << i;
        }
    } else if (base_i + kCompactItems <= N) {
        const uint2 v = *reinterpret_cast<const uint2 *>(mask + base_i);  // 8 mask bytes
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bits |= ((v.x >> (8 * i)) & 0xff) ? (1u << i) : 0u;
            bits |= ((v.y >> (8 * i)) & 0xff) ? (1u << (4 + i)) : 0u;
        }
    } else {
        for (int i = 0; i < kCompactItems; ++i)
            if (base_i + i < N && mask[base_i + i]) bits |= 1u << i;
    }
    const uint32_t c = __popc(bits);
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t wex = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kCompactThreads / 32; ++w) {
        wex += w < warp ? s_warp[w] : 0u;
        total += s_warp[w];
    }
    if (warp == 0) {
        const uint64_t excl = lookback_exclusive(scan_state, tile, total);
        if (lane == 0) {
            s_base = (uint32_t)excl;
            if (tile == (N - 1) / kCompactTile) *count_out = (int32_t)(excl + total);
        }
    }
    __syncthreads();
    uint32_t pos = s_base + wex + incl - c;
    for (int i = 0; i < kCompactItems; ++i)
        if (bits & (1u << i)) out_idx[pos++] = base_i + i;
}


// ---- densification statistics (scene/gaussian_model.py:696-713 `training_statis`) --------------------------
// One thread per visible anchor: opacity_accum += sum_k max(neural_opacity, 0), anchor_demon += 1.
__global__ void __launch_bounds__(256)
statis_anchor_kernel(const int *__restrict__ vis_idx, const int32_t *__restrict__ n_vis, int K,
                     const float *__restrict__ opacity, float *__restrict__ opacity_accum, float *__restrict__ anchor_demon)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= *n_vis) return;
    const int a = vis_idx[v];
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += fmaxf(opacity[(size_t)v * K + k], 0.f);
    opacity_accum[a] += s;
    anchor_demon[a] += 1.0f;
}

// One thread per emitted Gaussian p (= the p-th kept (visible anchor, offset) slot): if it was drawn
// (update_filter), add the norm of its screen-space gradient to the offset's accumulator.
__global__ void __launch_bounds__(256)
statis_offset_kernel(const int *__restrict__ vis_idx, const int *__restrict__ kept_slot, const int32_t *__restrict__ n_kept,
                     int P, int K, const float *__restrict__ vgrad, const uint8_t *__restrict__ update_filter,
                     float *__restrict__ grad_accum, float *__restrict__ denom)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P || p >= *n_kept || !update_filter[p]) return;
    const int slot = kept_slot[p];
    const int v = slot / K, k = slot - v * K;
    const size_t dst = (size_t)vis_idx[v] * K + k;
    const float gx = vgrad[3 * (size_t)p], gy = vgrad[3 * (size_t)p + 1];
    grad_accum[dst] += sqrtf(gx * gx + gy * gy);
    denom[dst] += 1.0f;
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_neural_gaussians_packed_floats(void) { return kPackedFloats; }

extern "C" size_t cgs_neural_gaussians_workspace_bytes(int Nv)
{
    const size_t tiles = (size_t)(Nv > 0 ? (Nv + kTM - 1) / kTM : 1);
    return align_up(tiles * 8) + align_up(16) + align_up(tiles * 4);
}

extern "C" int cgs_neural_gaussians_forward(const float *packed_weights, const int32_t *vis_idx, int Nv,
                                            const float *anchor, const float *feat, const float *offsets,
                                            const float *scaling, const float *mask, const float *campos_host,
                                            float *o_xyz, float *o_color, float *o_opacity, float *o_scaling,
                                            float *o_rot, float *o_neural_opacity, uint8_t *o_mask,
                                            int32_t *count_dev, void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CGS_CHECK_PTR(count_dev);
    if (Nv <= 0) {
        cudaMemsetAsync(count_dev, 0, sizeof(int32_t), st);
        return check_launch(__func__);
    }
    CGS_CHECK_PTR(packed_weights);
    CGS_CHECK_PTR(anchor);
    CGS_CHECK_PTR(feat);
    CGS_CHECK_PTR(offsets);
    CGS_CHECK_PTR(scaling);
    CGS_CHECK_PTR(mask);
    CGS_CHECK_PTR(campos_host);
    CGS_CHECK_PTR(o_xyz);
    CGS_CHECK_PTR(o_color);
    CGS_CHECK_PTR(o_opacity);
    CGS_CHECK_PTR(o_scaling);
    CGS_CHECK_PTR(o_rot);
    CGS_CHECK_PTR(o_neural_opacity);
    CGS_CHECK_PTR(o_mask);
    CGS_CHECK_PTR(workspace);
    if (workspace_bytes < cgs_neural_gaussians_workspace_bytes(Nv)) {
        set_error("%s: workspace too small", __func__);
        return -3;
    }
    const int tiles = (Nv + kTM - 1) / kTM;
    char *ws = static_cast<char *>(workspace);
    unsigned long long *scan_state = reinterpret_cast<unsigned long long *>(ws);
    uint32_t *ticket = reinterpret_cast<uint32_t *>(ws + align_up((size_t)tiles * 8));
    uint32_t *tile_prefix = reinterpret_cast<uint32_t *>(ws + align_up((size_t)tiles * 8) + align_up(16));
    cudaMemsetAsync(ws, 0, align_up((size_t)tiles * 8) + align_up(16), st);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(neural_gaussians_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(NgSmem));
        attr_set = true;
    }
    StageScope sc(ST_G1_FWD, st, 1);
    neural_gaussians_forward_kernel<<<tiles, kMlpThreads, sizeof(NgSmem), st>>>(
        packed_weights, vis_idx, Nv, anchor, feat, offsets, scaling, mask, campos_host[0], campos_host[1],
        campos_host[2], o_xyz, o_color, o_opacity, o_scaling, o_rot, o_neural_opacity, o_mask, tile_prefix, scan_state,
        ticket, count_dev);
    return check_launch(__func__);
}

extern "C" size_t cgs_compact_workspace_bytes(int N)
{
    const size_t tiles = (size_t)(N > 0 ? (N + kCompactTile - 1) / kCompactTile : 1);
    return align_up(tiles * 8) + align_up(16);
}

extern "C" int cgs_compact_indices(const uint8_t *mask, int N, int32_t *out_idx, int32_t *count_dev, void *workspace,
                                   size_t workspace_bytes, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CGS_CHECK_PTR(count_dev);
    if (N <= 0) {
        cudaMemsetAsync(count_dev, 0, sizeof(int32_t), st);
        return check_launch(__func__);
    }
    CGS_CHECK_PTR(mask);
    CGS_CHECK_PTR(out_idx);
    CGS_CHECK_PTR(workspace);
    if (reinterpret_cast<uintptr_t>(mask) & 7) {
        set_error("%s: mask must be 8-byte aligned", __func__);
        return -2;
    }
    if (workspace_bytes < cgs_compact_workspace_bytes(N)) {
        set_error("%s: workspace too small", __func__);
        return -3;
    }
    const int tiles = (N + kCompactTile - 1) / kCompactTile;
    char *ws = static_cast<char *>(workspace);
    cudaMemsetAsync(ws, 0, cgs_compact_workspace_bytes(N), st);
    StageScope sc(ST_COMPACT, st, 1);
    compact_indices_kernel<uint8_t><<<tiles, kCompactThreads, 0, st>>>(
        mask, N, out_idx, reinterpret_cast<unsigned long long *>(ws),
        reinterpret_cast<uint32_t *>(ws + align_up((size_t)tiles * 8)), count_dev);
    return check_launch(__func__);
}

extern "C" int cgs_compact_positive_i32(const int32_t *values, int N, int32_t *out_idx, int32_t *count_dev,
                                        void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CGS_CHECK_PTR(count_dev);
    if (N <= 0) {
        cudaMemsetAsync(count_dev, 0, sizeof(int32_t), st);
        return check_launch(__func__);
    }
    CGS_CHECK_PTR(values);
    CGS_CHECK_PTR(out_idx);
    CGS_CHECK_PTR(workspace);
    if (reinterpret_cast<uintptr_t>(values) & 15) {
        set_error("%s: values must be 16-byte aligned", __func__);
        return -2;
    }
    if (workspace_bytes < cgs_compact_workspace_bytes(N)) {
        set_error("%s: workspace too small", __func__);
        return -3;
    }
    const int tiles = (N + kCompactTile - 1) / kCompactTile;
    char *ws = static_cast<char *>(workspace);
    cudaMemsetAsync(ws, 0, cgs_compact_workspace_bytes(N), st);
    StageScope sc(ST_COMPACT, st, 1);
    compact_indices_kernel<int32_t><<<tiles, kCompactThreads, 0, st>>>(
        values, N, out_idx, reinterpret_cast<unsigned long long *>(ws),
        reinterpret_cast<uint32_t *>(ws + align_up((size_t)tiles * 8)), count_dev);
    return check_launch(__func__);
}

extern "C" size_t cgs_training_statis_workspace_bytes(int N, int K)
{
    const size_t n = (size_t)(N > 0 ? N : 1), slots = n * (size_t)(K > 0 ? K : 1);
    return align_up(n * 4) + align_up(slots * 4) + 2 * cgs_compact_workspace_bytes((int)slots) + align_up(16);
}

extern "C" int cgs_training_statis(int N, int K, const uint8_t *anchor_visible, const uint8_t *offset_selection, int n_slots,
                                   const float *neural_opacity, const float *viewspace_grad, const uint8_t *update_filter,
                                   int P, float *opacity_accum, float *anchor_demon, float *offset_gradient_accum,
                                   float *offset_denom, void *workspace, size_t workspace_bytes, void *stream)
{
    if (N <= 0 || n_slots <= 0) return 0;
    CGS_CHECK_PTR(anchor_visible); CGS_CHECK_PTR(offset_selection); CGS_CHECK_PTR(neural_opacity);
    CGS_CHECK_PTR(opacity_accum); CGS_CHECK_PTR(anchor_demon); CGS_CHECK_PTR(offset_gradient_accum);
    CGS_CHECK_PTR(offset_denom); CGS_CHECK_PTR(workspace);
    if (P > 0) {
        CGS_CHECK_PTR(viewspace_grad);
        CGS_CHECK_PTR(update_filter);
    }
    if (K <= 0 || n_slots % K != 0 || n_slots > (int64_t)N * K) {
        set_error("%s: n_slots must be (visible anchors) * K", __func__);
        return -2;
    }
    if (workspace_bytes < cgs_training_statis_workspace_bytes(N, K)) {
        set_error("%s: workspace too small", __func__);
        return -3;
    }
    if ((reinterpret_cast<uintptr_t>(anchor_visible) | reinterpret_cast<uintptr_t>(offset_selection)) & 7) {
        set_error("%s: masks must be 8-byte aligned", __func__);
        return -2;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *ws = static_cast<char *>(workspace);
    int *vis_idx = reinterpret_cast<int *>(ws);
    ws += align_up((size_t)N * 4);
    int *kept = reinterpret_cast<int *>(ws);
    ws += align_up((size_t)N * K * 4);
    int32_t *counts = reinterpret_cast<int32_t *>(ws);   // [0] visible anchors, [1] kept slots
    ws += align_up(16);
    const size_t cws = cgs_compact_workspace_bytes(N * K);
    if (int e = cgs_compact_indices(anchor_visible, N, vis_idx, counts, ws, cws, stream)) return e;
    if (int e = cgs_compact_indices(offset_selection, n_slots, kept, counts + 1, ws + cws, cws, stream)) return e;
    StageScope sc(ST_ELEMWISE, st, 2);
    const int nv_cap = n_slots / K;
    statis_anchor_kernel<<<(nv_cap + 255) / 256, 256, 0, st>>>(vis_idx, counts, K, neural_opacity, opacity_accum, anchor_demon);
    if (P > 0)
        statis_offset_kernel<<<(P + 255) / 256, 256, 0, st>>>(vis_idx, kept, counts + 1, P, K, viewspace_grad, update_filter,
                                                            offset_gradient_accum, offset_denom);
    return check_launch(__func__);
}

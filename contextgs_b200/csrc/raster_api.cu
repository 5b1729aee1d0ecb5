// C-ABI entry points of the rasterizer (include/contextgs_b200.h) and the error plumbing.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace cgs {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
        return -100 - (int)e;
    }
    return 0;
}

// ---- stage timing --------------------------------------------------------------------------
static const char *const kStageNames[ST_COUNT] = {
    "visible_filter", "compact_indices", "neural_gaussians_fwd", "neural_gaussians_bwd", "preprocess", "depth_sort",
    "scan_emit_pairs", "bin_expand", "pair_sort", "tile_ranges", "render_fwd", "render_bwd", "preprocess_bwd",
    "entropy_bottleneck", "context_level_fwd", "context_level_bwd", "gaussian_bits", "elementwise", "level_divide",
    "entropy_codec", "l1_ssim_loss", "anchor_growing"};

struct StageTimer {
    std::mutex mu;
    bool enabled = false;
    std::vector<cudaEvent_t> pool;      // event pairs: [2*i] begin, [2*i+1] end
    std::vector<int> slot_stage;        // stage of each used pair
    double ms_sum[ST_COUNT] = {0};
    int64_t scopes[ST_COUNT] = {0};
};
static StageTimer g_timer;
static std::atomic<int64_t> g_launches[ST_COUNT];

StageScope::StageScope(int stage_, cudaStream_t st_, int kernels) : stage(stage_), st(st_), slot(-1)
{
    g_launches[stage].fetch_add(kernels, std::memory_order_relaxed);
    if (!g_timer.enabled) return;
    std::lock_guard<std::mutex> lk(g_timer.mu);
    slot = (int)g_timer.slot_stage.size();
    if (g_timer.pool.size() < (size_t)(2 * slot + 2)) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        g_timer.pool.push_back(a);
        g_timer.pool.push_back(b);
    }
    g_timer.slot_stage.push_back(stage);
    cudaEventRecord(g_timer.pool[2 * slot], st);
}

StageScope::~StageScope()
{
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_timer.mu);
    cudaEventRecord(g_timer.pool[2 * slot + 1], st);
}

// launchers defined in the kernel translation units
void launch_preprocess(const CamParams &, int, const uint32_t *, const float *, const float *, const float *,
                       const float *, const float *, int32_t *, float *, uint32_t *, uint2 *, cudaStream_t);
void launch_filter(const CamParams &, int, const float *, const float *, const float *, int32_t *, cudaStream_t);
void launch_mark_visible(const CamParams &, int, const float *, uint8_t *, cudaStream_t);
void launch_prefilter_anchors(const CamParams &, int, const float *, const float *, int, const float *, uint8_t *, int *,
                              unsigned long long *, uint32_t *, int32_t *, cudaStream_t);
void launch_preprocess_backward(const CamParams &, int, const float *, const float *, const float *, const int32_t *,
                                const float *, float *, float *, float *, float *, float *, float *, cudaStream_t);
void launch_scan_emit_pairs(const uint32_t *, const uint2 *, int, const uint32_t *, int64_t, int, uint32_t *, uint32_t *,
                            unsigned long long *, uint32_t *, unsigned long long *, uint32_t *, uint32_t *, int32_t *,
                            cudaStream_t);
int bin_chunks_cap(int64_t R_cap);
void launch_bin_expand(const uint32_t *, const uint32_t *, const uint2 *, const uint32_t *, int64_t, int, int, uint32_t *,
                       uint32_t *, uint32_t *, uint32_t *, uint32_t *, uint32_t *, uint32_t *, uint32_t *, cudaStream_t);
int scan_tiles_count(int P);
void launch_render_forward(const CamParams &, const uint32_t *, const uint32_t *, const float *, float *, float *,
                           uint32_t *, cudaStream_t);
void launch_render_backward(const CamParams &, const uint32_t *, const uint32_t *, const float *, const float *,
                            const uint32_t *, const float *, float *, cudaStream_t);

__global__ void set_u32_kernel(uint32_t *p, uint32_t v) { *p = v; }
// device-side Gaussian count, clamped to the capacity the buffers were sized for
__global__ void copy_count_kernel(uint32_t *p, const int32_t *src, uint32_t cap, int32_t *status)
{
    const int32_t v = *src;
    const uint32_t c = v < 0 ? 0u : (uint32_t)v;
    *p = c < cap ? c : cap;
    status[3] = v;                    // CGS_STATUS_NUM_GAUSSIANS
    status[4] = c > cap ? 1 : 0;      // CGS_STATUS_GAUSSIAN_OVERFLOW
}

static int tile_bits(int tiles)
{
    int b = 1;
    while ((1 << b) < tiles) ++b;
    return b;
}

// Workspace layout of the forward pass.  Everything that must start at zero is grouped at the
// front so that one memset covers it.
struct RasterPlan {
    SortPlan depth_sort, pair_sort;
    size_t depth_ws_off, pair_ws_off, scan_state_off, scan_ticket_off, tile_total_off, seg_off, zero_bytes;
    size_t depth_keys_off[3], depth_vals_off[2], rects_off;    // keys: in / tmp / out ; vals: tmp / out
    size_t pair_keys_off[3], pair_vals_off[3];                 // in / tmp / out
    size_t tail_cnt_off, run_prefix_off, tile_start_off;
    size_t total;
    int sbits, num_super;
};

static RasterPlan make_plan(int P, int64_t R_cap, int W, int H)
{
    RasterPlan p;
    const int gx = (W + kTile - 1) / kTile, gy = (H + kTile - 1) / kTile;
    p.num_super = ((gx + CGS_SUPER_X - 1) / CGS_SUPER_X) * ((gy + CGS_SUPER_Y - 1) / CGS_SUPER_Y);
    p.sbits = tile_bits(p.num_super);
    const int64_t Pc = P > 0 ? P : 1, Rc = R_cap > 0 ? R_cap : 1;
    p.depth_sort = make_sort_plan(Pc, 0, 32);
    p.pair_sort = make_sort_plan(Rc, 0, p.sbits);
    size_t off = 0;
    p.depth_ws_off = off;
    off += p.depth_sort.total_bytes;
    p.pair_ws_off = off;
    off += p.pair_sort.total_bytes;
    p.scan_state_off = off;
    off += align_up((size_t)scan_tiles_count((int)Pc) * sizeof(unsigned long long));
    p.scan_ticket_off = off;   // ticket, P, pair count, done counter, [4..5] 64-bit instance accumulator
    off += align_up(8 * sizeof(uint32_t));
    p.tile_total_off = off;
    off += align_up((size_t)gx * gy * sizeof(uint32_t));
    p.seg_off = off;           // list begin / end per super-tile
    off += align_up((size_t)p.num_super * 2 * sizeof(uint32_t));
    p.zero_bytes = off;
    for (int i = 0; i < 3; ++i) {
        p.depth_keys_off[i] = off;
        off += align_up((size_t)Pc * 4);
    }
    for (int i = 0; i < 2; ++i) {
        p.depth_vals_off[i] = off;
        off += align_up((size_t)Pc * 4);
    }
    p.rects_off = off;
    off += align_up((size_t)Pc * 8);
    for (int i = 0; i < 3; ++i) {
        p.pair_keys_off[i] = off;
        off += align_up((size_t)Rc * 4);
    }
    for (int i = 0; i < 3; ++i) {
        p.pair_vals_off[i] = off;
        off += align_up((size_t)Rc * 4);
    }
    const size_t chunks = (size_t)bin_chunks_cap(Rc) + 2;
    p.tail_cnt_off = off;
    off += align_up(chunks * 32 * sizeof(uint32_t));
    p.run_prefix_off = off;
    off += align_up(chunks * 32 * sizeof(uint32_t));
    p.tile_start_off = off;
    off += align_up((size_t)gx * gy * sizeof(uint32_t));
    p.total = off;
    return p;
}

}  // namespace cgs

using namespace cgs;

extern "C" int cgs_abi_version(void) { return CGS_ABI_VERSION; }
extern "C" const char *cgs_last_error(void) { return g_error; }

extern "C" int cgs_stage_count(void) { return ST_COUNT; }
extern "C" const char *cgs_stage_name(int i) { return i >= 0 && i < ST_COUNT ? kStageNames[i] : ""; }

extern "C" int cgs_stage_timing_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_timer.mu);
    g_timer.enabled = on != 0;
    g_timer.slot_stage.clear();
    for (int i = 0; i < ST_COUNT; ++i) {
        g_timer.ms_sum[i] = 0;
        g_timer.scopes[i] = 0;
    }
    return 0;
}

extern "C" int cgs_stage_timing_read(double *ms_sum, int64_t *scopes)
{
    std::lock_guard<std::mutex> lk(g_timer.mu);
    for (size_t i = 0; i < g_timer.slot_stage.size(); ++i) {
        float ms = 0.f;
        if (cudaEventSynchronize(g_timer.pool[2 * i + 1]) != cudaSuccess ||
            cudaEventElapsedTime(&ms, g_timer.pool[2 * i], g_timer.pool[2 * i + 1]) != cudaSuccess) {
            set_error("cgs_stage_timing_read: event query failed");
            cudaGetLastError();
            return -1;
        }
        g_timer.ms_sum[g_timer.slot_stage[i]] += ms;
        g_timer.scopes[g_timer.slot_stage[i]] += 1;
    }
    g_timer.slot_stage.clear();
    for (int i = 0; i < ST_COUNT; ++i) {
        if (ms_sum) ms_sum[i] = g_timer.ms_sum[i];
        if (scopes) scopes[i] = g_timer.scopes[i];
    }
    return 0;
}

extern "C" int cgs_launch_counts(int64_t *launches, int reset)
{
    for (int i = 0; i < ST_COUNT; ++i) {
        if (launches) launches[i] = g_launches[i].load();
        if (reset) g_launches[i].store(0);
    }
    return 0;
}

static int validate_settings(const cgs_raster_settings *s, const char *fn)
{
    if (!s) {
        set_error("%s: null settings", fn);
        return -1;
    }
    if (s->image_width <= 0 || s->image_height <= 0) {
        set_error("%s: invalid image size %dx%d", fn, s->image_width, s->image_height);
        return -2;
    }
    if ((int64_t)((s->image_width + kTile - 1) / kTile) * ((s->image_height + kTile - 1) / kTile) > (1 << 30)) {
        set_error("%s: too many tiles", fn);
        return -2;
    }
    return 0;
}

extern "C" int cgs_visible_filter(const cgs_raster_settings *s, int N, const float *means3D, const float *scales,
                                  const float *rotations, int32_t *radii, void *stream)
{
    if (int e = validate_settings(s, __func__)) return e;
    if (N <= 0) return 0;
    CGS_CHECK_PTR(means3D);
    CGS_CHECK_PTR(scales);
    CGS_CHECK_PTR(rotations);
    CGS_CHECK_PTR(radii);
    StageScope sc(ST_FILTER, static_cast<cudaStream_t>(stream), 1);
    launch_filter(make_cam(s), N, means3D, scales, rotations, radii, static_cast<cudaStream_t>(stream));
    return check_launch(__func__);
}

extern "C" size_t cgs_prefilter_workspace_bytes(int N)
{
    const size_t tiles = (size_t)(N > 0 ? (N + 255) / 256 : 1);
    return align_up(tiles * 8) + align_up(16);
}

extern "C" int cgs_prefilter_anchors(const cgs_raster_settings *s, int N, const float *anchor, const float *scales,
                                     int scale_stride, const float *rotation_row, uint8_t *visible, int32_t *vis_idx,
                                     int32_t *count_dev, void *workspace, size_t workspace_bytes, void *stream)
{
    if (int e = validate_settings(s, __func__)) return e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CGS_CHECK_PTR(count_dev);
    if (N <= 0) {
        cudaMemsetAsync(count_dev, 0, sizeof(int32_t), st);
        return check_launch(__func__);
    }
    CGS_CHECK_PTR(anchor);
    CGS_CHECK_PTR(scales);
    CGS_CHECK_PTR(rotation_row);
    CGS_CHECK_PTR(visible);
    CGS_CHECK_PTR(vis_idx);
    CGS_CHECK_PTR(workspace);
    if (scale_stride < 3) {
        set_error("%s: scale_stride %d < 3", __func__, scale_stride);
        return -2;
    }
    if (workspace_bytes < cgs_prefilter_workspace_bytes(N)) {
        set_error("%s: workspace too small", __func__);
        return -3;
    }
    const size_t tiles = (size_t)(N + 255) / 256;
    char *ws = static_cast<char *>(workspace);
    cudaMemsetAsync(ws, 0, cgs_prefilter_workspace_bytes(N), st);
    StageScope sc(ST_FILTER, st, 1);
    launch_prefilter_anchors(make_cam(s), N, anchor, scales, scale_stride, rotation_row, visible, vis_idx,
                             reinterpret_cast<unsigned long long *>(ws),
                             reinterpret_cast<uint32_t *>(ws + align_up(tiles * 8)), count_dev, st);
    return check_launch(__func__);
}

extern "C" int cgs_render_anchors_forward(const cgs_raster_settings *s, const float *packed_weights,
                                          const int32_t *vis_idx, int Nv_cap, const int32_t *nv_dev, int P_cap,
                                          const float *anchor, const float *feat, const float *offsets,
                                          const float *scaling, const float *mask, float *g_xyz, float *g_color,
                                          float *g_opacity, float *g_scaling, float *g_rot, int32_t *g_count,
                                          void *g1_workspace, size_t g1_workspace_bytes, int64_t R_cap, float *out_color,
                                          int32_t *radii, float *geom, uint32_t *point_list, uint32_t *ranges,
                                          float *final_T, uint32_t *n_contrib, int32_t *status, void *workspace,
                                          size_t workspace_bytes, void *stream)
{
    if (int e = validate_settings(s, __func__)) return e;
    if (int e = cgs_neural_gaussians_umma_forward_dev(packed_weights, vis_idx, Nv_cap, nv_dev, P_cap, anchor, feat, offsets,
                                                      scaling, mask, s->campos, g_xyz, g_color, g_opacity, g_scaling, g_rot,
                                                      nullptr, nullptr, g_count, g1_workspace, g1_workspace_bytes, stream))
        return e;
    return cgs_rasterize_forward_dev(s, P_cap, g_count, g_xyz, g_color, g_opacity, g_scaling, g_rot, R_cap, out_color, radii,
                                     geom, point_list, ranges, final_T, n_contrib, status, workspace, workspace_bytes,
                                     stream);
}

extern "C" int cgs_mark_visible(const cgs_raster_settings *s, int N, const float *means3D, uint8_t *visible,
                                void *stream)
{
    if (int e = validate_settings(s, __func__)) return e;
    if (N <= 0) return 0;
    CGS_CHECK_PTR(means3D);
    CGS_CHECK_PTR(visible);
    StageScope sc(ST_FILTER, static_cast<cudaStream_t>(stream), 1);
    launch_mark_visible(make_cam(s), N, means3D, visible, static_cast<cudaStream_t>(stream));
    return check_launch(__func__);
}

extern "C" size_t cgs_raster_workspace_bytes(int P, int64_t R_cap, int W, int H)
{
    return make_plan(P, R_cap, W, H).total;
}

extern "C" int cgs_rasterize_forward(const cgs_raster_settings *s, int P, const float *means3D, const float *colors,
                                     const float *opacities, const float *scales, const float *rotations,
                                     int64_t R_cap, float *out_color, int32_t *radii, float *geom,
                                     uint32_t *point_list, uint32_t *ranges, float *final_T, uint32_t *n_contrib,
                                     int32_t *status, void *workspace, size_t workspace_bytes, void *stream)
{
    return cgs_rasterize_forward_dev(s, P, nullptr, means3D, colors, opacities, scales, rotations, R_cap, out_color, radii,
                                     geom, point_list, ranges, final_T, n_contrib, status, workspace, workspace_bytes,
                                     stream);
}

extern "C" int cgs_rasterize_forward_dev(const cgs_raster_settings *s, int P, const int32_t *P_dev, const float *means3D,
                                         const float *colors, const float *opacities, const float *scales,
                                         const float *rotations, int64_t R_cap, float *out_color, int32_t *radii,
                                         float *geom, uint32_t *point_list, uint32_t *ranges, float *final_T,
                                         uint32_t *n_contrib, int32_t *status, void *workspace, size_t workspace_bytes,
                                         void *stream)
{
    if (int e = validate_settings(s, __func__)) return e;
    CGS_CHECK_PTR(out_color);
    CGS_CHECK_PTR(ranges);
    CGS_CHECK_PTR(final_T);
    CGS_CHECK_PTR(n_contrib);
    CGS_CHECK_PTR(status);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CamParams cam = make_cam(s);
    const int tiles = cam.grid_x * cam.grid_y;
    if (P < 0 || R_cap < 0 || R_cap >= (1ll << 30)) {
        set_error("%s: invalid size (P %d, R_cap %lld; R_cap must be in [0, 2^30))", __func__, P, (long long)R_cap);
        return -2;
    }
    cudaMemsetAsync(status, 0, CGS_STATUS_WORDS * sizeof(int32_t), st);
    cudaMemsetAsync(ranges, 0, (size_t)tiles * 2 * sizeof(uint32_t), st);
    if (P > 0) {
        CGS_CHECK_PTR(means3D);
        CGS_CHECK_PTR(colors);
        CGS_CHECK_PTR(opacities);
        CGS_CHECK_PTR(scales);
        CGS_CHECK_PTR(rotations);
        CGS_CHECK_PTR(radii);
        CGS_CHECK_PTR(geom);
        CGS_CHECK_PTR(workspace);
        if (R_cap > 0) CGS_CHECK_PTR(point_list);
        const RasterPlan plan = make_plan(P, R_cap, cam.W, cam.H);
        if (workspace_bytes < plan.total) {
            set_error("%s: workspace %zu < %zu bytes", __func__, workspace_bytes, plan.total);
            return -3;
        }
        char *ws = static_cast<char *>(workspace);
        cudaMemsetAsync(ws, 0, plan.zero_bytes, st);
        auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(ws + off); };
        uint32_t *dkeys_in = u32(plan.depth_keys_off[0]), *dkeys_tmp = u32(plan.depth_keys_off[1]);
        uint32_t *dkeys_out = u32(plan.depth_keys_off[2]);
        uint32_t *dvals_tmp = u32(plan.depth_vals_off[0]), *order = u32(plan.depth_vals_off[1]);
        uint2 *rects = reinterpret_cast<uint2 *>(ws + plan.rects_off);
        uint32_t *pkeys_in = u32(plan.pair_keys_off[0]), *pkeys_tmp = u32(plan.pair_keys_off[1]);
        uint32_t *pkeys_out = u32(plan.pair_keys_off[2]);
        uint32_t *pvals_in = u32(plan.pair_vals_off[0]), *pvals_tmp = u32(plan.pair_vals_off[1]);
        uint32_t *pvals_out = u32(plan.pair_vals_off[2]);
        unsigned long long *scan_state = reinterpret_cast<unsigned long long *>(ws + plan.scan_state_off);
        uint32_t *scan_ticket = reinterpret_cast<uint32_t *>(ws + plan.scan_ticket_off);
        // device-side counts parked in the zeroed ticket block
        uint32_t *p_dev = scan_ticket + 1, *n_pairs_dev = scan_ticket + 2, *scan_done = scan_ticket + 3;
        unsigned long long *r_acc = reinterpret_cast<unsigned long long *>(scan_ticket + 4);
        uint32_t *tile_total = u32(plan.tile_total_off), *seg_begin = u32(plan.seg_off);
        uint32_t *seg_end = seg_begin + plan.num_super;

        {
            StageScope sc(ST_PREPROCESS, st, 2);
            if (P_dev) copy_count_kernel<<<1, 1, 0, st>>>(p_dev, P_dev, (uint32_t)P, status);
            else set_u32_kernel<<<1, 1, 0, st>>>(p_dev, (uint32_t)P);
            launch_preprocess(cam, P, p_dev, means3D, colors, opacities, scales, rotations, radii, geom, dkeys_in, rects, st);
        }
        // 1. Gaussians by depth (value = Gaussian id, implicit iota on the first pass)
        {
            StageScope sc(ST_DEPTH_SORT, st, 2 + plan.depth_sort.npass);
            if (int e = sort_pairs(dkeys_in, nullptr, dkeys_out, order, dkeys_tmp, dvals_tmp, p_dev, P, 0, 32,
                                   ws + plan.depth_ws_off, false, st))
                return e;
        }
        // 2. (super-tile, id) pairs in depth order; R and the pair count stay on the device
        {
            StageScope sc(ST_SCAN, st, 1);
            launch_scan_emit_pairs(order, rects, P, p_dev, R_cap, (cam.grid_x + CGS_SUPER_X - 1) / CGS_SUPER_X, pkeys_in,
                                   pvals_in, scan_state, scan_ticket, r_acc, scan_done, n_pairs_dev, status, st);
        }
        if (R_cap > 0) {
            // 3. stable sort of the pairs by super-tile
            {
                StageScope sc(ST_TILE_SORT, st, 2 + plan.pair_sort.npass);
                if (int e = sort_pairs(pkeys_in, pvals_in, pkeys_out, pvals_out, pkeys_tmp, pvals_tmp, n_pairs_dev, R_cap, 0,
                                       plan.sbits, ws + plan.pair_ws_off, false, st))
                    return e;
            }
            // 4. expansion into per-tile lists + ranges
            StageScope sc(ST_EMIT, st, 4);
            launch_bin_expand(pkeys_out, pvals_out, rects, n_pairs_dev, R_cap, cam.grid_x, cam.grid_y,
                              u32(plan.tail_cnt_off), u32(plan.run_prefix_off), seg_begin, seg_end, tile_total,
                              u32(plan.tile_start_off), ranges, point_list, st);
        }
    }
    {
        StageScope sc(ST_RENDER_FWD, st, 1);
        launch_render_forward(cam, ranges, point_list, geom, out_color, final_T, n_contrib, st);
    }
    return check_launch(__func__);
}

extern "C" size_t cgs_raster_backward_workspace_bytes(int P) { return align_up((size_t)(P > 0 ? P : 1) * 9 * 4); }

extern "C" int cgs_rasterize_backward(const cgs_raster_settings *s, int P, const float *means3D, const float *scales,
                                      const float *rotations, const int32_t *radii, const float *geom,
                                      const uint32_t *point_list, const uint32_t *ranges, const float *final_T,
                                      const uint32_t *n_contrib, const float *dL_dpix, float *dL_dmeans3D,
                                      float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity, float *dL_dscales,
                                      float *dL_drots, void *workspace, size_t workspace_bytes, void *stream)
{
    if (int e = validate_settings(s, __func__)) return e;
    if (P <= 0) return 0;
    CGS_CHECK_PTR(means3D);
    CGS_CHECK_PTR(scales);
    CGS_CHECK_PTR(rotations);
    CGS_CHECK_PTR(radii);
    CGS_CHECK_PTR(geom);
    CGS_CHECK_PTR(ranges);
    CGS_CHECK_PTR(final_T);
    CGS_CHECK_PTR(n_contrib);
    CGS_CHECK_PTR(dL_dpix);
    CGS_CHECK_PTR(dL_dmeans3D);
    CGS_CHECK_PTR(dL_dmeans2D);
    CGS_CHECK_PTR(dL_dcolors);
    CGS_CHECK_PTR(dL_dopacity);
    CGS_CHECK_PTR(dL_dscales);
    CGS_CHECK_PTR(dL_drots);
    CGS_CHECK_PTR(workspace);
    if (workspace_bytes < cgs_raster_backward_workspace_bytes(P)) {
        set_error("%s: workspace too small", __func__);
        return -3;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const CamParams cam = make_cam(s);
    float *acc = static_cast<float *>(workspace);
    cudaMemsetAsync(acc, 0, (size_t)P * 9 * sizeof(float), st);
    {
        StageScope sc(ST_RENDER_BWD, st, 1);
        launch_render_backward(cam, ranges, point_list, geom, final_T, n_contrib, dL_dpix, acc, st);
    }
    StageScope sc(ST_PRE_BWD, st, 1);
    launch_preprocess_backward(cam, P, means3D, scales, rotations, radii, acc, dL_dmeans3D, dL_dmeans2D, dL_dcolors,
                               dL_dopacity, dL_dscales, dL_drots, st);
    return check_launch(__func__);
}

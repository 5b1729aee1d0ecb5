// Fused photometric loss of the training step: L1 + SSIM of train.py:200-204,
// i.e. utils/loss_utils.py `l1_loss` (:17-18) and `ssim` / `_ssim` (:33-64): 11x11 Gaussian window
// (sigma 1.5, separable, zero padding 5), C1 = 0.01^2, C2 = 0.03^2, mean over all pixels and channels.
//
// The reference runs five grouped 11x11 convolutions (121 taps each), ~15 elementwise kernels and the
// autograd mirror of all of them.  Here:
//   forward : one kernel; a CTA stages a (16+10)^2 tile of both images in shared memory, runs the
//             separable filter for the five moments (x, y, x^2, y^2, xy) and writes, per pixel, the three
//             partial derivatives of the SSIM map w.r.t. the filtered moments of x
//                 dm = d ssim / d (G*x),  dp = d ssim / d (G*x^2),  dq = d ssim / d (G*xy)
//             plus the two fp64 sums (|x - y|, ssim);
//   backward: one kernel; the adjoint of a zero-padded symmetric filter is the same filter, so
//                 d ssim_mean / d x = ( G*dm + 2x G*dp + y G*dq ) / n
//             and dL/dx = g_l1 * sign(x - y) / n + g_ssim * that.
// Only the rendered image receives a gradient (the ground truth is data).
#include "common.cuh"

namespace cgs {
namespace loss {

constexpr int kWin = 11, kHalo = 5, kTileL = 16, kExt = kTileL + 2 * kHalo;   // 26
constexpr float kC1 = 0.01f * 0.01f, kC2 = 0.03f * 0.03f;

struct Window {
    float g[kWin];
};

__global__ void __launch_bounds__(kTileL * kTileL)
l1_ssim_forward_kernel(Window win, const float *__restrict__ img, const float *__restrict__ gt, int H, int W,
                       float *__restrict__ dm, float *__restrict__ dp, float *__restrict__ dq, double *__restrict__ sums)
{
    __shared__ float sx[kExt][kExt + 1], sy[kExt][kExt + 1];
    __shared__ float sh[5][kExt][kTileL + 1];
    __shared__ float red[2][kTileL * kTileL / 32];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kTileL, y0 = blockIdx.y * kTileL;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kTileL + tx;
    const float *ic = img + (size_t)c * H * W, *gc = gt + (size_t)c * H * W;
    for (int i = tid; i < kExt * kExt; i += kTileL * kTileL) {
        const int ly = i / kExt, lx = i - ly * kExt;
        const int gy = y0 + ly - kHalo, gx = x0 + lx - kHalo;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        sx[ly][lx] = in ? ic[(size_t)gy * W + gx] : 0.f;
        sy[ly][lx] = in ? gc[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    // horizontal pass: 26 rows x 16 columns x 5 moments
    for (int i = tid; i < kExt * kTileL; i += kTileL * kTileL) {
        const int ly = i / kTileL, lx = i - ly * kTileL;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win.g[k], x = sx[ly][lx + k], y = sy[ly][lx + k];
            a += w * x; b += w * y; aa += w * (x * x); bb += w * (y * y); ab += w * (x * y);
        }
        sh[0][ly][lx] = a; sh[1][ly][lx] = b; sh[2][ly][lx] = aa; sh[3][ly][lx] = bb; sh[4][ly][lx] = ab;
    }
    __syncthreads();
    float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
        const float w = win.g[k];
        mu1 += w * sh[0][ty + k][tx]; mu2 += w * sh[1][ty + k][tx]; e11 += w * sh[2][ty + k][tx];
        e22 += w * sh[3][ty + k][tx]; e12 += w * sh[4][ty + k][tx];
    }
    const int gx = x0 + tx, gy = y0 + ty;
    float l1 = 0.f, sv = 0.f;
    if (gx < W && gy < H) {
        const float s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
        const float A = 2.f * mu1 * mu2 + kC1, B = 2.f * s12 + kC2, C = mu1 * mu1 + mu2 * mu2 + kC1, D = s11 + s22 + kC2;
        const float inv_cd = 1.0f / (C * D);
        sv = A * B * inv_cd;
        l1 = fabsf(sx[ty + kHalo][tx + kHalo] - sy[ty + kHalo][tx + kHalo]);
        if (dm) {
            const size_t p = (size_t)c * H * W + (size_t)gy * W + gx;
            dm[p] = 2.f * mu2 * (B - A) * inv_cd - sv * 2.f * mu1 * (D - C) * inv_cd;
            dp[p] = -sv / D;
            dq[p] = 2.f * A * inv_cd;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = l1;
        red[1][tid >> 5] = sv;
    }
    __syncthreads();
    if (tid < 2) {
        float t = 0.f;
        for (int w = 0; w < kTileL * kTileL / 32; ++w) t += red[tid][w];
        atomicAdd(&sums[tid], (double)t);
    }
}

__global__ void __launch_bounds__(kTileL * kTileL)
l1_ssim_backward_kernel(Window win, const float *__restrict__ img, const float *__restrict__ gt, int H, int W,
                        const float *__restrict__ dm, const float *__restrict__ dp, const float *__restrict__ dq,
                        const float *__restrict__ g_l1, const float *__restrict__ g_ssim, float *__restrict__ d_img)
{
    __shared__ float s[3][kExt][kExt + 1];
    __shared__ float sh[3][kExt][kTileL + 1];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * kTileL, y0 = blockIdx.y * kTileL;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * kTileL + tx;
    const size_t plane = (size_t)c * H * W;
    const bool with_ssim = dm != nullptr && g_ssim != nullptr;
    float acc_m = 0.f, acc_p = 0.f, acc_q = 0.f;
    if (with_ssim) {
        for (int i = tid; i < kExt * kExt; i += kTileL * kTileL) {
            const int ly = i / kExt, lx = i - ly * kExt;
            const int gy = y0 + ly - kHalo, gx = x0 + lx - kHalo;
            const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
            const size_t p = plane + (size_t)gy * W + gx;
            s[0][ly][lx] = in ? dm[p] : 0.f;
            s[1][ly][lx] = in ? dp[p] : 0.f;
            s[2][ly][lx] = in ? dq[p] : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < kExt * kTileL; i += kTileL * kTileL) {
            const int ly = i / kTileL, lx = i - ly * kTileL;
            float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                const float w = win.g[k];
                a += w * s[0][ly][lx + k]; b += w * s[1][ly][lx + k]; d += w * s[2][ly][lx + k];
            }
            sh[0][ly][lx] = a; sh[1][ly][lx] = b; sh[2][ly][lx] = d;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win.g[k];
            acc_m += w * sh[0][ty + k][tx]; acc_p += w * sh[1][ty + k][tx]; acc_q += w * sh[2][ty + k][tx];
        }
    }
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx < W && gy < H) {
        const size_t p = plane + (size_t)gy * W + gx;
        const float x = img[p], y = gt[p];
        const float inv_n = 1.0f / (3.0f * (float)H * (float)W);
        float g = 0.f;
        if (g_l1) g += __ldg(g_l1) * (x > y ? 1.f : (x < y ? -1.f : 0.f)) * inv_n;
        if (with_ssim) g += __ldg(g_ssim) * (acc_m + 2.f * x * acc_p + y * acc_q) * inv_n;
        d_img[p] = g;
    }
}

static Window make_window()
{
    // utils/loss_utils.py:23-25: exp(-(x - 5)^2 / (2 * 1.5^2)) normalised, evaluated in double and rounded to
    // float32 like torch.Tensor([...]) / sum
    Window w;
    float g[kWin], sum = 0.f;
    for (int i = 0; i < kWin; ++i) {
        g[i] = (float)exp(-(double)((i - kWin / 2) * (i - kWin / 2)) / (2.0 * 1.5 * 1.5));
        sum += g[i];
    }
    for (int i = 0; i < kWin; ++i) w.g[i] = g[i] / sum;
    return w;
}

}  // namespace loss
}  // namespace cgs

using namespace cgs;

extern "C" int cgs_l1_ssim_forward(const float *img, const float *gt, int H, int W, float *dm, float *dp, float *dq,
                                   double *sums, void *stream)
{
    CGS_CHECK_PTR(img); CGS_CHECK_PTR(gt); CGS_CHECK_PTR(sums);
    if (H <= 0 || W <= 0) {
        set_error("%s: invalid image size %dx%d", __func__, W, H);
        return -2;
    }
    if ((dm == nullptr) != (dp == nullptr) || (dm == nullptr) != (dq == nullptr)) {
        set_error("%s: dm, dp, dq must all be given or all be NULL", __func__);
        return -2;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(sums, 0, 2 * sizeof(double), st);
    dim3 grid((W + loss::kTileL - 1) / loss::kTileL, (H + loss::kTileL - 1) / loss::kTileL, 3), block(loss::kTileL, loss::kTileL);
    StageScope sc(ST_LOSS, st, 1);
    loss::l1_ssim_forward_kernel<<<grid, block, 0, st>>>(loss::make_window(), img, gt, H, W, dm, dp, dq, sums);
    return check_launch(__func__);
}

extern "C" int cgs_l1_ssim_backward(const float *img, const float *gt, int H, int W, const float *dm, const float *dp,
                                    const float *dq, const float *g_l1, const float *g_ssim, float *d_img, void *stream)
{
    CGS_CHECK_PTR(img); CGS_CHECK_PTR(gt); CGS_CHECK_PTR(d_img);
    if (H <= 0 || W <= 0) {
        set_error("%s: invalid image size %dx%d", __func__, W, H);
        return -2;
    }
    if (g_ssim && !(dm && dp && dq)) {
        set_error("%s: the SSIM gradient needs the dm / dp / dq maps of the forward", __func__);
        return -2;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((W + loss::kTileL - 1) / loss::kTileL, (H + loss::kTileL - 1) / loss::kTileL, 3), block(loss::kTileL, loss::kTileL);
    StageScope sc(ST_LOSS, st, 1);
    loss::l1_ssim_backward_kernel<<<grid, block, 0, st>>>(loss::make_window(), img, gt, H, W, dm, dp, dq, g_l1, g_ssim, d_img);
    return check_launch(__func__);
}

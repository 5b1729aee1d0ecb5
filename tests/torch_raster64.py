"""Differentiable float64 PyTorch restatement of the rasterizer forward, used ONLY to pin the
C oracle's backward with autograd (tests/test_oracle_raster.py).  Tiny scenes only."""
import math

import torch


def render64(st, means, colors, opac, scales, rots, radii, W, H):
    """st: oracle RefSettings.  All tensors float64 with requires_grad where wanted.
    `radii` (int, from the oracle) defines the non-differentiable tile footprint."""
    dt = torch.float64
    vm = torch.tensor(list(st.view), dtype=dt).reshape(4, 4)  # transposed storage: vm[r][c] = flat[4r+c]
    pm = torch.tensor(list(st.proj), dtype=dt).reshape(4, 4)
    ones = torch.ones(means.shape[0], 1, dtype=dt)
    ph = torch.cat([means, ones], 1)
    t = ph @ vm  # [P,4] view space (x_view = sum_j p_j * flat[4j+0])
    hom = ph @ pm
    pw = 1.0 / (hom[:, 3] + 1e-7)
    ndc = hom[:, :2] * pw[:, None]
    px = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    py = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    r, x, y, z = rots[:, 0], rots[:, 1], rots[:, 2], rots[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    N = R * (st.scale_modifier * scales)[:, None, :]
    Sigma = N @ N.transpose(1, 2)
    fx = W / (2.0 * st.tanfovx)
    fy = H / (2.0 * st.tanfovy)
    tz = t[:, 2]
    limx, limy = 1.3 * st.tanfovx, 1.3 * st.tanfovy
    tx = torch.clamp(t[:, 0] / tz, -limx, limx) * tz
    ty = torch.clamp(t[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz), zero, fy / tz, -(fy * ty) / (tz * tz)], 1).reshape(-1, 2, 3)
    Rv = vm[:3, :3].T  # Rv(i,j) = flat[i + 4j]
    A = J @ Rv
    cov = A @ Sigma @ A.transpose(1, 2)
    cx = cov[:, 0, 0] + 0.3
    cy = cov[:, 0, 1]
    cz = cov[:, 1, 1] + 0.3
    det = cx * cz - cy * cy
    ca, cb, cc = cz / det, -cy / det, cx / det
    order = torch.argsort(tz.detach(), stable=True)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    T = torch.ones(H, W, dtype=dt)
    C = torch.zeros(3, H, W, dtype=dt)
    done = torch.zeros(H, W, dtype=torch.bool)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    for g in order.tolist():
        rad = int(radii[g])
        if rad <= 0:
            continue
        pxf, pyf = float(torch.tensor(px[g].item(), dtype=torch.float32)), float(torch.tensor(py[g].item(), dtype=torch.float32))
        x0 = min(gx, max(0, int((pxf - rad) / 16.0)))
        y0 = min(gy, max(0, int((pyf - rad) / 16.0)))
        x1 = min(gx, max(0, int((pxf + rad + 15) / 16.0)))
        y1 = min(gy, max(0, int((pyf + rad + 15) / 16.0)))
        foot = torch.zeros(H, W, dtype=torch.bool)
        foot[y0 * 16:y1 * 16, x0 * 16:x1 * 16] = True
        dx = px[g] - xs
        dy = py[g] - ys
        power = -0.5 * (ca[g] * dx * dx + cc[g] * dy * dy) - cb[g] * dx * dy
        G = torch.exp(power)
        alpha_raw = opac[g, 0] * G
        # upstream: the 0.99 clamp is NOT differentiated through (gradient passes as if unclamped)
        alpha = alpha_raw + (torch.clamp(alpha_raw, max=0.99) - alpha_raw).detach()
        valid = foot & (~done) & (power.detach() <= 0) & (alpha.detach() >= 1.0 / 255.0)
        test_T = T * (1 - alpha)
        stop = valid & (test_T.detach() < 1e-4)
        done = done | stop
        use = valid & (~stop)
        w = torch.where(use, alpha * T, torch.zeros_like(T))
        C = C + colors[g][:, None, None] * w[None]
        T = torch.where(use, test_T, T)
    bg = torch.tensor(list(st.bg), dtype=dt)
    return C + T[None] * bg[:, None, None]

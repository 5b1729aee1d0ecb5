"""GPU integration: a short training run wired exactly like the reference's loop (train.py:150-259) --
prefilter_voxel -> render -> L1 + SSIM (+ bit-rate term) -> backward -> training_statis -> adjust_anchor -> Adam step --
through the drop-in surface only.  Growing / pruning replaces every per-anchor Parameter mid-run, so this is the test
that the version-keyed caches (level plan, packed MLP weights, global means, frame scratch) follow the model."""
import types

import pytest
import torch

from contextgs_b200 import codec, synthetic
from contextgs_b200.gaussian_model import GaussianModel
from contextgs_b200.loss_utils import l1_ssim
from contextgs_b200.renderer import prefilter_voxel, render
from tests.test_checkpoint_cpu import _args

pytestmark = pytest.mark.gpu


def test_short_training_run_with_growing_and_pruning():
    N, W, H = 6000, 256, 160
    scene = synthetic.make_scene("chair", N, seed=7, gaussian_scale=4.0)
    torch.manual_seed(11)
    pc = GaussianModel.from_tensors(scene, device="cuda").train()
    pc.training_setup(_args())
    cams = synthetic.make_cameras("chair", 6, device="cuda", W=W, H=H)
    pipe = types.SimpleNamespace(debug=False)
    bg = torch.zeros(3, device="cuda")
    with torch.no_grad():          # targets: the initial renders, recoloured -- reachable by training colours / opacities
        gts = []
        for cam in cams:
            vis = prefilter_voxel(cam, pc, pipe, bg)
            img = render(cam, pc, pipe, bg, visible_mask=vis, retain_grad=False, step=0)["render"]
            gts.append((0.6 * img.flip(0)).clamp(0, 1).detach())
    losses, sizes = [], [pc._anchor.shape[0]]
    for it in range(1, 61):
        pc.update_learning_rate(it)
        cam, gt = cams[it % len(cams)], gts[it % len(cams)]
        vis = prefilter_voxel(cam, pc, pipe, bg)
        step = it if it <= 40 else 10000 + it           # the context model joins the loss after iteration 10000
        pkg = render(cam, pc, pipe, bg, visible_mask=vis, retain_grad=True, step=step)
        Ll1, ssim_value = l1_ssim(pkg["render"], gt)
        loss = 0.8 * Ll1 + 0.2 * (1.0 - ssim_value) + 0.01 * pkg["scaling"].prod(dim=1).mean()
        if pkg["bit_per_param"] is not None:
            loss = loss + 0.004 * pkg["bit_per_param"] + 5e-4 * torch.mean(torch.sigmoid(pc._mask))
        loss.backward()
        losses.append(float(loss.detach()))
        with torch.no_grad():
            pc.training_statis(pkg["viewspace_points"], pkg["neural_opacity"], pkg["visibility_filter"], pkg["selection_mask"], vis)
            if it in (24, 42, 55):
                pc.adjust_anchor(check_interval=10, success_threshold=0.8, grad_threshold=2e-6, min_opacity=0.005)
                sizes.append(pc._anchor.shape[0])
            pc.optimizer.step()
            pc.optimizer.zero_grad(set_to_none=True)
    n = pc._anchor.shape[0]
    assert len(set(sizes)) > 1, sizes                                   # the anchor set actually changed
    for name, cols in (("_offset", (n, 10, 3)), ("_mask", (n, 10, 1)), ("_anchor_feat", (n, 50)), ("_hyper_latent", (n, 12)),
                       ("_scaling", (n, 6)), ("_rotation", (n, 4)), ("_opacity", (n, 1))):
        t = getattr(pc, name)
        assert tuple(t.shape) == cols and bool(torch.isfinite(t).all()), name
    assert pc.opacity_accum.shape == (n, 1) and pc.offset_denom.shape == (n * 10, 1)
    # the same six cameras before any growing: 18 Adam steps must have lowered the loss
    assert sum(losses[17:23]) < 0.97 * sum(losses[:6]), (losses[:6], losses[17:23])
    # the grown model goes through the rest of the surface: scoring, bitstream round trip, inference render
    pc.eval()
    sums = pc.estimate_final_bits(return_values=True)
    assert all(float(v) >= 0 for v in sums) and float(sums[2]) > 0
    enc = codec.encode_model(pc)
    fresh = GaussianModel(device="cuda")
    fresh.load_state_dict({k: v for k, v in pc.state_dict().items() if not k.startswith("_")}, strict=False)
    out = codec.decode_model(fresh, enc.meta, enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes, enc.hyper_lens,
                             enc.levels)
    for k in ("feat", "scaling", "hyper", "anchor"):
        assert torch.equal(out[k], enc.quantised[k]), k
    with torch.no_grad():
        vis = prefilter_voxel(cams[0], pc, pipe, bg)
        img = render(cams[0], pc, pipe, bg, visible_mask=vis)["render"]
    assert img.shape == (3, H, W) and bool(torch.isfinite(img).all())

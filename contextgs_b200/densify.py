"""Optimiser set-up, anchor growing and pruning of the reference's `GaussianModel`
(scene/gaussian_model.py:426-555 training_setup / update_learning_rate, :657-694 optimiser surgery,
:715-759 pruning, :762-854 anchor_growing, :856-910 adjust_anchor), as a mixin of
contextgs_b200.gaussian_model.GaussianModel with the same method names, arguments and Adam group names
(densify / prune rewrite the optimiser state by group name).

What runs where: the candidate test of one growing depth is a handful of elementwise torch ops (and torch's own
generator, so a seeded run draws what the reference draws); everything the reference does per depth after that --
the cell coordinates, torch.unique(dim=0), the chunked all-pairs occupancy test and torch_scatter.scatter_max --
is ONE library call (csrc/anchor_growing.cu) followed by ONE read of the new-anchor count.
"""
import math

import torch
import torch.nn as nn

from . import _lib

ANCHOR_GROUPS = ("anchor", "offset", "mask", "anchor_feat", "hyper_latent", "opacity", "scaling", "rotation")
_SKIP = ("mlp", "conv", "feat_base", "encoding", "codec")      # groups that are not per-anchor (:676, :718)


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000, step_sub=0):
    """utils/general_utils.py:49-82: log-linear interpolation from lr_init to lr_final over max_steps."""
    def helper(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        delay = 1.0
        if lr_delay_steps > 0:
            delay = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0), 1))
        t = min(max((step - step_sub) / (max_steps - step_sub), 0), 1)
        return delay * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)
    return helper


class DensifyMixin:
    update_depth, update_init_factor, update_hierachy_factor = 3, 16, 4      # arguments/__init__.py:53-55
    spatial_lr_scale = 1.0
    percent_dense = 0.0
    optimizer = None

    # ---- optimiser (scene/gaussian_model.py:426-555) -----------------------------------------------------
    def training_setup(self, training_args):
        a, s = training_args, self.spatial_lr_scale
        self.percent_dense = a.percent_dense
        self.opacity_accum = self.anchor_demon = None
        self._ensure_statis()
        groups = [
            {"params": [self._anchor], "lr": a.position_lr_init * s, "name": "anchor"},
            {"params": [self._offset], "lr": a.offset_lr_init * s, "name": "offset"},
            {"params": [self._mask], "lr": a.mask_lr_init * s, "name": "mask"},
            {"params": [self._anchor_feat], "lr": a.feature_lr, "name": "anchor_feat"},
            {"params": [self._hyper_latent], "lr": a.hyper_latent_lr, "name": "hyper_latent"},
            {"params": [self._opacity], "lr": a.opacity_lr, "name": "opacity"},
            {"params": [self._scaling], "lr": a.scaling_lr, "name": "scaling"},
            {"params": [self._rotation], "lr": a.rotation_lr, "name": "rotation"},
            {"params": self.mlp_opacity.parameters(), "lr": a.mlp_opacity_lr_init, "name": "mlp_opacity"},
            {"params": self.mlp_cov.parameters(), "lr": a.mlp_cov_lr_init, "name": "mlp_cov"},
            {"params": self.mlp_color.parameters(), "lr": a.mlp_color_lr_init, "name": "mlp_color"},
            {"params": self.latent_codec.parameters(), "lr": a.latent_codec_lr_init, "name": "latent_codec"},
            {"params": self.mlp_grid.parameters(), "lr": a.mlp_grid_lr_init, "name": "mlp_grid"},
        ]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        sched = lambda n, scale=1.0: get_expon_lr_func(
            lr_init=getattr(a, n + "_lr_init") * scale, lr_final=getattr(a, n + "_lr_final") * scale,
            lr_delay_mult=getattr(a, n + "_lr_delay_mult"), max_steps=getattr(a, n + "_lr_max_steps"))
        self._schedules = {"anchor": sched("position", s), "offset": sched("offset", s), "mask": sched("mask", s),
                           "mlp_opacity": sched("mlp_opacity"), "mlp_cov": sched("mlp_cov"), "mlp_color": sched("mlp_color"),
                           "latent_codec": sched("latent_codec"), "mlp_grid": sched("mlp_grid")}

    def update_learning_rate(self, iteration):
        for group in self.optimizer.param_groups:
            f = self._schedules.get(group["name"])
            if f is not None:
                group["lr"] = f(iteration)

    # ---- optimiser surgery (:657-694, :715-745) -------------------------------------------------------------
    def _per_anchor_groups(self):
        for group in self.optimizer.param_groups:
            if any(t in group["name"] for t in _SKIP):
                continue
            assert len(group["params"]) == 1
            yield group

    def _swap(self, group, new_value, state):
        old = group["params"][0]
        self.optimizer.state.pop(old, None)
        p = nn.Parameter(new_value.requires_grad_(True))
        group["params"][0] = p
        if state is not None:
            self.optimizer.state[p] = state
        return p

    def _rebind(self, tensors):
        for name in ANCHOR_GROUPS:
            if name in tensors:
                setattr(self, "_" + name, tensors[name])

    def replace_tensor_to_optimizer(self, tensor, name):
        out = {}
        for group in self.optimizer.param_groups:
            if group["name"] == name:
                state = self.optimizer.state.get(group["params"][0], None)
                state["exp_avg"], state["exp_avg_sq"] = torch.zeros_like(tensor), torch.zeros_like(tensor)
                out[name] = self._swap(group, tensor, state)
        return out

    def cat_tensors_to_optimizer(self, tensors_dict):
        out = {}
        for group in self._per_anchor_groups():
            ext = tensors_dict[group["name"]]
            old = group["params"][0]
            state = self.optimizer.state.get(old, None)
            if state is not None:
                for k in ("exp_avg", "exp_avg_sq"):
                    state[k] = torch.cat((state[k], torch.zeros_like(ext)), dim=0)
            out[group["name"]] = self._swap(group, torch.cat((old.detach(), ext), dim=0), state)
        return out

    def _prune_anchor_optimizer(self, mask):
        out = {}
        for group in self._per_anchor_groups():
            old = group["params"][0]
            state = self.optimizer.state.get(old, None)
            if state is not None:
                for k in ("exp_avg", "exp_avg_sq"):
                    state[k] = state[k][mask]
            kept = old.detach()[mask]
            if group["name"] == "scaling":
                kept[:, 3:].clamp_(max=0.05)              # :729-733 (log-space columns 3..5)
            out[group["name"]] = self._swap(group, kept, state)
        return out

    def prune_anchor(self, mask):
        self._rebind(self._prune_anchor_optimizer(~mask))

    # ---- growing (:762-854) ----------------------------------------------------------------------------------
    @torch.no_grad()
    def grow_cells(self, candidate_mask, cur_size, n_candidates=None):
        """One depth of :778-816 on the device: returns (candidate_anchor[M,3], new_feat[M,50], new_hyper[M,12]) for the
        unoccupied cells of size `cur_size` that the candidate (anchor, offset) slots fall into, in the order of the
        sorted unique cells."""
        L = _lib.lib()
        N, K, dev = self._anchor.shape[0], self.n_offsets, self._anchor.device
        cand = candidate_mask.reshape(-1).contiguous()
        if cand.numel() != N * K:
            raise ValueError("grow_cells: candidate_mask must hold one flag per (anchor, offset) slot")
        cand = cand.view(torch.uint8) if cand.dtype == torch.bool else cand.to(torch.uint8)
        M = int(cand.sum()) if n_candidates is None else int(n_candidates)
        H = self._hyper_latent.shape[1]
        if M == 0 or N == 0:
            e = lambda c: torch.zeros((0, c), device=dev)
            return e(3), e(self.feat_dim), e(H)
        anchor_q = self.get_anchor.detach().contiguous()
        scaling = self.get_scaling.detach().contiguous()
        new_anchor = torch.empty((M, 3), device=dev)
        new_feat = torch.empty((M, self.feat_dim), device=dev)
        new_hyper = torch.empty((M, H), device=dev)
        status = torch.empty(5, dtype=torch.int32, device=dev)
        ws = torch.empty((L.cgs_anchor_growing_workspace_bytes(N, K, M),), dtype=torch.uint8, device=dev)
        offset, feat = self._offset.detach().contiguous(), self._anchor_feat.detach().contiguous()   # named: outlive the launch
        hyper = self._hyper_latent.detach().contiguous()
        _lib.check(L.cgs_anchor_growing(
            _lib.ptr(anchor_q), _lib.ptr(offset), _lib.ptr(scaling), scaling.shape[1],
            _lib.ptr(feat), self.feat_dim, _lib.ptr(hyper),
            H, _lib.ptr(cand), N, K, float(cur_size), M, _lib.ptr(new_anchor), _lib.ptr(new_feat), _lib.ptr(new_hyper), M,
            _lib.ptr(status), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "cgs_anchor_growing")
        n_new, bad_range, _, _, overflow = status.tolist()
        if bad_range:
            raise _lib.CgsError("anchor_growing: a cell coordinate left the supported +-2^20 range "
                                f"(cell size {cur_size}); the scene extent / voxel size ratio is too large")
        if overflow:
            raise _lib.CgsError("anchor_growing: candidate count exceeded the sized workspace")
        return new_anchor[:n_new], new_feat[:n_new], new_hyper[:n_new]

    @torch.no_grad()
    def anchor_growing(self, grads, threshold, offset_mask, rand=None):
        """Same effect as :762-854.  rand: optional list of pre-drawn uniform tensors, one per depth (parity tests);
        by default they are drawn with torch.rand_like in the reference's order."""
        K, dev = self.n_offsets, self._anchor.device
        init_length = self._anchor.shape[0] * K
        for i in range(self.update_depth):
            cur_threshold = threshold * ((self.update_hierachy_factor // 2) ** i)
            candidate = (grads >= cur_threshold) & offset_mask
            r = rand[i].to(dev) if rand is not None else torch.rand_like(candidate.float())
            candidate &= r > (0.5 ** (i + 1))
            length_inc = self._anchor.shape[0] * K - init_length
            if length_inc == 0:
                if i > 0:
                    continue                     # :774-776: finer depths only run once a coarser one has added anchors
            else:
                candidate = torch.cat([candidate, torch.zeros(length_inc, dtype=torch.bool, device=dev)], dim=0)
            cur_size = self.voxel_size * (self.update_init_factor // (self.update_hierachy_factor ** i))
            candidate_anchor, new_feat, new_hyper = self.grow_cells(candidate, cur_size)
            m = candidate_anchor.shape[0]
            if m == 0:
                continue
            rot = torch.zeros((m, 4), device=dev)
            rot[:, 0] = 1.0
            d = {
                "anchor": candidate_anchor,
                "scaling": torch.log(torch.full((m, 6), 1.0, device=dev) * cur_size),
                "rotation": rot,
                "anchor_feat": new_feat,
                "hyper_latent": new_hyper,
                "offset": torch.zeros((m, K, 3), device=dev),
                "mask": torch.ones((m, K, 1), device=dev),
                "opacity": torch.log(torch.full((m, 1), 0.1, device=dev) / (1 - torch.full((m, 1), 0.1, device=dev))),
            }
            pad = torch.zeros((m, 1), device=dev)
            self.anchor_demon = torch.cat([self.anchor_demon, pad], dim=0)
            self.opacity_accum = torch.cat([self.opacity_accum, pad], dim=0)
            self._rebind(self.cat_tensors_to_optimizer(d))

    @torch.no_grad()
    def adjust_anchor(self, check_interval=100, success_threshold=0.8, grad_threshold=0.0002, min_opacity=0.005, rand=None):
        """:856-910 (train.py:246-247)."""
        K = self.n_offsets
        grads = self.offset_gradient_accum / self.offset_denom
        grads[grads.isnan()] = 0.0
        grads_norm = torch.norm(grads, dim=-1)
        offset_mask = (self.offset_denom > check_interval * success_threshold * 0.5).squeeze(dim=1)

        self.anchor_growing(grads_norm, grad_threshold, offset_mask, rand=rand)

        n_slots = self._anchor.shape[0] * K
        for name in ("offset_denom", "offset_gradient_accum"):
            t = getattr(self, name)
            t[offset_mask] = 0
            setattr(self, name, torch.cat([t, torch.zeros((n_slots - t.shape[0], 1), dtype=t.dtype, device=t.device)], dim=0))

        prune_mask = (self.opacity_accum < min_opacity * self.anchor_demon).squeeze(dim=1)
        anchors_mask = (self.anchor_demon > check_interval * success_threshold).squeeze(dim=1)
        prune_mask = torch.logical_and(prune_mask, anchors_mask)
        keep = ~prune_mask
        for name in ("offset_denom", "offset_gradient_accum"):
            setattr(self, name, getattr(self, name).view(-1, K)[keep].reshape(-1, 1))
        self.opacity_accum[anchors_mask] = 0
        self.anchor_demon[anchors_mask] = 0
        self.opacity_accum = self.opacity_accum[keep]
        self.anchor_demon = self.anchor_demon[keep]
        if prune_mask.shape[0] > 0:
            self.prune_anchor(prune_mask)
        self.max_radii2D = torch.zeros((self._anchor.shape[0],), device=self._anchor.device)

"""Host-side mirror of the hot-path surface of the reference's `GaussianModel`
(scene/gaussian_model.py:46-190, getters :288-345): the per-anchor parameters, the three decoder
MLPs, the three context ("grid") MLPs and the hyper-prior entropy bottleneck, with the same
attribute names the reference's glue reads (gaussian_renderer/__init__.py:31-142).

The rows of SURVEY.md 8f are mixed in from their own modules: optimiser set-up, anchor growing and pruning
(densify.py), the bitstream codec (codec.py), point_cloud.ply IO (ply_io.py), densification statistics (below).
MLP checkpoint files stay torch.save / torch.load of the state dict.
"""
import math

import torch
import torch.nn as nn

from . import _lib
from .densify import DensifyMixin
from .encodings import Quantize_anchor


class EntropyBottleneck(nn.Module):
    """Parameter container + CUDA forward of compressai's EntropyBottleneck (channels = feat_dim //
    hyper_divisor; reference: scene/gaussian_model.py:135,1556).  Same parameter names/shapes as
    CompressAI (matrices / biases / factors / quantiles) so that checkpoints map one-to-one."""

    def __init__(self, channels, filters=(3, 3, 3, 3), init_scale=10.0, tail_mass=1e-9):
        super().__init__()
        self.channels, self.filters = int(channels), tuple(int(f) for f in filters)
        f = (1,) + self.filters + (1,)
        scale = init_scale ** (1 / (len(self.filters) + 1))
        self.matrices, self.biases, self.factors = nn.ParameterList(), nn.ParameterList(), nn.ParameterList()
        for i in range(len(self.filters) + 1):
            init = math.log(math.expm1(1 / scale / f[i + 1]))
            self.matrices.append(nn.Parameter(torch.full((channels, f[i + 1], f[i]), init)))
            self.biases.append(nn.Parameter(torch.rand(channels, f[i + 1], 1) - 0.5))
            if i < len(self.filters):
                self.factors.append(nn.Parameter(torch.zeros(channels, f[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.tensor([-init_scale, 0.0, init_scale]).repeat(channels, 1, 1))
        self._packed = None
        self._packed_key = None

    def packed(self):
        """[C, 59] parameter block of the CUDA kernel (softplus / tanh pre-applied), cached per
        parameter version."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is None or self._packed_key != key:
            with torch.no_grad():
                cols = []
                for i in range(len(self.filters) + 1):
                    cols.append(torch.nn.functional.softplus(self.matrices[i]).reshape(self.channels, -1))
                    cols.append(self.biases[i].reshape(self.channels, -1))
                    if i < len(self.filters):
                        cols.append(torch.tanh(self.factors[i]).reshape(self.channels, -1))
                cols.append(self.quantiles[:, 0, 1:2])
                self._packed = torch.cat(cols, dim=1).float().contiguous()
            assert self._packed.shape[1] == _lib.lib().cgs_eb_param_floats()
            self._packed_key = key
        return self._packed

    def packed_diff(self):
        """The same [C, 59] block built with autograd recording: the backward kernel returns the gradient
        of this tensor (cgs_eb_backward) and autograd carries it through softplus / tanh into the parameters."""
        cols = []
        for i in range(len(self.filters) + 1):
            cols.append(torch.nn.functional.softplus(self.matrices[i]).reshape(self.channels, -1))
            cols.append(self.biases[i].reshape(self.channels, -1))
            if i < len(self.filters):
                cols.append(torch.tanh(self.factors[i]).reshape(self.channels, -1))
        cols.append(self.quantiles[:, 0, 1:2])
        return torch.cat(cols, dim=1).float().contiguous()

    @torch.no_grad()
    def forward(self, x, training=None, noise=None, choose=None, bit_sum=None):
        """x [N,C] -> (x_hat, likelihood).  training=True adds U(-.5,.5) (pass `noise` [N,C] for
        reproducibility), otherwise rounds about the median.  No autograd: the differentiable
        training path lives in contextgs_b200.context_model."""
        if training is None:
            training = self.training
        x = x.contiguous()
        N, C = x.shape
        if noise is not None and (not noise.is_cuda or noise.shape != x.shape):
            raise ValueError("EntropyBottleneck: noise must be a CUDA tensor shaped like x")
        if training and noise is None:
            noise = torch.empty_like(x).uniform_(-0.5, 0.5)
        out, lik = torch.empty_like(x), torch.empty_like(x)
        noise = noise.contiguous() if training else None   # named: a copy made here must outlive the launch
        _lib.check(_lib.lib().cgs_eb_forward(
            _lib.ptr(self.packed()), C, _lib.ptr(x), _lib.ptr(noise), N,
            _lib.ptr(out), _lib.ptr(lik), _lib.ptr(choose), _lib.ptr(bit_sum), _lib.stream_ptr()), "cgs_eb_forward")
        return out, lik


def _compressai_keys(sd):
    """State-dict keys of CompressAI's EntropyBottleneck -> ours (same shapes): `_matrix3` -> `matrices.3`, ..."""
    out = {}
    for k, v in sd.items():
        for old, new in (("_matrix", "matrices."), ("_bias", "biases."), ("_factor", "factors.")):
            if k.startswith(old) and k[len(old):].isdigit():
                k = new + k[len(old):]
        out[k] = v
    return out


def _mlp(i, h, o, act=None):
    layers = [nn.Linear(i, h), nn.ReLU(True), nn.Linear(h, o)]
    if act is not None:
        layers.append(act)
    return nn.Sequential(*layers)


class GaussianModel(DensifyMixin, nn.Module):
    def __init__(self, feat_dim=50, n_offsets=10, voxel_size=0.001, update_depth=3, update_init_factor=16,
                 update_hierachy_factor=4, use_feat_bank=False, n_features_per_level=2, resolutions_list=None,
                 resolutions_list_2D=None, ste_binary=True, ste_multistep=False, add_noise=False, Q=1, use_2D=True,
                 decoded_version=False, level_num=3, adaptQ_per_channel=False, hyper_divisor=4, target_ratio=0.2,
                 disable_hyper=False, device="cuda"):
        """Same positional order and keywords as the reference's constructor (scene/gaussian_model.py:61-83), so the
        calls at train.py:94-107 / :450, test.py:149 and decompress.py:149 work unchanged; the defaults are the
        values ContextGS trains with (arguments/__init__.py:50-55,67-68; train.py:595-616), not the reference's
        signature defaults.  `device` is the one addition (the reference hard-codes .cuda())."""
        super().__init__()
        if (feat_dim, n_offsets, hyper_divisor, level_num) != (50, 10, 4, 3):
            raise NotImplementedError("the CUDA kernels are specialised for the ContextGS defaults "
                                      "feat_dim=50, n_offsets=10, hyper_divisor=4, level_num=3 "
                                      "(arguments/__init__.py:50-52,67-68; train.py:595)")
        if use_feat_bank or adaptQ_per_channel:
            raise NotImplementedError("use_feat_bank / adaptQ_per_channel are off in every ContextGS launch script "
                                      "(scripts/*.py, train_scripts/*.py) and are not part of the hot path")
        if target_ratio is None:
            target_ratio = 0.2
        self.feat_dim, self.n_offsets, self.voxel_size = feat_dim, n_offsets, voxel_size
        # densification constants (instance attributes like the reference; DensifyMixin's class values are the defaults)
        self.update_depth, self.update_init_factor = int(update_depth), int(update_init_factor)
        self.update_hierachy_factor, self.use_feat_bank = int(update_hierachy_factor), bool(use_feat_bank)
        self.n_features_per_level, self.ste_binary, self.ste_multistep = n_features_per_level, ste_binary, ste_multistep
        self.add_noise, self.Q, self.use_2D = add_noise, Q, use_2D
        self.resolutions_list, self.resolutions_list_2D = resolutions_list, resolutions_list_2D
        self.level_num, self.hyper_divisor, self.target_ratio = level_num, hyper_divisor, target_ratio
        self.decoded_version = decoded_version
        self.level_scale = None
        self.disable_hyper = bool(disable_hyper)
        self.adaptQ_per_channel = False
        self.x_bound_min = torch.zeros(1, 3, device=device)
        self.x_bound_max = torch.ones(1, 3, device=device)
        e = torch.empty(0, device=device)
        self._anchor = self._offset = self._mask = self._anchor_feat = self._hyper_latent = e
        self._scaling = self._rotation = self._opacity = e
        self.rotation_activation = torch.nn.functional.normalize
        self.latent_codec = EntropyBottleneck(feat_dim // hyper_divisor).to(device)
        d = feat_dim + 3 + 1
        self.mlp_opacity = _mlp(d, feat_dim, n_offsets, nn.Tanh()).to(device)
        self.mlp_cov = _mlp(d, feat_dim, 7 * n_offsets).to(device)
        self.mlp_color = _mlp(d, feat_dim, 3 * n_offsets, nn.Sigmoid()).to(device)
        self.mlp_grid = nn.ModuleList()
        out = (feat_dim + 6 + 3 * n_offsets) * 2 + 3
        H = feat_dim // hyper_divisor
        for i in range(level_num):
            fin = H + 3 if i == level_num - 1 else feat_dim + 6 + 3 + H
            self.mlp_grid.append(_mlp(fin, feat_dim * 2, out).to(device))

    # ---- construction from synthetic tensors (tests, bench) ---------------------------------
    @classmethod
    def from_tensors(cls, scene, mlps=None, eb=None, device="cuda", **kw):
        """scene: dict(anchor, feat, hyper, offset, scaling, mask, voxel_size) as produced by
        contextgs_b200.synthetic.make_scene; mlps/eb: optional weights in the oracle's layout
        (lists [W1,b1,W2,b2]; EntropyBottleneckRef) so that both sides share identical parameters."""
        m = cls(voxel_size=scene["voxel_size"], device=device, **kw)
        P = lambda t, g=True: nn.Parameter(t.detach().clone().float().to(device).contiguous(), requires_grad=g)
        m._anchor, m._anchor_feat, m._hyper_latent = P(scene["anchor"]), P(scene["feat"]), P(scene["hyper"])
        m._offset, m._mask, m._scaling = P(scene["offset"]), P(scene["mask"]), P(scene["scaling"])
        N = m._anchor.shape[0]
        rot = torch.zeros(N, 4)
        rot[:, 0] = 1
        m._rotation = P(rot, False)
        m._opacity = P(torch.zeros(N, 1), False)
        if mlps is not None:
            with torch.no_grad():
                for name in ("opacity", "cov", "color"):
                    seq = getattr(m, "mlp_" + name)
                    W1, b1, W2, b2 = mlps[name]
                    seq[0].weight.copy_(W1); seq[0].bias.copy_(b1); seq[2].weight.copy_(W2); seq[2].bias.copy_(b2)
                for i, (W1, b1, W2, b2) in enumerate(mlps["grid"]):
                    seq = m.mlp_grid[i]
                    seq[0].weight.copy_(W1); seq[0].bias.copy_(b1); seq[2].weight.copy_(W2); seq[2].bias.copy_(b2)
        if eb is not None:
            with torch.no_grad():
                for i, t in enumerate(eb.matrices):
                    m.latent_codec.matrices[i].copy_(t)
                for i, t in enumerate(eb.biases):
                    m.latent_codec.biases[i].copy_(t)
                for i, t in enumerate(eb.factors):
                    m.latent_codec.factors[i].copy_(t)
                m.latent_codec.quantiles.copy_(eb.quantiles)
        m.update_anchor_bound()
        return m

    # ---- getters (scene/gaussian_model.py:288-345) -------------------------------------------
    @property
    def get_scaling(self):
        if self.decoded_version:
            return self._scaling
        return 1.0 * torch.exp(self._scaling)

    @property
    def get_mask(self):
        if self.decoded_version:
            return self._mask
        if torch.is_grad_enabled() and self._mask.requires_grad:   # training: autograd needs the expression (STE)
            mask_sig = torch.sigmoid(self._mask)
            return ((mask_sig > 0.01).float() - mask_sig).detach() + mask_sig
        return self._binary_mask_state()[2]

    def _binary_mask_state(self):
        """(parameter, version, mask [N,K,1], anchor-valid bool [N], [all valid: bool or None]) of the no-grad evaluation
        of scene/gaussian_model.py:295-310, shared by `get_mask`, `get_mask_anchor` and the callers that need to know whether
        every anchor is valid (scoring pass, encoder).  The reference re-evaluates the expression on every access (SURVEY.md
        G2); here it is evaluated once per version of `_mask`.  The entry holds the parameter object itself, so "same
        object, same version" cannot be confused by a new tensor that reuses the address."""
        ent = self.__dict__.get("_cgs_mask_state")
        if (ent is None or ent[0] is not self._mask or ent[1] != self._mask._version
                or ent[5] != self._mask.data_ptr()):   # (`p.data = t` swaps the storage without a version bump)
            with torch.no_grad():
                mask_sig = torch.sigmoid(self._mask)
                mask = ((mask_sig > 0.01).float() - mask_sig) + mask_sig
                valid = torch.sum(mask, dim=1)[:, 0] > 0
            ent = [self._mask, self._mask._version, mask, valid, None, self._mask.data_ptr()]
            self.__dict__["_cgs_mask_state"] = ent
        return ent

    def all_anchors_valid(self):
        """bool(get_mask_anchor.all()) with the host read-back done once per version of `_mask`."""
        if self.decoded_version:
            return bool(self.get_mask_anchor.all())
        ent = self._binary_mask_state()
        if ent[4] is None:
            ent[4] = bool(ent[3].all())
        return ent[4]

    @property
    def get_mask_anchor(self):
        with torch.no_grad():
            if self.decoded_version:
                return (torch.sum(self._mask, dim=1)[:, 0]) > 0
            return self._binary_mask_state()[3]

    @property
    def get_opacity_mlp(self):
        return self.mlp_opacity

    @property
    def get_cov_mlp(self):
        return self.mlp_cov

    @property
    def get_color_mlp(self):
        return self.mlp_color

    @property
    def get_grid_mlp(self):
        return self.mlp_grid

    @property
    def get_rotation(self):
        return self.rotation_activation(self._rotation)

    @property
    def get_anchor(self):
        if self.decoded_version:
            return self._anchor
        anchor, _ = Quantize_anchor.apply(self._anchor, self.x_bound_min, self.x_bound_max)
        return anchor

    @torch.no_grad()
    def update_anchor_bound(self):
        """scene/gaussian_model.py:352-360."""
        mn = torch.min(self._anchor, dim=0, keepdim=True)[0].detach()
        mx = torch.max(self._anchor, dim=0, keepdim=True)[0].detach()
        self.x_bound_min = torch.where(mn < 0, mn * 1.2, mn * 0.8)
        self.x_bound_max = torch.where(mx > 0, mx * 1.2, mx * 0.8)
        self._cgs_bound_version = getattr(self, "_cgs_bound_version", 0) + 1   # invalidates the cached level plan

    @torch.no_grad()
    def replace_with_decoded(self, anchor, hyper, feat, offsets, scaling, masks):
        """Parameter replacement at the end of `conduct_decoding` (scene/gaussian_model.py:1503-1533):
        the model afterwards holds DECODED values (quantised anchors, dequantised feat / offsets,
        scaling already exponentiated, binary masks) and `decoded_version` is True, so rendering
        skips the context model (gaussian_renderer/__init__.py:103-104) -- the published-FPS path."""
        P = lambda t: nn.Parameter(t.detach().clone().float().contiguous())
        self._hyper_latent, self._anchor_feat, self._offset = P(hyper), P(feat), P(offsets)
        self.decoded_version = True
        self._anchor, self._scaling, self._mask = P(anchor), P(scaling), P(masks)
        if self._rotation.shape[0] != self._anchor.shape[0]:   # identity rotations, as conduct_decoding creates them
            rot = torch.zeros((self._anchor.shape[0], 4), device=self._anchor.device)
            rot[:, 0] = 1
            self._rotation = nn.Parameter(rot, requires_grad=False)
        if hasattr(self, "_cgs_level_plan"):
            del self._cgs_level_plan
        return self

    # ---- densification statistics (scene/gaussian_model.py:429-433, 696-713) ---------------------------------
    def _ensure_statis(self):
        N, K, dev = self._anchor.shape[0], self.n_offsets, self._anchor.device
        if getattr(self, "opacity_accum", None) is None or self.opacity_accum.shape[0] != N:
            self.opacity_accum = torch.zeros((N, 1), device=dev)
            self.anchor_demon = torch.zeros((N, 1), device=dev)
            self.offset_gradient_accum = torch.zeros((N * K, 1), device=dev)
            self.offset_denom = torch.zeros((N * K, 1), device=dev)

    @torch.no_grad()
    def training_statis(self, viewspace_point_tensor, opacity, update_filter, offset_selection_mask, anchor_visible_mask):
        """Same signature and effect as scene/gaussian_model.py:696-713, one library call, no host synchronisation
        (the reference synchronises on every boolean-index assignment)."""
        self._ensure_statis()
        L = _lib.lib()
        N, K = self._anchor.shape[0], self.n_offsets
        grad = viewspace_point_tensor.grad
        P = int(update_filter.shape[0])
        if P and (grad is None or grad.shape[0] != P):
            raise ValueError("training_statis: viewspace_point_tensor.grad must hold one row per emitted Gaussian")
        u8 = lambda t: t.contiguous().view(torch.uint8) if t.dtype == torch.bool else t.contiguous().to(torch.uint8)
        vis, keep, upd = u8(anchor_visible_mask), u8(offset_selection_mask.reshape(-1)), u8(update_filter)
        op = opacity.detach().reshape(-1).float().contiguous()
        if op.numel() != keep.numel():
            raise ValueError("training_statis: opacity and offset_selection_mask must cover the same (visible anchor, offset) slots")
        ws = torch.empty((L.cgs_training_statis_workspace_bytes(N, K),), dtype=torch.uint8, device=vis.device)
        g = grad.detach().float().contiguous() if P else None
        _lib.check(L.cgs_training_statis(N, K, _lib.ptr(vis), _lib.ptr(keep), keep.numel(), _lib.ptr(op), _lib.ptr(g),
                                         _lib.ptr(upd) if P else None, P, _lib.ptr(self.opacity_accum),
                                         _lib.ptr(self.anchor_demon), _lib.ptr(self.offset_gradient_accum),
                                         _lib.ptr(self.offset_denom), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "cgs_training_statis")

    def save_ply(self, path):
        """scene/gaussian_model.py:578-597 (numpy writer, same 119-column vertex schema)."""
        from .ply_io import save_ply
        save_ply(self, path)

    def load_ply_sparse_gaussian(self, path):
        """scene/gaussian_model.py:599-656."""
        from .ply_io import load_ply_sparse_gaussian
        load_ply_sparse_gaussian(self, path)

    def conduct_encoding(self, pre_path_name):
        """scene/gaussian_model.py:1005-1300 on the GPU codec (contextgs_b200/codec.py); returns the size summary."""
        from .codec import conduct_encoding
        return conduct_encoding(self, pre_path_name)[1]

    def conduct_decoding(self, pre_path_name):
        """scene/gaussian_model.py:1302-1538."""
        from .codec import conduct_decoding
        conduct_decoding(self, pre_path_name)
        return ""

    # ---- initialisation from an SfM point cloud (scene/gaussian_model.py:377-423) -----------------------------------
    def voxelize_sample(self, data=None, voxel_size=0.01):
        """:377-380 (numpy; the shuffle does not change the sorted unique result but keeps the reference's RNG use)."""
        import numpy as np
        np.random.shuffle(data)
        return np.unique(np.round(data / voxel_size), axis=0) * voxel_size

    def create_from_pcd(self, pcd, spatial_lr_scale):
        """:382-423: anchors = voxelised points; scales from the mean squared distance to the 3 nearest anchors."""
        import numpy as np
        from .knn import distCUDA2
        dev = self.mlp_opacity[0].weight.device
        self.spatial_lr_scale = spatial_lr_scale
        points = np.asarray(pcd.points if hasattr(pcd, "points") else pcd)
        if self.voxel_size <= 0:
            init_dist = distCUDA2(torch.tensor(points).float().to(dev))
            self.voxel_size = torch.kthvalue(init_dist, int(init_dist.shape[0] * 0.5))[0].item()
        points = self.voxelize_sample(points, voxel_size=self.voxel_size)
        fused = torch.tensor(np.asarray(points)).float().to(dev)
        n, K = fused.shape[0], self.n_offsets
        dist2 = torch.clamp_min(distCUDA2(fused), 0.0000001)
        scales = torch.log(torch.sqrt(dist2))[..., None].repeat(1, 6)
        rots = torch.zeros((n, 4), device=dev)
        rots[:, 0] = 1
        opacities = torch.log(torch.full((n, 1), 0.1, device=dev) / (1 - torch.full((n, 1), 0.1, device=dev)))
        P = lambda t, g=True: nn.Parameter(t.contiguous().requires_grad_(g))
        self._anchor, self._offset = P(fused), P(torch.zeros((n, K, 3), device=dev))
        self._mask = P(torch.ones((n, K, 1), device=dev))
        self._anchor_feat = P(torch.zeros((n, self.feat_dim), device=dev))
        self._hyper_latent = P(torch.zeros((n, self.feat_dim // self.hyper_divisor), device=dev))
        self._scaling, self._rotation, self._opacity = P(scales), P(rots, False), P(opacities, False)
        self.max_radii2D = torch.zeros((n,), device=dev)
        self.update_anchor_bound()

    # ---- checkpoints (scene/gaussian_model.py:221-286 capture / restore, :912-951 MLP checkpoint) ------------------
    def capture(self):
        """The same 19-tuple as scene/gaussian_model.py:221-249, so `torch.save((gaussians.capture(), iteration), ...)`
        (train.py:278) writes the reference's checkpoint layout."""
        return (self._anchor, self._anchor_feat, self._hyper_latent, self._offset, self._mask, self._scaling, self._rotation,
                self._opacity, getattr(self, "max_radii2D", None), self.optimizer.state_dict(), self.spatial_lr_scale,
                self.mlp_opacity.state_dict(), self.mlp_cov.state_dict(), self.mlp_color.state_dict(),
                self.latent_codec.state_dict(), self.mlp_grid.state_dict(), self.x_bound_min, self.x_bound_max,
                self.level_scale)

    def restore(self, model_args, training_args):
        """scene/gaussian_model.py:251-285."""
        (anchor, feat, hyper, offset, mask, scaling, rotation, opacity, self.max_radii2D, opt_dict, self.spatial_lr_scale,
         sd_opacity, sd_cov, sd_color, sd_codec, sd_grid, self.x_bound_min, self.x_bound_max, self.level_scale) = model_args
        dev = self.mlp_opacity[0].weight.device
        P = lambda t: t if isinstance(t, nn.Parameter) and t.device == dev else nn.Parameter(
            t.detach().to(dev).float().contiguous(), requires_grad=bool(t.requires_grad))
        self._anchor, self._anchor_feat, self._hyper_latent, self._offset = P(anchor), P(feat), P(hyper), P(offset)
        self._mask, self._scaling, self._rotation, self._opacity = P(mask), P(scaling), P(rotation), P(opacity)
        self.x_bound_min, self.x_bound_max = self.x_bound_min.to(dev), self.x_bound_max.to(dev)
        self.training_setup(training_args)
        self.optimizer.load_state_dict(opt_dict)
        self.mlp_opacity.load_state_dict(sd_opacity)
        self.mlp_cov.load_state_dict(sd_cov)
        self.mlp_color.load_state_dict(sd_color)
        self.latent_codec.load_state_dict(_compressai_keys(sd_codec), strict=False)
        self.mlp_grid.load_state_dict(sd_grid)

    def save_mlp_checkpoints(self, path):
        """scene/gaussian_model.py:912-936: same file keys."""
        import os
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save({"opacity_mlp": self.mlp_opacity.state_dict(), "cov_mlp": self.mlp_cov.state_dict(),
                    "color_mlp": self.mlp_color.state_dict(), "latent_codec": self.latent_codec.state_dict(),
                    "grid_mlp": self.mlp_grid.state_dict(), "bound": [self.x_bound_min, self.x_bound_max],
                    "level_scale": self.level_scale}, path)

    def load_mlp_checkpoints(self, path):
        """scene/gaussian_model.py:939-951.  Accepts the reference's files: CompressAI's buffers (`_offset`,
        `_quantized_cdf`, `_cdf_length`, `target`) are ignored (strict=False there too) and both of CompressAI's
        parameter namings (`_matrix0` ... of 1.1.x, `matrices.0` ... of >= 1.2) are understood."""
        dev = self.mlp_opacity[0].weight.device
        ck = torch.load(path, map_location=dev, weights_only=False)
        self.mlp_opacity.load_state_dict(ck["opacity_mlp"])
        self.mlp_cov.load_state_dict(ck["cov_mlp"])
        self.mlp_color.load_state_dict(ck["color_mlp"])
        self.latent_codec.load_state_dict(_compressai_keys(ck["latent_codec"]), strict=False)
        self.mlp_grid.load_state_dict(ck["grid_mlp"])
        self.x_bound_min, self.x_bound_max = (t.to(dev) for t in ck["bound"])
        self.level_scale = ck["level_scale"]

    def eval(self):
        for m in (self.mlp_opacity, self.mlp_cov, self.mlp_color, self.latent_codec, self.mlp_grid):
            m.eval()
        return self

    def train(self, mode=True):
        for m in (self.mlp_opacity, self.mlp_cov, self.mlp_color, self.latent_codec, self.mlp_grid):
            m.train(mode)
        return self

    @torch.no_grad()
    def estimate_final_bits(self, return_values=False):
        """scene/gaussian_model.py:981-1004 (without the side-effect files at :1681-1682)."""
        from .context_model import multi_scale_generating
        sel = self.get_mask_anchor
        tensors = (self.get_anchor, self._hyper_latent, self._anchor_feat, self._offset, self.get_scaling, self.get_mask)
        if not self.all_anchors_valid():  # one index list for the six gathers (the reference runs six boolean-mask selects)
            idx = torch.nonzero(sel)[:, 0]
            tensors = tuple(t.index_select(0, idx) for t in tensors)
        a, h, f, o, s, m = tensors
        sums = multi_scale_generating(self, a, h, f, o, s, binary_grid_masks=m, predict_bpp=True, return_sum_bits=True)
        if return_values:
            return sums
        names = ["anchor", "hyper", "feat", "scaling", "offsets", "masks"]
        mb = 8 * 1024 * 1024
        return "\nEstimated sizes in MB: " + ", ".join(f"{n} {round(v / mb, 4)}" for n, v in zip(names, sums))

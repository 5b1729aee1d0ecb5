"""TEST INFRASTRUCTURE.  Loads the reference's own modules from the code objects oracle/build_ref.py produced
(oracle/_ref/*.refbin) and wires their third-party imports:

  diff_gaussian_rasterization -> contextgs_b200.dropin.diff_gaussian_rasterization   (the thing under test)
  simple_knn._C               -> contextgs_b200.dropin.simple_knn._C
  compressai.EntropyBottleneck-> EntropyBottleneckTorch below (CompressAI's forward restated with torch ops;
                                 third-party and absent: PARITY UNPINNED, as everywhere else)
  torchac, plyfile, torch_scatter, compressai.latent_codecs -> empty stand-ins (never called on the hot path)

`load()` returns a namespace with the modules; nothing here is imported by the product."""
import importlib.util
import marshal
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
ORDER = ["utils.general_utils", "utils.graphics_utils", "utils.system_utils", "utils.encodings", "utils.entropy_models",
         "utils.multi_level", "utils.loss_utils", "scene.gaussian_model", "gaussian_renderer"]


def available():
    return all(os.path.exists(os.path.join(OUT, n + ".refbin")) for n in ORDER)


class EntropyBottleneckTorch(nn.Module):
    """compressai.entropy_models.EntropyBottleneck.forward restated (filters (3,3,3,3), init_scale 10, likelihood
    floor 1e-9), as an nn.Module so that the reference's `.cuda()` / state_dict calls work.  Same parameter
    shapes as oracle.entropy_ref.EntropyBottleneckRef."""

    def __init__(self, channels, *a, filters=(3, 3, 3, 3), init_scale=10.0, **kw):
        super().__init__()
        self.channels, self.filters = int(channels), tuple(filters)
        f = (1,) + self.filters + (1,)
        scale = init_scale ** (1 / (len(self.filters) + 1))
        self.matrices, self.biases, self.factors = nn.ParameterList(), nn.ParameterList(), nn.ParameterList()
        for i in range(len(self.filters) + 1):
            init = float(np.log(np.expm1(1 / scale / f[i + 1])))
            self.matrices.append(nn.Parameter(torch.full((channels, f[i + 1], f[i]), init)))
            self.biases.append(nn.Parameter(torch.rand(channels, f[i + 1], 1) - 0.5))
            if i < len(self.filters):
                self.factors.append(nn.Parameter(torch.zeros(channels, f[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.tensor([-init_scale, 0.0, init_scale]).repeat(channels, 1, 1))
        self.likelihood_bound = 1e-9

    def load_ref(self, eb):
        with torch.no_grad():
            for dst, src in ((self.matrices, eb.matrices), (self.biases, eb.biases), (self.factors, eb.factors)):
                for d, s in zip(dst, src):
                    d.copy_(s)
            self.quantiles.copy_(eb.quantiles)
        return self

    def _logits(self, v):
        for i in range(len(self.filters) + 1):
            v = torch.matmul(F.softplus(self.matrices[i]), v) + self.biases[i]
            if i < len(self.filters):
                v = v + torch.tanh(self.factors[i]) * torch.tanh(v)
        return v

    def forward(self, x, training=None):
        training = self.training if training is None else training
        v = x.permute(1, 0).contiguous().reshape(self.channels, 1, -1)
        if training:
            out = v + torch.empty_like(v).uniform_(-0.5, 0.5)
        else:
            med = self.quantiles[:, :, 1:2].detach()
            out = torch.round(v - med) + med
        lower, upper = self._logits(out - 0.5), self._logits(out + 0.5)
        sign = -torch.sign(lower + upper).detach()
        lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower)).clamp(min=self.likelihood_bound)
        N = x.shape[0]
        return out.reshape(self.channels, N).permute(1, 0), lik.reshape(self.channels, N).permute(1, 0)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _not_on_path(what):
    def f(*a, **k):
        raise RuntimeError(f"{what} is not part of the hot path (stand-in)")
    return f


_loaded = None


def load():
    """Install the stand-ins and the reference modules into sys.modules (idempotent)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise FileNotFoundError("oracle/_ref/*.refbin missing: run `python -m oracle.build_ref` where /root/reference exists")
    _stub("torchac")
    _stub("plyfile", PlyData=object, PlyElement=object)
    _stub("torch_scatter", scatter_max=_not_on_path("torch_scatter.scatter_max"))
    ca = _stub("compressai")
    ca.entropy_models = _stub("compressai.entropy_models", EntropyBottleneck=EntropyBottleneckTorch, GaussianConditional=object)
    ca.latent_codecs = _stub("compressai.latent_codecs", LatentCodec=object, HyperLatentCodec=object)
    import contextgs_b200.dropin.diff_gaussian_rasterization as dgr
    sys.modules["diff_gaussian_rasterization"] = dgr
    try:
        import contextgs_b200.dropin.simple_knn as sk
        import contextgs_b200.dropin.simple_knn._C as skc
        sys.modules["simple_knn"], sys.modules["simple_knn._C"] = sk, skc
    except Exception:
        sk = _stub("simple_knn")
        sk._C = _stub("simple_knn._C", distCUDA2=_not_on_path("simple_knn.distCUDA2"))
    for pkg in ("utils", "scene"):
        if pkg not in sys.modules or not getattr(sys.modules[pkg], "_cgs_ref_pkg", False):
            m = _stub(pkg)
            m.__path__ = []
            m._cgs_ref_pkg = True
    mods = {}
    for name in ORDER:
        with open(os.path.join(OUT, name + ".refbin"), "rb") as f:
            f.read(16)                      # magic, flags, mtime, size (PEP 552 header)
            code = marshal.load(f)
        m = types.ModuleType(name)
        m.__file__ = f"reference:{name}"
        if name == "gaussian_renderer":
            m.__path__ = []
        sys.modules[name] = m
        exec(code, m.__dict__)
        if "." in name:
            setattr(sys.modules[name.split(".")[0]], name.split(".")[1], m)
        mods[name] = m
    _loaded = types.SimpleNamespace(gaussian_renderer=mods["gaussian_renderer"], gaussian_model=mods["scene.gaussian_model"],
                                    encodings=mods["utils.encodings"], entropy_models=mods["utils.entropy_models"],
                                    multi_level=mods["utils.multi_level"], loss_utils=mods["utils.loss_utils"],
                                    EntropyBottleneck=EntropyBottleneckTorch)
    return _loaded


def reference_model(scene, mlps=None, eb=None):
    """The reference's OWN GaussianModel (scene/gaussian_model.py:46-190; hard-codes .cuda()) filled with a synthetic
    scene and, optionally, the oracle's MLP / bottleneck weights (lists [W1, b1, W2, b2]; EntropyBottleneckRef)."""
    ref = load()
    m = ref.gaussian_model.GaussianModel(50, 10, scene["voxel_size"], 3, 16, 4, False, n_features_per_level=2, level_num=3,
                                         hyper_divisor=4, target_ratio=0.2, disable_hyper=False)
    P = lambda t, g=True: nn.Parameter(t.detach().clone().float().cuda().contiguous(), requires_grad=g)
    m._anchor, m._anchor_feat, m._hyper_latent = P(scene["anchor"]), P(scene["feat"]), P(scene["hyper"])
    m._offset, m._mask, m._scaling = P(scene["offset"]), P(scene["mask"]), P(scene["scaling"])
    N = m._anchor.shape[0]
    rot = torch.zeros(N, 4)
    rot[:, 0] = 1
    m._rotation, m._opacity = P(rot, False), P(torch.zeros(N, 1), False)
    if mlps is not None:
        with torch.no_grad():
            for name, seq in (("opacity", m.mlp_opacity), ("cov", m.mlp_cov), ("color", m.mlp_color)):
                W1, b1, W2, b2 = mlps[name]
                seq[0].weight.copy_(W1); seq[0].bias.copy_(b1); seq[2].weight.copy_(W2); seq[2].bias.copy_(b2)
            for i, (W1, b1, W2, b2) in enumerate(mlps["grid"]):
                seq = m.mlp_grid[i]
                seq[0].weight.copy_(W1); seq[0].bias.copy_(b1); seq[2].weight.copy_(W2); seq[2].bias.copy_(b2)
    if eb is not None:
        m.latent_codec.load_ref(eb)
    m.update_anchor_bound()
    return m

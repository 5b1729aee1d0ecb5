"""`point_cloud.ply` of a ContextGS model: the 119-column vertex schema of
`GaussianModel.construct_list_of_attributes` / `save_ply` / `load_ply_sparse_gaussian`
(scene/gaussian_model.py:561-656), read and written with numpy alone (the reference goes through `plyfile`, which is
not a dependency here): binary_little_endian 1.0, one `vertex` element, every property a 32-bit float, in the order
x y z | nx ny nz | f_offset_0..29 | f_mask_0..9 | f_anchor_feat_0..49 | f_hyper_latent_0..11 | opacity | scale_0..5 | rot_0..3.
Offsets are stored channel-major (`_offset.transpose(1, 2).flatten(1)`: f_offset_{c*K + k}), masks as [N, 1, K] flattened.
Plain file IO on the host: not part of the GPU hot path."""
import os

import numpy as np
import torch
import torch.nn as nn


def attribute_names(n_offsets=10, feat_dim=50, hyper_dim=12, scale_dim=6, rot_dim=4):
    """scene/gaussian_model.py:561-576."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_offset_{i}" for i in range(3 * n_offsets)]
    names += [f"f_mask_{i}" for i in range(n_offsets)]
    names += [f"f_anchor_feat_{i}" for i in range(feat_dim)]
    names += [f"f_hyper_latent_{i}" for i in range(hyper_dim)]
    names += ["opacity"]
    names += [f"scale_{i}" for i in range(scale_dim)]
    names += [f"rot_{i}" for i in range(rot_dim)]
    return names


def write_ply(path, columns, names):
    """columns: float32 [N, len(names)]."""
    columns = np.ascontiguousarray(columns, dtype="<f4")
    if columns.ndim != 2 or columns.shape[1] != len(names):
        raise ValueError("write_ply: one column per property expected")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {columns.shape[0]}"]
    header += [f"property float {n}" for n in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(columns.tobytes())


_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1", "char": "i1",
              "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2", "int": "<i4", "int32": "<i4",
              "uint": "<u4", "uint32": "<u4"}


def read_ply(path):
    """-> dict name -> float32 array [N] of the first element of a binary_little_endian PLY (scalar properties only)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_first, elements = None, None, [], False, 0
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements += 1
                in_first = elements == 1
                if in_first:
                    count = int(tok[2])
            elif tok[0] == "property" and in_first:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not part of the ContextGS schema")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian":
            raise ValueError(f"{path}: only binary_little_endian PLY files are supported (got {fmt})")
        data = np.frombuffer(f.read(count * np.dtype(props).itemsize), dtype=np.dtype(props), count=count)
    return {name: np.asarray(data[name], dtype=np.float32) for name, _ in props}


def save_ply(pc, path):
    """scene/gaussian_model.py:578-597."""
    t = lambda x: x.detach().float().cpu().numpy()
    N = pc._anchor.shape[0]
    opacity = t(pc._opacity) if pc._opacity.numel() == N else np.zeros((N, 1), np.float32)
    cols = np.concatenate([t(pc._anchor), np.zeros((N, 3), np.float32),
                           t(pc._offset.transpose(1, 2).flatten(start_dim=1)), t(pc._mask.transpose(1, 2).flatten(start_dim=1)),
                           t(pc._anchor_feat), t(pc._hyper_latent), opacity, t(pc._scaling), t(pc._rotation)], axis=1)
    names = attribute_names(pc._offset.shape[1], pc._anchor_feat.shape[1], pc._hyper_latent.shape[1], pc._scaling.shape[1],
                            pc._rotation.shape[1])
    write_ply(path, cols, names)


def load_ply_sparse_gaussian(pc, path, device=None):
    """scene/gaussian_model.py:599-656: fills the model's per-anchor parameters from a `point_cloud.ply`."""
    d = read_ply(path)
    device = device if device is not None else pc.latent_codec.quantiles.device

    def group(prefix):
        names = sorted([n for n in d if n.startswith(prefix)], key=lambda n: int(n.split("_")[-1]))
        return np.stack([d[n] for n in names], axis=1) if names else np.zeros((len(d["x"]), 0), np.float32)

    P = lambda a, g=True: nn.Parameter(torch.tensor(a, dtype=torch.float, device=device).contiguous(), requires_grad=g)
    anchor = np.stack([d["x"], d["y"], d["z"]], axis=1)
    offsets = group("f_offset").reshape(anchor.shape[0], 3, -1)
    masks = group("f_mask").reshape(anchor.shape[0], 1, -1)
    pc._anchor_feat, pc._hyper_latent = P(group("f_anchor_feat")), P(group("f_hyper_latent"))
    pc._offset = nn.Parameter(torch.tensor(offsets, dtype=torch.float, device=device).transpose(1, 2).contiguous())
    pc._mask = nn.Parameter(torch.tensor(masks, dtype=torch.float, device=device).transpose(1, 2).contiguous())
    pc._anchor, pc._opacity = P(anchor), P(d["opacity"][:, None], False)
    pc._scaling, pc._rotation = P(group("scale_")), P(group("rot"), False)
    if hasattr(pc, "_cgs_level_plan"):
        del pc._cgs_level_plan
    return pc

"""point_cloud.ply of the reference's schema (scene/gaussian_model.py:561-656): header, column order, round trip."""
import numpy as np
import torch

from contextgs_b200 import ply_io, synthetic
from contextgs_b200.gaussian_model import GaussianModel


def test_schema_matches_the_reference_attribute_list():
    names = ply_io.attribute_names()
    assert len(names) == 119 and names[:6] == ["x", "y", "z", "nx", "ny", "nz"]
    assert names[6] == "f_offset_0" and names[36] == "f_mask_0" and names[46] == "f_anchor_feat_0"
    assert names[96] == "f_hyper_latent_0" and names[108] == "opacity" and names[109] == "scale_0" and names[115] == "rot_0"


def test_save_load_round_trip(tmp_path):
    scene = synthetic.make_scene("chair", 300, seed=4)
    torch.manual_seed(0)
    pc = GaussianModel.from_tensors(scene, device="cpu")
    path = tmp_path / "point_cloud" / "iteration_30000" / "point_cloud.ply"
    pc.save_ply(str(path))
    raw = open(path, "rb").read()
    head = raw[:raw.index(b"end_header\n") + 11].decode("ascii").split("\n")
    assert head[0] == "ply" and head[1] == "format binary_little_endian 1.0" and head[2] == "element vertex 300"
    assert head[3] == "property float x" and head[3 + 118] == "property float rot_3"
    assert len(raw) == raw.index(b"end_header\n") + 11 + 300 * 119 * 4
    d = ply_io.read_ply(str(path))
    # channel-major offsets: f_offset_{c*K + k} = _offset[:, k, c]
    assert np.array_equal(d["f_offset_13"], pc._offset[:, 3, 1].detach().numpy())
    assert np.array_equal(d["f_mask_7"], pc._mask[:, 7, 0].detach().numpy())
    other = GaussianModel(device="cpu")
    other.load_ply_sparse_gaussian(str(path))
    for k in ("_anchor", "_anchor_feat", "_hyper_latent", "_offset", "_mask", "_scaling", "_rotation"):
        assert torch.equal(getattr(other, k).detach(), getattr(pc, k).detach()), k
    assert other._offset.shape == (300, 10, 3) and other._mask.shape == (300, 10, 1)

"""TEST INFRASTRUCTURE -- CPU restatement of the proposal density field (HashMLPDensityField,
nerfstudio/fields/density_fields.py:40-121): tcnn HashGrid -> ReLU MLP without biases -> trunc_exp (= exp in the forward,
field_components/activations.py:24-42).  tiny-cuda-nn is not vendored: PARITY UNPINNED (restates its published semantics)."""
import torch

from . import hashgrid
from .field import scene_contraction


def density_field(positions, weights, table, hidden, n_hidden, n_levels, n_features, log2_hashmap_size, base_res, per_level_scale, aabb=None,
                  contraction=None):
    x = positions.reshape(-1, 3)
    if aabb is not None:
        x01 = (x - aabb[0]) / (aabb[1] - aabb[0])
    else:
        x01 = (scene_contraction(x, contraction) + 2.0) / 4.0
    meta = hashgrid.tcnn_grid_meta(n_levels, n_features, log2_hashmap_size, base_res, per_level_scale)
    feat = hashgrid.encode_tcnn_layout(x01, table.view(-1, n_features), meta, n_features, False)
    in_dim = n_levels * n_features
    in_pad = (in_dim + 15) // 16 * 16
    o = 0
    W0 = weights[o: o + hidden * in_pad].view(hidden, in_pad)[:, :in_dim]
    o += hidden * in_pad
    h = torch.relu(feat @ W0.t())
    for _ in range(n_hidden - 1):
        W = weights[o: o + hidden * hidden].view(hidden, hidden)
        o += hidden * hidden
        h = torch.relu(h @ W.t())
    out = h @ weights[o: o + hidden]
    return torch.exp(out).view(*positions.shape[:-1], 1), out.view(*positions.shape[:-1], 1)

"""TEST INFRASTRUCTURE -- CPU restatement of ``SDFField``.

Follows nerfstudio/fields/sdf_field.py:380-698, field_components/encodings.py:167-208 (NeRFEncoding),
field_components/spatial_distortions.py:66-73 (SceneContraction).  Functional style on plain tensors; parameters
are taken from a dict whose keys equal the reference ``state_dict`` names (``glin0.weight_g`` ...), so a reference module's
weights can be loaded verbatim for pinning.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import hashgrid

# 21 icosahedron directions of mip-NeRF-360's off-axis encoding (encodings.py:129-153), stored [3, 21].
OFF_AXIS_P = torch.tensor(
    [
        [0.8506508, 0, 0.5257311], [0.809017, 0.5, 0.309017], [0.5257311, 0.8506508, 0], [1, 0, 0],
        [0.809017, 0.5, -0.309017], [0.8506508, 0, -0.5257311], [0.309017, 0.809017, -0.5],
        [0, 0.5257311, -0.8506508], [0.5, 0.309017, -0.809017], [0, 1, 0], [-0.5257311, 0.8506508, 0],
        [-0.309017, 0.809017, -0.5], [0, 0.5257311, 0.8506508], [-0.309017, 0.809017, 0.5],
        [0.309017, 0.809017, 0.5], [0.5, 0.309017, 0.809017], [0.5, -0.309017, 0.809017], [0, 0, 1],
        [-0.5, 0.309017, 0.809017], [-0.809017, 0.5, 0.309017], [-0.809017, 0.5, -0.309017],
    ]
).T.contiguous()  # fmt: skip


@dataclass
class FieldSpec:
    """The subset of ``SDFFieldConfig`` (sdf_field.py:121-185) that changes the arithmetic."""

    num_layers: int = 8
    hidden_dim: int = 256
    geo_feat_dim: int = 256
    num_layers_color: int = 4
    hidden_dim_color: int = 256
    appearance_embedding_dim: int = 32
    use_appearance_embedding: bool = False
    use_grid_feature: bool = False
    position_encoding_max_degree: int = 6
    use_position_encoding: bool = True
    off_axis: bool = False
    use_numerical_gradients: bool = False
    use_diffuse_color: bool = False
    use_specular_tint: bool = False
    use_reflections: bool = False
    use_n_dot_v: bool = False
    rgb_padding: float = 0.001
    num_levels: int = 16
    max_res: int = 2048
    base_res: int = 16
    log2_hashmap_size: int = 19
    hash_features_per_level: int = 2
    hash_smoothstep: bool = True
    weight_norm: bool = True
    skip_in: List[int] = field(default_factory=lambda: [4])
    grid_layout: str = "torch"  # "torch" (reference HashEncoding) | "tcnn"
    contraction: Optional[str] = None  # None | "linf" | "l2"  (SceneContraction order)

    @property
    def grid_out_dim(self):
        return self.num_levels * self.hash_features_per_level

    @property
    def pe_dim(self):
        per = 21 if self.off_axis else 3
        return per * self.position_encoding_max_degree * 2

    @property
    def geo_in_dim(self):
        return 3 + self.pe_dim + self.grid_out_dim

    @property
    def dir_enc_dim(self):
        return 3 * 4 * 2 + 3

    @property
    def color_in_dim(self):
        if self.use_diffuse_color:
            d = self.dir_enc_dim + self.geo_feat_dim + self.appearance_embedding_dim
        else:
            d = 3 + self.dir_enc_dim + 3 + self.geo_feat_dim + self.appearance_embedding_dim
        return d + (1 if self.use_n_dot_v else 0)

    def geo_dims(self):
        return [self.geo_in_dim] + [self.hidden_dim] * self.num_layers + [1 + self.geo_feat_dim]

    def color_dims(self):
        return [self.color_in_dim] + [self.hidden_dim_color] * self.num_layers_color + [3]


def nerf_encoding(x, num_frequencies: int, min_exp: float, max_exp: float, include_input: bool, off_axis: bool = False):
    """encodings.py:167-208: sin(x*2^k) for all (dim,k), then sin(.+pi/2) for all (dim,k), then optionally x."""
    freqs = (2 ** torch.linspace(min_exp, max_exp, num_frequencies)).to(x.dtype)
    base = x @ OFF_AXIS_P.to(x.dtype) if off_axis else x
    scaled = (base[..., None] * freqs).reshape(*base.shape[:-1], -1)
    enc = torch.sin(torch.cat([scaled, scaled + torch.pi / 2.0], dim=-1))
    if include_input:
        enc = torch.cat([enc, x], dim=-1)
    return enc


def scene_contraction(x, order: Optional[str]):
    """spatial_distortions.py:66-73.  order None -> identity (no distortion module); 'linf' / 'l2'."""
    if order is None:
        return x
    mag = torch.linalg.norm(x, ord=float("inf") if order == "linf" else None, dim=-1)
    mask = mag >= 1
    out = x.clone()
    out[mask] = (2 - (1 / mag[mask][..., None])) * (x[mask] / mag[mask][..., None])
    return out


def softplus_beta100(z):
    """nn.Softplus(beta=100) (sdf_field.py:365) with PyTorch's threshold 20."""
    return torch.nn.functional.softplus(z, beta=100)


def folded_weight(params: Dict[str, torch.Tensor], name: str, weight_norm: bool):
    """nn.utils.weight_norm, dim=0 (sdf_field.py:312-313): W[o,:] = g[o] * v[o,:] / ||v[o,:]||."""
    if weight_norm and (name + ".weight_g") in params:
        v, g = params[name + ".weight_v"], params[name + ".weight_g"]
        return torch._weight_norm(v, g, 0)
    return params[name + ".weight"]


class OracleField:
    """CPU restatement of SDFField's arithmetic.  ``params`` keys follow the reference state_dict; the grid table is
    ``params['hash_table']`` ([L*T, F] torch layout, or flat [total, F] tcnn layout)."""

    def __init__(self, spec: FieldSpec, params: Dict[str, torch.Tensor], dtype=torch.float32):
        self.spec = spec
        self.dtype = dtype
        self.p = {k: (v.detach().to(dtype) if v.is_floating_point() else v.detach()) for k, v in params.items()}
        self.hash_mask = torch.ones(spec.grid_out_dim, dtype=dtype)
        self.cos_anneal_ratio = 1.0
        self.numerical_gradients_delta = 0.0001
        self.training = False
        self.use_average_appearance_embedding = False
        if spec.use_grid_feature:
            if spec.grid_layout == "torch":
                # scalings are computed in fp32 exactly like the reference, then cast (encodings.py:303)
                g = hashgrid.growth_factor(spec.num_levels, spec.base_res, spec.max_res)
                max_res = spec.base_res * g ** (spec.num_levels - 1)  # what the tcnn stand-in feeds HashEncoding
                self.scalings = hashgrid.torch_layout_scalings(spec.num_levels, spec.base_res, max_res)
            else:
                g = hashgrid.growth_factor(spec.num_levels, spec.base_res, spec.max_res)
                self.meta = hashgrid.tcnn_grid_meta(spec.num_levels, spec.hash_features_per_level, spec.log2_hashmap_size, spec.base_res, g)

    # -- sdf_field.py:376-378
    def update_mask(self, level: int):
        self.hash_mask[:] = 1.0
        self.hash_mask[level * self.spec.hash_features_per_level :] = 0

    def grid_features(self, positions01):
        s = self.spec
        if s.grid_layout == "torch":
            return hashgrid.encode_torch_layout(positions01, self.p["hash_table"], self.scalings, 1 << s.log2_hashmap_size, s.hash_smoothstep)
        return hashgrid.encode_tcnn_layout(positions01, self.p["hash_table"], self.meta, s.hash_features_per_level, s.hash_smoothstep)

    # -- sdf_field.py:380-410
    def forward_geonetwork(self, x):
        s = self.spec
        if s.use_grid_feature:
            feat = self.grid_features((x + 2.0) / 4.0) * self.hash_mask
        else:
            feat = torch.zeros(x.shape[0], s.grid_out_dim, dtype=x.dtype)
        pe = nerf_encoding(x, s.position_encoding_max_degree, 0.0, s.position_encoding_max_degree - 1, False, s.off_axis)
        if not s.use_position_encoding:
            pe = torch.zeros_like(pe)
        inputs = torch.cat((x, pe, feat), dim=-1)
        h = inputs
        n_lin = s.num_layers + 1
        for l in range(n_lin):
            if l in s.skip_in:
                h = torch.cat([h, inputs], 1) / np.sqrt(2)
            W = folded_weight(self.p, f"glin{l}", s.weight_norm)
            h = torch.nn.functional.linear(h, W, self.p[f"glin{l}.bias"])
            if l < n_lin - 1:
                h = softplus_beta100(h)
        return h

    # -- sdf_field.py:424-465
    def gradient(self, x, skip_spatial_distortion=False, return_sdf=False):
        s = self.spec
        if s.contraction is not None and not skip_spatial_distortion:
            x = scene_contraction(x, s.contraction)
        points_sdf = None
        if s.use_numerical_gradients:
            d = self.numerical_gradients_delta
            offs = torch.tensor([[d, 0, 0], [-d, 0, 0], [0, d, 0], [0, -d, 0], [0, 0, d], [0, 0, -d]], dtype=x.dtype)
            pts = (x[None] + offs[:, None, :]).reshape(-1, 3)
            points_sdf = self.forward_geonetwork(pts)[..., 0].view(6, *x.shape[:-1])
            grads = torch.stack(
                [0.5 * (points_sdf[0] - points_sdf[1]) / d, 0.5 * (points_sdf[2] - points_sdf[3]) / d, 0.5 * (points_sdf[4] - points_sdf[5]) / d],
                dim=-1,
            )
        else:
            with torch.enable_grad():
                xg = x.detach().clone().requires_grad_(True)
                y = self.forward_geonetwork(xg)[:, :1]
                grads = torch.autograd.grad(y, xg, torch.ones_like(y))[0]
        if return_sdf:
            return grads, points_sdf
        return grads

    # -- sdf_field.py:57-71
    def get_beta(self):
        return self.p["laplace_density.beta"].abs() + self.p["laplace_density.beta_min"]

    def laplace_density(self, sdf, beta=None):
        if beta is None:
            beta = self.get_beta()
        alpha = 1.0 / beta
        return alpha * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))

    # -- sdf_field.py:116-118
    def get_variance(self):
        return torch.exp(self.p["deviation_network.variance"] * 10.0).clip(1e-6, 1e6)

    # -- sdf_field.py:527-530
    @staticmethod
    def get_occupancy(sdf):
        return torch.sigmoid(-10.0 * sdf)

    # -- sdf_field.py:476-525 (sdf & gradients given)
    def get_alpha(self, directions, deltas, sdf, gradients):
        inv_s = self.get_variance()
        true_cos = (directions * gradients).sum(-1, keepdim=True)
        r = self.cos_anneal_ratio
        iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - r) + torch.relu(-true_cos) * r)
        est_next = sdf + iter_cos * deltas * 0.5
        est_prev = sdf - iter_cos * deltas * 0.5
        prev_cdf = torch.sigmoid(est_prev * inv_s)
        next_cdf = torch.sigmoid(est_next * inv_s)
        p = prev_cdf - next_cdf
        c = prev_cdf
        return ((p + 1e-5) / (c + 1e-5)).clip(0.0, 1.0)

    # -- sdf_field.py:532-612
    def get_colors(self, points, directions, gradients, geo_features, camera_indices=None):
        s = self.spec
        N = points.shape[0]
        normals = torch.nn.functional.normalize(gradients, p=2, dim=-1)
        if s.use_diffuse_color:
            raw_diffuse = torch.nn.functional.linear(geo_features, self.p["diffuse_color_pred.weight"], self.p["diffuse_color_pred.bias"])
        if s.use_specular_tint:
            tint = torch.sigmoid(torch.nn.functional.linear(geo_features, self.p["specular_tint_pred.weight"], self.p["specular_tint_pred.bias"]))
        if s.use_reflections:
            refdirs = 2.0 * torch.sum(normals * -directions, dim=-1, keepdim=True) * normals + directions
            d = nerf_encoding(refdirs, 4, 0.0, 3.0, True)
        else:
            d = nerf_encoding(directions, 4, 0.0, 3.0, True)
        if self.training:
            emb = self.p["embedding_appearance.embedding.weight"][camera_indices]
            if not s.use_appearance_embedding:
                emb = torch.zeros_like(emb)
        elif self.use_average_appearance_embedding:
            emb = torch.ones(N, s.appearance_embedding_dim, dtype=points.dtype) * self.p["embedding_appearance.embedding.weight"].mean(dim=0)
        else:
            emb = torch.zeros(N, s.appearance_embedding_dim, dtype=points.dtype)
        if s.use_diffuse_color:
            h = [d, geo_features, emb]
        else:
            h = [points, d, gradients, geo_features, emb]
        if s.use_n_dot_v:
            h.append(torch.sum(normals * directions, dim=-1, keepdim=True))
        h = torch.cat(h, dim=-1)
        n_lin = s.num_layers_color + 1
        for l in range(n_lin):
            W = folded_weight(self.p, f"clin{l}", s.weight_norm)
            h = torch.nn.functional.linear(h, W, self.p[f"clin{l}.bias"])
            if l < n_lin - 1:
                h = torch.relu(h)
        rgb = torch.sigmoid(h)
        if s.use_diffuse_color:
            diffuse_linear = torch.sigmoid(raw_diffuse - math.log(3.0))
            specular_linear = tint * rgb if s.use_specular_tint else 0.5 * rgb
            rgb = torch.clamp(specular_linear + diffuse_linear, 0.0, 1.0)
        return rgb * (1 + 2 * s.rgb_padding) - s.rgb_padding

    # -- sdf_field.py:412-418 (NOTE: un-contracted start positions, like the reference)
    def get_sdf(self, origins, directions, starts):
        """origins/directions [R,3], starts [R,S] -> sdf [R,S]."""
        R, S = starts.shape
        pos = origins[:, None, :] + directions[:, None, :] * starts[..., None]
        return self.forward_geonetwork(pos.reshape(-1, 3))[:, 0].view(R, S)

    # -- sdf_field.py:614-689
    def get_outputs(self, origins, directions, starts, deltas, camera_indices=None, return_alphas=False, return_occupancy=False):
        """origins/directions [R,3]; starts, deltas [R,S].  Returns dict of [R,S,k] tensors (reference FieldHeadNames)."""
        s = self.spec
        R, S = starts.shape
        pos = (origins[:, None, :] + directions[:, None, :] * starts[..., None]).reshape(-1, 3)
        dirs = directions[:, None, :].expand(R, S, 3).reshape(-1, 3)
        x = scene_contraction(pos, s.contraction)
        points_norm = x.norm(dim=-1)
        with torch.enable_grad():
            xg = x.detach().clone().requires_grad_(True)
            h = self.forward_geonetwork(xg)
            sdf, geo = h[:, :1], h[:, 1:]
            if s.use_numerical_gradients:
                grads, sampled = self.gradient(xg.detach(), skip_spatial_distortion=True, return_sdf=True)
                sampled = sampled.view(-1, R, S).permute(1, 2, 0).contiguous()
            else:
                grads = torch.autograd.grad(sdf, xg, torch.ones_like(sdf), retain_graph=False)[0]
                sampled = None
        sdf, geo, x = sdf.detach(), geo.detach(), xg.detach()
        cam = None if camera_indices is None else camera_indices.reshape(R, 1).expand(R, S).reshape(-1)
        rgb = self.get_colors(x, dirs, grads, geo, cam)
        density = self.laplace_density(sdf)
        out = {
            "rgb": rgb.view(R, S, 3),
            "density": density.view(R, S, 1),
            "sdf": sdf.view(R, S, 1),
            "gradients": grads.view(R, S, 3),
            "normals": torch.nn.functional.normalize(grads, p=2, dim=-1).view(R, S, 3),
            "points_norm": points_norm.view(R, S, 1),
            "sampled_sdf": sampled,
            "geo_feature": geo.view(R, S, -1),
        }
        if return_alphas:
            out["alphas"] = self.get_alpha(directions[:, None, :].expand(R, S, 3), deltas[..., None], out["sdf"], out["gradients"])
        if return_occupancy:
            out["occupancy"] = self.get_occupancy(out["sdf"])
        return out


# ----------------------------------------------------------------------------------------------------------------
# parameter initialisation restated from SDFField.__init__ (sdf_field.py:284-363) -- used to build synthetic fields on
# the GPU box where the reference cannot be imported.
# ----------------------------------------------------------------------------------------------------------------
def init_params(spec: FieldSpec, num_images: int = 49, bias: float = 0.5, beta_init: float = 0.3, inside_outside: bool = False,
                seed: int = 0, hash_init_scale: float = 1e-3, perturb: float = 0.0) -> Dict[str, torch.Tensor]:
    """Geometric init + kaiming colour net + uniform hash table.  ``perturb`` > 0 adds N(0, perturb) to every MLP weight
    and fills the (otherwise zero) PE/grid columns of the first layer so that all inputs matter (a 'briefly trained'
    stand-in for parity tests)."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, torch.Tensor] = {}
    dims = spec.geo_dims()
    n_lin = len(dims) - 1
    for l in range(n_lin):
        out_dim = dims[l + 1] - dims[0] if (l + 1) in spec.skip_in else dims[l + 1]
        W = torch.empty(out_dim, dims[l])
        b = torch.zeros(out_dim)
        if l == n_lin - 1:
            mean = np.sqrt(np.pi) / np.sqrt(dims[l])
            W.normal_(mean=(-mean if inside_outside else mean), std=0.0001, generator=g)
            b.fill_(bias if inside_outside else -bias)
        elif l == 0:
            W.zero_()
            W[:, :3].normal_(0.0, np.sqrt(2) / np.sqrt(out_dim), generator=g)
        elif l in spec.skip_in:
            W.normal_(0.0, np.sqrt(2) / np.sqrt(out_dim), generator=g)
            W[:, -(dims[0] - 3):] = 0.0
        else:
            W.normal_(0.0, np.sqrt(2) / np.sqrt(out_dim), generator=g)
        if perturb > 0:
            W = W + perturb * torch.randn(W.shape, generator=g)
            b = b + perturb * torch.randn(b.shape, generator=g)
        p[f"glin{l}.weight_v"] = W
        p[f"glin{l}.weight_g"] = W.norm(dim=1, keepdim=True)
        p[f"glin{l}.bias"] = b
    dims = spec.color_dims()
    for l in range(len(dims) - 1):
        W = torch.empty(dims[l + 1], dims[l])
        bound = math.sqrt(6.0 / dims[l])  # kaiming_uniform_, a=0, fan_in
        W.uniform_(-bound, bound, generator=g)
        b = torch.zeros(dims[l + 1])
        if perturb > 0:
            b = b + perturb * torch.randn(b.shape, generator=g)
        p[f"clin{l}.weight_v"] = W
        p[f"clin{l}.weight_g"] = W.norm(dim=1, keepdim=True)
        p[f"clin{l}.bias"] = b
    if spec.use_diffuse_color:
        p["diffuse_color_pred.weight"] = torch.randn(3, spec.geo_feat_dim, generator=g) / math.sqrt(spec.geo_feat_dim)
        p["diffuse_color_pred.bias"] = torch.zeros(3)
    if spec.use_specular_tint:
        p["specular_tint_pred.weight"] = torch.randn(3, spec.geo_feat_dim, generator=g) / math.sqrt(spec.geo_feat_dim)
        p["specular_tint_pred.bias"] = torch.zeros(3)
    p["laplace_density.beta_min"] = torch.tensor([0.0001])
    p["laplace_density.beta"] = torch.tensor([beta_init])
    p["deviation_network.variance"] = torch.tensor([beta_init])
    p["embedding_appearance.embedding.weight"] = torch.randn(num_images, spec.appearance_embedding_dim, generator=g)
    if spec.use_grid_feature:
        if spec.grid_layout == "torch":
            n = spec.num_levels << spec.log2_hashmap_size
        else:
            gf = hashgrid.growth_factor(spec.num_levels, spec.base_res, spec.max_res)
            n = hashgrid.tcnn_grid_meta(spec.num_levels, spec.hash_features_per_level, spec.log2_hashmap_size, spec.base_res, gf)["total"]
        p["hash_table"] = (torch.rand(n, spec.hash_features_per_level, generator=g) * 2 - 1) * hash_init_scale
    return p

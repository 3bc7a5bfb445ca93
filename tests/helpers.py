"""Shared helpers for the parity tests: build the *product* SDFField for a seeded oracle case."""
import os

import numpy as np
import torch

from oracle import cases
from oracle.field import FieldSpec, OracleField, init_params

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name, device="cpu"):
    return {k: torch.from_numpy(v).to(device) for k, v in np.load(os.path.join(GOLDEN_DIR, name + ".npz")).items()}


def product_field(spec: FieldSpec, params, kw, device="cuda", precision="fp32", table_dtype="fp32"):
    """sdfstudio_b200.SDFField with the oracle's seeded parameters loaded (names match the reference state_dict)."""
    import sdfstudio_b200 as sb

    cfg = sb.SDFFieldConfig(
        num_layers=spec.num_layers, hidden_dim=spec.hidden_dim, geo_feat_dim=spec.geo_feat_dim, num_layers_color=spec.num_layers_color,
        hidden_dim_color=spec.hidden_dim_color, appearance_embedding_dim=spec.appearance_embedding_dim,
        use_appearance_embedding=spec.use_appearance_embedding, bias=kw.get("bias", 0.5), inside_outside=kw.get("inside_outside", False),
        use_grid_feature=spec.use_grid_feature, beta_init=kw.get("beta_init", 0.3), position_encoding_max_degree=spec.position_encoding_max_degree,
        use_diffuse_color=spec.use_diffuse_color, use_specular_tint=spec.use_specular_tint, use_reflections=spec.use_reflections,
        use_n_dot_v=spec.use_n_dot_v, rgb_padding=spec.rgb_padding, off_axis=spec.off_axis, use_numerical_gradients=spec.use_numerical_gradients,
        num_levels=spec.num_levels, max_res=spec.max_res, base_res=spec.base_res, log2_hashmap_size=spec.log2_hashmap_size,
        hash_features_per_level=spec.hash_features_per_level, hash_smoothstep=spec.hash_smoothstep, use_position_encoding=spec.use_position_encoding,
        grid_layout=spec.grid_layout, precision=precision, table_dtype=table_dtype,
    )  # fmt: skip

    class _Contraction:  # duck-typed SceneContraction (only `.order` is read)
        def __init__(self, order):
            self.order = order

    distortion = None
    if spec.contraction is not None:
        distortion = _Contraction(float("inf") if spec.contraction == "linf" else None)
    f = sb.SDFField(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=49, spatial_distortion=distortion)
    sd = {}
    for k, v in params.items():
        if k == "hash_table":
            if spec.grid_layout == "torch":
                sd["encoding.hash_table"] = v
            else:
                sd["encoding.params"] = v.reshape(-1)
        else:
            sd[k] = v
    missing, unexpected = f.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(m.startswith("encoding") or m == "aabb" for m in missing), missing
    f = f.to(device).eval()
    if "mask_level" in kw:
        f.update_mask(kw["mask_level"])
    if "num_grad_delta" in kw:
        f.set_numerical_gradients_delta(kw["num_grad_delta"])
    return f


def build_case(name, device="cuda", precision="fp32", table_dtype="fp32"):
    spec, kw, o, d, cam, nears, fars = cases.case_inputs(name)
    params = init_params(spec, **cases.init_kwargs(kw))
    if table_dtype == "fp16" and "hash_table" in params:
        params["hash_table"] = params["hash_table"].half().float()   # the model IS the fp16-representable table (tcnn semantics)
    oracle = OracleField(spec, params)
    if "mask_level" in kw:
        oracle.update_mask(kw["mask_level"])
    if "num_grad_delta" in kw:
        oracle.numerical_gradients_delta = kw["num_grad_delta"]
    field = product_field(spec, params, kw, device, precision, table_dtype)
    return spec, kw, o, d, cam, nears, fars, oracle, field


def make_bundle(o, d, cam, nears, fars, device="cuda"):
    import sdfstudio_b200 as sb

    R = o.shape[0]
    return sb.RayBundle(origins=o.to(device), directions=d.to(device), pixel_area=torch.ones(R, 1, device=device),
                        directions_norm=torch.ones(R, 1, device=device), camera_indices=cam.view(R, 1).to(device), nears=nears.to(device),
                        fars=fars.to(device))


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor)"""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(floor)).max())


def oracle64(spec, params, kw):
    """float64 instance of the oracle: the 'exact' answer used to calibrate fp32 noise."""
    o = OracleField(spec, params, dtype=torch.float64)
    if "mask_level" in kw:
        o.update_mask(kw["mask_level"])
    if "num_grad_delta" in kw:
        o.numerical_gradients_delta = kw["num_grad_delta"]
    return o


def assert_within_noise(cuda, ref32, ref64, what, factor=4.0, floor=2e-6):
    """|cuda - exact| must not exceed `factor` x the reference's OWN fp32 rounding noise |ref32 - exact| (max-norm), with an
    absolute floor.  Used where the quantity is ill-conditioned in fp32 (numerical gradients, inverse-CDF sampling), so
    that a fixed relative bound would be either meaningless or unattainable by any fp32 implementation."""
    cuda, ref32, ref64 = (t.detach().double().cpu() for t in (cuda, ref32, ref64))
    noise = float((ref32 - ref64).abs().max())
    err = float((cuda - ref64).abs().max())
    bound = max(floor, factor * noise)
    assert err <= bound, f"{what}: |cuda-exact| = {err:.3e} > {bound:.3e} (reference fp32 noise {noise:.3e})"


def cdf_consistency(existing_spacing, weights, new_bins, u, hist_pad, eps=1e-5):
    """max |cdf(new_bin) - u| evaluated in float64: the forward map of PDF sampling is well conditioned even where its
    inverse (the bin position) is not."""
    w = weights.double() + hist_pad
    ws = w.sum(-1, keepdim=True)
    pad = torch.relu(eps - ws)
    w = w + pad / w.shape[-1]
    ws = ws + pad
    cdf = torch.cumsum(w / ws, -1).clamp(max=1.0)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    eb = existing_spacing.double()
    nb = new_bins.double()
    idx = torch.searchsorted(eb.contiguous(), nb.contiguous(), right=True).clamp(1, eb.shape[1] - 1)
    b0, b1 = torch.gather(eb, 1, idx - 1), torch.gather(eb, 1, idx)
    c0, c1 = torch.gather(cdf, 1, idx - 1), torch.gather(cdf, 1, idx)
    t = ((nb - b0) / (b1 - b0).clamp_min(1e-30)).clamp(0, 1)
    return float((c0 + t * (c1 - c0) - u.double()).abs().max())

"""TEST INFRASTRUCTURE -- mints tests/golden/*.npz by running the UNMODIFIED reference (via oracle/ref_import.py).

Container-only (needs /root/reference).  Usage:  python -m oracle.make_golden [case ...]

Each fixture holds the reference's outputs for one seeded case of oracle/cases.py.  Parameters and rays are not
stored: they are regenerated bit-identically from the seed (oracle.field.init_params / oracle.cases.synthetic_rays)
and loaded INTO the reference module with load_state_dict, so the fixture pins reference *arithmetic*, not its RNG.
"""
import os
import sys

import numpy as np
import torch

from . import cases
from .field import FieldSpec, init_params
from .ref_import import ref_modules

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_reference_field(R, spec: FieldSpec, params, kw):
    """Instantiate the reference SDFField for `spec` and load the seeded parameters into it."""
    cfg = R.sdf_field.SDFFieldConfig(
        num_layers=spec.num_layers, hidden_dim=spec.hidden_dim, geo_feat_dim=spec.geo_feat_dim,
        num_layers_color=spec.num_layers_color, hidden_dim_color=spec.hidden_dim_color,
        appearance_embedding_dim=spec.appearance_embedding_dim, use_appearance_embedding=spec.use_appearance_embedding,
        bias=kw["bias"], inside_outside=kw.get("inside_outside", False), use_grid_feature=spec.use_grid_feature,
        beta_init=kw["beta_init"], position_encoding_max_degree=spec.position_encoding_max_degree,
        use_diffuse_color=spec.use_diffuse_color, use_specular_tint=spec.use_specular_tint,
        use_reflections=spec.use_reflections, use_n_dot_v=spec.use_n_dot_v, rgb_padding=spec.rgb_padding,
        off_axis=spec.off_axis, use_numerical_gradients=spec.use_numerical_gradients, num_levels=spec.num_levels,
        max_res=spec.max_res, base_res=spec.base_res, log2_hashmap_size=spec.log2_hashmap_size,
        hash_features_per_level=spec.hash_features_per_level, hash_smoothstep=spec.hash_smoothstep,
        use_position_encoding=spec.use_position_encoding,
    )  # fmt: skip
    distortion = None
    if spec.contraction is not None:
        distortion = R.spatial_distortions.SceneContraction(order=float("inf") if spec.contraction == "linf" else None)
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    field = R.sdf_field.SDFField(cfg, aabb, num_images=49, spatial_distortion=distortion).eval()
    sd = {}
    for k, v in params.items():
        sd["encoding.enc.hash_table" if k == "hash_table" else k] = v
    sd["aabb"] = aabb
    if not spec.use_grid_feature:
        sd["encoding.enc.hash_table"] = field.encoding.enc.hash_table.data
    missing, unexpected = field.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(m.startswith("encoding") for m in missing), missing
    if "mask_level" in kw:
        field.update_mask(kw["mask_level"])
    if "num_grad_delta" in kw:
        field.set_numerical_gradients_delta(kw["num_grad_delta"])
    return field


def make_bundle(R, o, d, cam, nears, fars):
    n = o.shape[0]
    return R.RayBundle(origins=o, directions=d, pixel_area=torch.ones(n, 1), directions_norm=torch.ones(n, 1),
                       camera_indices=cam.view(n, 1).long(), nears=nears.clone(), fars=fars.clone())


def npy(t):
    return None if t is None else t.detach().cpu().numpy()


def mint(name: str):
    R = ref_modules()
    H = R.FieldHeadNames
    spec, kw, o, d, cam, nears, fars = cases.case_inputs(name)
    params = init_params(spec, **cases.init_kwargs(kw))
    field = build_reference_field(R, spec, params, kw)
    rb = make_bundle(R, o, d, cam, nears, fars)
    S = kw["S"]
    out = {}

    # ---- initial sampler + field forward (a1, a2, a9-a16) ----
    kind = kw.get("spacing", "uniform")
    sampler = {"uniform": R.ray_samplers.UniformSampler, "piecewise": R.ray_samplers.UniformLinDispPiecewiseSampler,
               "lindisp": R.ray_samplers.LinearDisparitySampler}[kind](num_samples=S).eval()
    rs = sampler(rb)
    out["spacing_bins"] = npy(torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[:, -1:, 0]], -1))
    out["euclid_bins"] = npy(torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[:, -1:, 0]], -1))
    fo = field(rs, return_alphas=True, return_occupancy=True)
    for key, nm in [(H.RGB, "rgb"), (H.DENSITY, "density"), (H.SDF, "sdf"), (H.NORMAL, "normals"), (H.GRADIENT, "gradients"),
                    ("points_norm", "points_norm"), (H.ALPHA, "alphas"), (H.OCCUPANCY, "occupancy")]:
        out[nm] = npy(fo[key])
    if fo["sampled_sdf"] is not None:
        out["sampled_sdf"] = npy(fo["sampled_sdf"])
    out["get_sdf"] = npy(field.get_sdf(rs))  # un-contracted positions (sdf_field.py:412-418)
    g = torch.Generator().manual_seed(kw["seed"] + 77)
    pts = (torch.rand(200, 3, generator=g) * 2 - 1) * (3.0 if spec.contraction else 1.0)
    out["points"] = npy(pts)
    out["geo_points"] = npy(field.forward_geonetwork(pts.clone()))
    out["grad_points"] = npy(field.gradient(pts.clone()))

    # ---- weights + renderers (a17-a20) ----
    w_a, T_a = rs.get_weights_and_transmittance_from_alphas(fo[H.ALPHA])
    w_d, T_d = rs.get_weights_and_transmittance(fo[H.DENSITY])
    out["weights_alpha"], out["trans_alpha"] = npy(w_a), npy(T_a)
    out["weights_density"], out["trans_density"] = npy(w_d), npy(T_d)
    white = torch.ones(3)
    out["render_rgb_white"] = npy(R.renderers.RGBRenderer(background_color=white).eval()(fo[H.RGB], w_a))
    out["render_rgb_last"] = npy(R.renderers.RGBRenderer(background_color="last_sample").eval()(fo[H.RGB], w_a))
    out["render_rgb_white_train"] = npy(R.renderers.RGBRenderer(background_color=white).train()(fo[H.RGB], w_d))
    out["render_depth_expected"] = npy(R.renderers.DepthRenderer("expected")(w_a, rs))
    out["render_depth_median"] = npy(R.renderers.DepthRenderer("median")(w_a, rs))
    out["render_acc"] = npy(R.renderers.AccumulationRenderer()(w_a))
    out["render_normal"] = npy(R.renderers.SemanticRenderer()(fo[H.NORMAL], w_a))

    # ---- samplers driven by the field (a3-a7) ----
    def bins_of(s):
        return (npy(torch.cat([s.spacing_starts[..., 0], s.spacing_ends[:, -1:, 0]], -1)),
                npy(torch.cat([s.frustums.starts[..., 0], s.frustums.ends[:, -1:, 0]], -1)))

    if name in ("neusfacto_c1", "volsdf_stock", "neusfacto_c1_init"):
        # PDFSampler with captured searchsorted indices
        inds_log = []
        orig_ss = torch.searchsorted

        def ss(*a, **k):
            r = orig_ss(*a, **k)
            inds_log.append(r.clone())
            return r

        wts = torch.rand(kw["R"], S, 1, generator=g) ** 4
        torch.searchsorted = ss
        try:
            new = R.ray_samplers.PDFSampler(include_original=False, histogram_padding=0.01).eval()(rb, rs, wts, num_samples=24)
            new2 = R.ray_samplers.PDFSampler(include_original=True, histogram_padding=1e-5).eval()(rb, rs, wts, num_samples=16)
        finally:
            torch.searchsorted = orig_ss
        out["pdf_weights"] = npy(wts)
        out["pdf_inds"] = npy(inds_log[0])
        out["pdf_spacing"], out["pdf_euclid"] = bins_of(new)
        out["pdf_inc_inds"] = npy(inds_log[1])
        out["pdf_inc_spacing"], out["pdf_inc_euclid"] = bins_of(new2)
        merged, sidx = R.ray_samplers.ErrorBoundedSampler().merge_ray_samples(rb, rs, new)
        out["merge_spacing"], out["merge_euclid"] = bins_of(merged)
        out["merge_sorted_index"] = npy(sidx)

        # NeuSSampler (eval: no jitter)
        ns = R.ray_samplers.NeuSSampler().eval()
        inds_log.clear()
        torch.searchsorted = ss
        try:
            rs_n = ns(rb, sdf_fn=field.get_sdf)
        finally:
            torch.searchsorted = orig_ss
        out["neus_spacing"], out["neus_euclid"] = bins_of(rs_n)
        out["neus_inds"] = np.stack([npy(i) for i in inds_log])

        # ErrorBoundedSampler (volsdf.py:35-40 numbers)
        eb = R.ray_samplers.ErrorBoundedSampler(num_samples=64, num_samples_eval=128, num_samples_extra=32).eval()
        torch.manual_seed(0)
        rs_e, eik = eb(rb, density_fn=field.laplace_density, sdf_fn=field.get_sdf)
        out["eb_spacing"], out["eb_euclid"] = bins_of(rs_e)

        # UniSurfSampler
        us = R.ray_samplers.UniSurfSampler().eval()
        rs_u, surf = us(make_bundle(R, o, d, cam, nears, fars), occupancy_fn=field.get_occupancy, sdf_fn=field.get_sdf, return_surface_points=True)
        out["uni_spacing"], out["uni_euclid"] = bins_of(rs_u)
        out["uni_surface"] = npy(surf)

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **{k: v for k, v in out.items() if v is not None})
    print(f"wrote {path}: {os.path.getsize(path)/1024:.0f} KiB, keys={sorted(out)}")


if __name__ == "__main__":
    names = sys.argv[1:] or list(cases.CASES)
    torch.set_num_threads(8)
    for n in names:
        mint(n)

"""Pins oracle/ (the CPU restatement) against golden vectors minted from the UNMODIFIED reference
(oracle/make_golden.py).  CPU-only; runs everywhere."""
import os

import numpy as np
import pytest
import torch

from oracle import cases, render, samplers
from oracle.field import OracleField, init_params

FIELD_CASES = list(cases.CASES) + list(cases.CPU_CASES)      # CPU_CASES: oracle-vs-reference only (e.g. the L2 SceneContraction)
SAMPLER_CASES = ["neusfacto_c1", "neusfacto_c1_init", "volsdf_stock"]


def load(golden_dir, name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, name + ".npz")).items()}


def build(name):
    spec, kw, o, d, cam, nears, fars = cases.case_inputs(name)
    params = init_params(spec, **cases.init_kwargs(kw))
    f = OracleField(spec, params)
    if "mask_level" in kw:
        f.update_mask(kw["mask_level"])
    if "num_grad_delta" in kw:
        f.numerical_gradients_delta = kw["num_grad_delta"]
    return spec, kw, o, d, cam, nears, fars, f


def close(a, b, rtol=2e-5, atol=2e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", FIELD_CASES)
def test_field_matches_reference(golden_dir, name):
    G = load(golden_dir, name)
    spec, kw, o, d, cam, nears, fars, f = build(name)
    b = samplers.spaced_sampler(nears, fars, kw["S"], kw.get("spacing", "uniform"))
    assert torch.equal(b.spacing, G["spacing_bins"])
    assert torch.equal(b.euclid, G["euclid_bins"])
    out = f.get_outputs(o, d, b.starts, b.deltas, cam, return_alphas=True, return_occupancy=True)
    # same torch-CPU kernels in the same order => (near) bit parity; tolerance only covers BLAS blocking differences
    for k in ["sdf", "density", "gradients", "normals", "points_norm", "rgb", "alphas", "occupancy"]:
        close(out[k], G[k])
    if "sampled_sdf" in G:
        close(out["sampled_sdf"], G["sampled_sdf"])
    close(f.get_sdf(o, d, b.starts)[..., None], G["get_sdf"])
    close(f.forward_geonetwork(G["points"]), G["geo_points"])
    close(f.gradient(G["points"]), G["grad_points"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", FIELD_CASES)
def test_weights_and_renderers_match_reference(golden_dir, name):
    G = load(golden_dir, name)
    b_st, b_en = G["euclid_bins"][:, :-1, None], G["euclid_bins"][:, 1:, None]
    w_a, T_a = samplers.weights_from_alphas(G["alphas"][..., 0])
    assert torch.equal(w_a[..., None], G["weights_alpha"]) and torch.equal(T_a[..., None], G["trans_alpha"])
    w_d, T_d = samplers.weights_from_density((b_en - b_st)[..., 0], G["density"][..., 0])
    assert torch.equal(w_d[..., None], G["weights_density"]) and torch.equal(T_d[..., None], G["trans_density"])
    w = G["weights_alpha"]
    assert torch.equal(render.render_rgb(G["rgb"], w, torch.ones(3)), G["render_rgb_white"])
    assert torch.equal(render.render_rgb(G["rgb"], w, "last_sample"), G["render_rgb_last"])
    assert torch.equal(render.render_rgb(G["rgb"], G["weights_density"], torch.ones(3), training=True), G["render_rgb_white_train"])
    assert torch.equal(render.render_depth(w, b_st, b_en, "expected"), G["render_depth_expected"])
    assert torch.equal(render.render_depth(w, b_st, b_en, "median"), G["render_depth_median"])
    assert torch.equal(render.render_accumulation(w), G["render_acc"])
    assert torch.equal(render.render_semantics(G["normals"], w), G["render_normal"])


@pytest.mark.parametrize("name", SAMPLER_CASES)
def test_pdf_and_merge_bit_exact(golden_dir, name):
    G = load(golden_dir, name)
    spec, kw, o, d, cam, nears, fars, f = build(name)
    b = samplers.spaced_sampler(nears, fars, kw["S"], "uniform")
    new, inds = samplers.pdf_sampler(b, G["pdf_weights"][..., 0], 24, histogram_padding=0.01, return_indices=True)
    assert torch.equal(inds, G["pdf_inds"])
    assert torch.equal(new.spacing, G["pdf_spacing"]) and torch.equal(new.euclid, G["pdf_euclid"])
    new2, inds2 = samplers.pdf_sampler(b, G["pdf_weights"][..., 0], 16, histogram_padding=1e-5, include_original=True, return_indices=True)
    assert torch.equal(inds2, G["pdf_inc_inds"])
    assert torch.equal(new2.spacing, G["pdf_inc_spacing"]) and torch.equal(new2.euclid, G["pdf_inc_euclid"])
    merged, sidx = samplers.merge_bins(b, new)
    assert torch.equal(sidx, G["merge_sorted_index"])
    assert torch.equal(merged.spacing, G["merge_spacing"]) and torch.equal(merged.euclid, G["merge_euclid"])


@pytest.mark.parametrize("name", SAMPLER_CASES)
def test_field_driven_samplers(golden_dir, name):
    G = load(golden_dir, name)
    spec, kw, o, d, cam, nears, fars, f = build(name)
    sdf_fn = lambda starts: f.get_sdf(o, d, starts)
    trace = []
    nb = samplers.neus_sampler(nears, fars, sdf_fn, trace=trace)
    assert torch.equal(torch.stack([t["inds"] for t in trace]), G["neus_inds"])
    close(nb.spacing, G["neus_spacing"], rtol=0, atol=1e-6)
    close(nb.euclid, G["neus_euclid"], rtol=0, atol=4e-6)
    eb = samplers.error_bounded_sampler(nears, fars, sdf_fn, f.get_beta())
    assert eb.spacing.shape == G["eb_spacing"].shape
    close(eb.spacing, G["eb_spacing"], rtol=0, atol=1e-6)
    close(eb.euclid, G["eb_euclid"], rtol=0, atol=4e-6)
    ub, surf, mask = samplers.unisurf_sampler(o, d, nears, fars, sdf_fn)
    close(ub.euclid, G["uni_euclid"], rtol=0, atol=4e-6)
    if mask.any():
        close(surf, G["uni_surface"], rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------------------------
# step before the path: camera rays, colliders, meshing lattice
# ---------------------------------------------------------------------------------------------------------------
def test_raygen_oracle_matches_reference_golden(golden_dir):
    from oracle import raygen

    g = load(golden_dir, "raygen")
    c = raygen.raygen_case()
    o, d, area, dnorm = raygen.generate_rays(c["fx"], c["fy"], c["cx"], c["cy"], c["cam_type"], c["c2w"], c["idx"], c["coords"])
    assert torch.equal(o, g["origins"])
    assert torch.equal(d, g["directions"])
    assert torch.equal(dnorm, g["directions_norm"])
    assert torch.equal(area, g["pixel_area"])
    aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
    n, f = raygen.collide_aabb(o, d, aabb, near_plane=0.05)
    assert torch.equal(n, g["aabb_train_nears"]) and torch.equal(f, g["aabb_train_fars"])
    n, f = raygen.collide_aabb(o, d, aabb, near_plane=0.0)
    assert torch.equal(n, g["aabb_eval_nears"]) and torch.equal(f, g["aabb_eval_fars"])
    n, f = raygen.collide_near_far(o, 0.5, 4.5)
    assert torch.equal(n, g["nf_nears"]) and torch.equal(f, g["nf_fars"])
    n, f = raygen.collide_sphere(o, d, 1.3, False)
    assert torch.equal(n, g["sph_nears"]) and torch.equal(f, g["sph_fars"])
    n, f = raygen.collide_sphere(o, d, 1.3, True)
    assert torch.equal(n, g["sphsoft_nears"]) and torch.equal(f, g["sphsoft_fars"])
    assert torch.equal(raygen.lattice((-1.0, -0.7, -1.0), (0.3, 1.0, 1.0), (9, 5, 7)), g["lattice"])

"""GPU parity of the step before the path: camera rays, colliders, meshing lattice / SDF grid.
Ray generation and colliders are compared with goldens minted from the unmodified reference (tests/golden/raygen.npz)."""
import numpy as np
import pytest
import torch

from oracle import raygen

from helpers import build_case, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _ulp_close(a, b, ulps=2):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    tol = ulps * torch.finfo(torch.float32).eps * b.abs().clamp_min(1e-30)
    return bool(((a - b).abs() <= tol).all())


def test_generate_rays_and_colliders_match_reference():
    import sdfstudio_b200 as sb

    g = load_golden("raygen")
    c = raygen.raygen_case()
    cams = sb.cameras.Cameras(c["c2w"], c["fx"], c["fy"], c["cx"], c["cy"], 384, 384, camera_type=c["cam_type"], device=torch.device("cuda"))
    rb = cams.generate_rays(c["idx"], c["coords"])
    assert torch.equal(rb.origins.cpu(), g["origins"])
    persp = (c["cam_type"][c["idx"]] == raygen.PERSPECTIVE)
    # perspective rays: same fp32 expression tree -> bit-exact directions; fisheye goes through sin / cos (<= 2 ulp).
    # pixel_area goes through torch.sqrt, whose vectorised CPU kernel is 1 ulp off the correctly rounded value in ~0.6 % of the
    # cases on AVX-512 hosts -> compared to a few ulp of the two factors
    assert torch.equal(rb.directions.cpu()[persp], g["directions"][persp])
    assert torch.equal(rb.directions_norm.cpu()[persp], g["directions_norm"][persp])
    assert _ulp_close(rb.pixel_area[persp], g["pixel_area"][persp], 4)
    assert (rb.pixel_area.cpu()[persp] == g["pixel_area"][persp]).float().mean() > 0.95
    assert (rb.directions.cpu()[~persp] - g["directions"][~persp]).abs().max() < 3e-7
    assert rel_err(rb.pixel_area[~persp], g["pixel_area"][~persp], floor=1e-9) < 2e-3      # difference of nearly equal unit vectors
    assert _ulp_close(rb.directions_norm[~persp], g["directions_norm"][~persp], 4)

    # colliders on the reference's own rays
    def bundle():
        return sb.RayBundle(origins=g["origins"].cuda(), directions=g["directions"].cuda(), pixel_area=g["pixel_area"].cuda())

    class _Box:
        aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])

    col = sb.AABBBoxCollider(_Box(), near_plane=0.05).train()
    b = col(bundle())
    assert torch.equal(b.nears.cpu(), g["aabb_train_nears"]) and torch.equal(b.fars.cpu(), g["aabb_train_fars"])
    b = col.eval()(bundle())
    assert torch.equal(b.nears.cpu(), g["aabb_eval_nears"]) and torch.equal(b.fars.cpu(), g["aabb_eval_fars"])
    b = sb.NearFarCollider(0.5, 4.5)(bundle())
    assert torch.equal(b.nears.cpu(), g["nf_nears"]) and torch.equal(b.fars.cpu(), g["nf_fars"])
    b = sb.SphereCollider(radius=1.3)(bundle())          # through torch.sqrt: 1 ulp of sqrt(under_sqrt), see above
    assert (b.nears.cpu() - g["sph_nears"]).abs().max() < 5e-7 and (b.fars.cpu() - g["sph_fars"]).abs().max() < 5e-7
    assert (b.fars.cpu() == g["sph_fars"]).float().mean() > 0.9
    b = sb.SphereCollider(radius=1.3, soft_intersection=True)(bundle())
    assert (b.nears.cpu() - g["sphsoft_nears"]).abs().max() < 5e-7 and (b.fars.cpu() - g["sphsoft_fars"]).abs().max() < 5e-7
    # a collider never overwrites nears / fars that are already set (scene_colliders.py:41-45)
    pre = bundle()
    pre.nears, pre.fars = torch.zeros(513, 1, device="cuda"), torch.ones(513, 1, device="cuda")
    assert sb.NearFarCollider(0.5, 4.5)(pre).nears.max() == 0


def test_whole_image_rays():
    import sdfstudio_b200 as sb

    c = raygen.raygen_case()
    cams = sb.cameras.Cameras(c["c2w"], c["fx"], c["fy"], c["cx"], c["cy"], 48, 32, camera_type=c["cam_type"], device=torch.device("cuda"))
    rb = cams.generate_rays(2)
    assert rb.origins.shape == (32 * 48, 3)
    ys, xs = torch.meshgrid(torch.arange(32.0) + 0.5, torch.arange(48.0) + 0.5, indexing="ij")
    coords = torch.stack([ys, xs], -1).reshape(-1, 2)
    o, d, area, dn = raygen.generate_rays(c["fx"], c["fy"], c["cx"], c["cy"], c["cam_type"], c["c2w"], torch.full((32 * 48,), 2), coords)
    assert torch.equal(rb.directions.cpu(), d) and _ulp_close(rb.pixel_area, area, 4)


def test_lattice_and_sdf_grid():
    import sdfstudio_b200 as sb

    g = load_golden("raygen")
    pts = sb.meshing.lattice_points((-1.0, -0.7, -1.0), (0.3, 1.0, 1.0), (9, 5, 7), 0, 315, "cuda")
    assert torch.equal(pts.cpu(), g["lattice"])
    part = sb.meshing.lattice_points((-1.0, -0.7, -1.0), (0.3, 1.0, 1.0), (9, 5, 7), 100, 57, "cuda")
    assert torch.equal(part.cpu(), g["lattice"][100:157])

    # SDF on a dense grid == the oracle's forward_geonetwork on the reference lattice (extract_mesh.py:97-126)
    spec, kw, o, d, cam, nears, fars, oracle, field = build_case("neusfacto_c1")
    res = (20, 17, 23)
    grid = sb.meshing.evaluate_sdf_grid(field, res, (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), chunk=3000)
    ref = oracle.forward_geonetwork(raygen.lattice((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0), res))[:, 0].view(*res)
    assert rel_err(grid, ref) < 1e-4
    fn = sb.meshing.sdf_fn(field, level=0.1)
    x = torch.rand(1000, 3, device="cuda") * 2 - 1
    assert rel_err(fn(x), oracle.forward_geonetwork(x.cpu())[:, 0] - 0.1) < 1e-4
    # the fused sdf-only tensor-core path gives the same grid at the fp32-noise level
    spec, kw, o, d, cam, nears, fars, oracle, field_tc = build_case("neusfacto_c1", precision="bf16x3")
    grid_tc = sb.meshing.evaluate_sdf_grid(field_tc, res, chunk=4096)
    assert float((grid_tc.cpu() - ref).abs().max()) < 5e-5


def test_surface_renderer_image_vs_oracle():
    """cameras -> AABB collider -> NeuS sampler -> SDFField -> renderers as ONE composition (SurfaceRenderer) against the CPU
    oracle chained the same way; chunked rendering (base_model.py:165-189) reproduces the unchunked image."""
    import sdfstudio_b200 as sb
    from oracle import render, samplers

    spec, kw, o_, d_, cam_, nears_, fars_, oracle, field = build_case("neusfacto_c1")
    c = raygen.raygen_case()
    H, W = 12, 16
    persp = torch.full_like(c["cam_type"], raygen.PERSPECTIVE)
    cams = sb.cameras.Cameras(c["c2w"], c["fx"] / 24, c["fy"] / 24, c["cx"] / 24, c["cy"] / 24, W, H, camera_type=persp, device=torch.device("cuda"))
    rb = cams.generate_rays(1)

    class _Box:
        aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])

    sampler = sb.NeuSSampler(num_samples=16, num_samples_importance=16, num_samples_outside=0, num_upsample_steps=2, base_variance=64).eval()
    model = sb.SurfaceRenderer(field, sampler, collider=sb.AABBBoxCollider(_Box()).eval(), kind="neus", background_color="white").eval()
    with torch.no_grad():
        img = model.get_outputs_for_camera_ray_bundle(rb, image_shape=(H, W))
        model.eval_num_rays_per_chunk = 50
        rb2 = cams.generate_rays(1)
        img_chunked = model.get_outputs_for_camera_ray_bundle(rb2, image_shape=(H, W))
    assert img["rgb"].shape == (H, W, 3) and img["depth"].shape == (H, W, 1)
    for k in ("rgb", "normal", "accumulation"):
        assert torch.equal(img[k], img_chunked[k]), k

    # oracle chain on the same rays
    ys, xs = torch.meshgrid(torch.arange(float(H)) + 0.5, torch.arange(float(W)) + 0.5, indexing="ij")
    o, d, area, dn = raygen.generate_rays(c["fx"] / 24, c["fy"] / 24, c["cx"] / 24, c["cy"] / 24, persp, c["c2w"], torch.full((H * W,), 1),
                                          torch.stack([ys, xs], -1).reshape(-1, 2))
    o, d = o.contiguous(), d.contiguous()
    nears, fars = raygen.collide_aabb(o, d, _Box.aabb, 0.0)
    bins = samplers.neus_sampler(nears, fars, lambda st: oracle.get_sdf(o, d, st), num_samples=16, num_samples_importance=16, num_upsample_steps=2,
                                 base_variance=64.0)
    oo = oracle.get_outputs(o, d, bins.starts, bins.deltas, torch.full((H * W,), 1), return_alphas=True)
    ow, _ = samplers.weights_from_alphas(oo["alphas"][..., 0])
    orgb = render.render_rgb(oo["rgb"], ow[..., None], torch.ones(3))
    odepth = render.render_depth(ow[..., None], bins.starts[..., None], bins.ends[..., None], "expected") / dn
    assert rel_err(img["rgb"].view(-1, 3), orgb, floor=1e-2) < 2e-4
    assert rel_err(img["depth"].view(-1, 1), odepth, floor=1e-2) < 2e-4
    assert rel_err(img["accumulation"].view(-1, 1), ow.sum(-1, keepdim=True), floor=1e-2) < 2e-4

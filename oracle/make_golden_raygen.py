"""TEST INFRASTRUCTURE -- mints tests/golden/raygen.npz from the UNMODIFIED reference (run in the build container only):
Cameras.generate_rays (cameras/cameras.py), AABBBoxCollider / NearFarCollider / SphereCollider (scene_colliders.py) and the
np.linspace lattice of marching_cubes.py, on the seeded case of oracle.raygen.raygen_case.

    python -m oracle.make_golden_raygen
"""
import os
import warnings

import numpy as np
import torch

from . import raygen, ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "raygen.npz")


def main():
    ref_import.install_shims()
    warnings.simplefilter("ignore")
    from nerfstudio.cameras.cameras import Cameras
    from nerfstudio.cameras.rays import RayBundle
    from nerfstudio.data.scene_box import SceneBox
    from nerfstudio.model_components import scene_colliders as sc

    c = raygen.raygen_case()
    cams = Cameras(camera_to_worlds=c["c2w"], fx=c["fx"], fy=c["fy"], cx=c["cx"], cy=c["cy"], width=384, height=384, camera_type=c["cam_type"])
    rb = cams.generate_rays(camera_indices=c["idx"][:, None], coords=c["coords"])
    out = {"origins": rb.origins, "directions": rb.directions, "pixel_area": rb.pixel_area, "directions_norm": rb.directions_norm}

    def fresh():
        # the tensors exactly as Cameras hands them over (origins is the strided view c2w[..., :3, 3]: torch's CPU norm rounds
        # differently for contiguous and strided inputs, so the layout is part of the reference result)
        return RayBundle(origins=rb.origins, directions=rb.directions, pixel_area=rb.pixel_area)

    box = SceneBox(aabb=torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]]))
    col = sc.AABBBoxCollider(box, near_plane=0.05).train()
    b = col(fresh())
    out["aabb_train_nears"], out["aabb_train_fars"] = b.nears, b.fars
    b = col.eval()(fresh())
    out["aabb_eval_nears"], out["aabb_eval_fars"] = b.nears, b.fars
    b = sc.NearFarCollider(0.5, 4.5)(fresh())
    out["nf_nears"], out["nf_fars"] = b.nears, b.fars
    b = sc.SphereCollider(radius=1.3)(fresh())
    out["sph_nears"], out["sph_fars"] = b.nears, b.fars
    b = sc.SphereCollider(radius=1.3, soft_intersection=True)(fresh())
    out["sphsoft_nears"], out["sphsoft_fars"] = b.nears, b.fars

    # marching_cubes.py:49-56 lattice (restated call sequence: the reference function itself needs CUDA)
    x = np.linspace(-1.0, 0.3, 9)
    y = np.linspace(-0.7, 1.0, 5)
    z = np.linspace(-1.0, 1.0, 7)
    xx, yy, zz = np.meshgrid(x, y, z, indexing="ij")
    out["lattice"] = torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)
    np.savez_compressed(GOLDEN, **{k: v.detach().cpu().numpy() for k, v in out.items()})
    print("wrote", GOLDEN, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE -- CPU restatement of the step before the hot path:
camera ray generation (nerfstudio/cameras/cameras.py:459-695, perspective + fisheye, no distortion parameters),
scene colliders (nerfstudio/model_components/scene_colliders.py:47-163) and the meshing lattice
(nerfstudio/utils/marching_cubes.py:49-56).  Pinned by tests/golden/raygen.npz (minted from the unmodified reference by
oracle/make_golden_raygen.py).
"""
import math

import numpy as np
import torch

PERSPECTIVE, FISHEYE = 1, 2  # CameraType values, cameras/cameras.py:38-43


def generate_rays(fx, fy, cx, cy, cam_type, c2w, camera_indices, coords):
    """per-camera [C] intrinsics / types, c2w [C,3,4]; camera_indices [N] long, coords [N,2] = (y, x).
    Returns origins, directions [N,3], pixel_area, directions_norm [N,1]."""
    idx = camera_indices.long()
    y, x = coords[..., 0], coords[..., 1]                                    # :549-550
    fx, fy, cx, cy = fx[idx], fy[idx], cx[idx], cy[idx]
    coord = torch.stack([(x - cx) / fx, -(y - cy) / fy], -1)                 # :574-576
    coord_x = torch.stack([(x - cx + 1) / fx, -(y - cy) / fy], -1)
    coord_y = torch.stack([(x - cx) / fx, -(y - cy + 1) / fy], -1)
    cs = torch.stack([coord, coord_x, coord_y], dim=0)                       # [3, N, 2]
    t = cam_type[idx]
    dirs = torch.empty(3, idx.shape[0], 3, dtype=coords.dtype)
    persp = t == PERSPECTIVE
    dirs[:, persp, 0] = cs[:, persp, 0]
    dirs[:, persp, 1] = cs[:, persp, 1]
    dirs[:, persp, 2] = -1.0                                                 # :616-621
    fish = t == FISHEYE
    if fish.any():                                                           # :623-634
        theta = torch.clip(torch.sqrt(torch.sum(cs**2, dim=-1)), 0.0, math.pi)
        st = torch.sin(theta)
        dirs[:, fish, 0] = (cs[..., 0] * st / theta)[:, fish]
        dirs[:, fish, 1] = (cs[..., 1] * st / theta)[:, fish]
        dirs[:, fish, 2] = -torch.cos(theta)[:, fish]
    m = c2w[idx]                                                             # [N,3,4]
    rot = m[:, :3, :3]
    dirs = torch.sum(dirs[..., None, :] * rot, dim=-1)                       # :662-664
    dnorm = torch.norm(dirs, dim=-1, keepdim=True)[0]
    dirs = torch.nn.functional.normalize(dirs, dim=-1)
    dx = torch.sqrt(torch.sum((dirs[0] - dirs[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((dirs[0] - dirs[2]) ** 2, dim=-1))
    return m[:, :3, 3], dirs[0], (dx * dy)[..., None], dnorm


def collide_aabb(o, d, aabb, near_plane=0.0):
    """scene_colliders.py:61-98.  aabb [2,3]."""
    inv = 1.0 / (d + 1e-6)
    t_lo = (aabb[0] - o) * inv
    t_hi = (aabb[1] - o) * inv
    nears = torch.max(torch.minimum(t_lo, t_hi), dim=1).values
    fars = torch.min(torch.maximum(t_lo, t_hi), dim=1).values
    nears = torch.clamp(nears, min=near_plane)
    fars = torch.maximum(fars, nears + 1e-6)
    return nears[..., None], fars[..., None]


def collide_near_far(o, near, far):
    ones = torch.ones_like(o[..., 0:1])
    return ones * near, ones * far


def collide_sphere(o, d, radius=1.0, soft=False):
    """scene_colliders.py:149-163."""
    rc = (d * o).sum(dim=-1, keepdim=True)
    us = rc**2 - (o.norm(p=2, dim=-1, keepdim=True) ** 2 - radius**2)
    us = us.clamp_min(0.01)
    if soft:
        us = torch.ones_like(us) * radius
    inter = (torch.sqrt(us) * torch.tensor([-1.0, 1.0], dtype=o.dtype) - rc).clamp_min(0.01)
    return inter[:, 0:1], inter[:, 1:2]


def lattice(bbox_min, bbox_max, res):
    """marching_cubes.py:49-56: np.linspace per axis, meshgrid 'ij', float32 points [rx*ry*rz, 3]."""
    res = (res,) * 3 if isinstance(res, int) else tuple(res)
    ax = [np.linspace(bbox_min[k], bbox_max[k], res[k]) for k in range(3)]
    xx, yy, zz = np.meshgrid(*ax, indexing="ij")
    return torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)


def raygen_case(seed=21, n_cams=7, n_rays=513):
    """Seeded synthetic cameras + pixel coordinates shared by the golden script and the tests."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.randn(n_cams, 3, generator=g)
    pos = pos / pos.norm(dim=-1, keepdim=True) * 2.7
    fwd = -pos / pos.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm(dim=-1, keepdim=True)
    up2 = torch.linalg.cross(right, fwd)
    c2w = torch.stack([right, up2, -fwd, pos], dim=-1)                       # [C,3,4], camera looks along -z
    fx = 900.0 + 50.0 * torch.rand(n_cams, generator=g)
    fy = 900.0 + 50.0 * torch.rand(n_cams, generator=g)
    cx = 190.0 + 20.0 * torch.rand(n_cams, generator=g)
    cy = 190.0 + 20.0 * torch.rand(n_cams, generator=g)
    cam_type = torch.full((n_cams,), PERSPECTIVE, dtype=torch.int64)
    cam_type[-2:] = FISHEYE
    idx = torch.randint(0, n_cams, (n_rays,), generator=g)
    coords = torch.floor(torch.rand(n_rays, 2, generator=g) * 384.0) + 0.5   # (y, x) pixel centres
    return dict(fx=fx, fy=fy, cx=cx, cy=cy, cam_type=cam_type, c2w=c2w.contiguous(), idx=idx, coords=coords)

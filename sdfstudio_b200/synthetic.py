"""Synthetic inputs for benchmarks / smoke tests (there is no dataset or checkpoint access on the GPU box).

``dtu_like_rays``: DTU-scan65-shaped rays (docs/sdfstudio-data.md:26-47 of the reference):
49 cameras on a sphere of radius 2.7 looking at the origin through a 384x384 pinhole (fx 925.5, fy 922.6, cx 199.4,
cy 198.1), uniformly random pixels, near 0.5 / far 4.5.
``perturb_field_``: turns the geometric initialisation into a 'briefly trained' stand-in so that the hash-grid and
positional-encoding columns of the first layer (exactly zero at init, sdf_field.py:300-303) contribute.
"""
import torch


def dtu_like_rays(n_rays: int, seed: int, radius: float = 2.7):
    g = torch.Generator().manual_seed(seed)
    n_views = 49
    cam = torch.randn(n_views, 3, generator=g)
    cam = cam / cam.norm(dim=-1, keepdim=True) * radius
    idx = torch.randint(0, n_views, (n_rays,), generator=g)
    o = cam[idx]
    fwd = -o / o.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    up2 = torch.linalg.cross(right, fwd)
    px = torch.rand(n_rays, 2, generator=g) * 384.0
    x = (px[:, 0:1] - 199.4) / 925.5
    y = (px[:, 1:2] - 198.1) / 922.6
    d = fwd + x * right + y * up2
    d = d / d.norm(dim=-1, keepdim=True)
    nears = torch.full((n_rays, 1), 0.5)
    fars = torch.full((n_rays, 1), 4.5)
    return o.contiguous(), d.contiguous(), idx, nears, fars


@torch.no_grad()
def perturb_field_(field, seed: int = 0, scale: float = 0.02, table_scale: float = 0.05):
    g = torch.Generator().manual_seed(seed)
    for name, p in field.named_parameters():
        if name.endswith("weight_v") or name.endswith(".bias"):
            p.add_((scale * torch.randn(p.shape, generator=g)).to(p.device))
    for l in range(field.num_layers - 1):
        lin = getattr(field, f"glin{l}")
        if hasattr(lin, "weight_g"):
            lin.weight_g.copy_(lin.weight_v.norm(dim=1, keepdim=True))
    t = field.encoding.table
    t.copy_(((torch.rand(t.shape, generator=g) * 2 - 1) * table_scale).to(t.device))
    return field

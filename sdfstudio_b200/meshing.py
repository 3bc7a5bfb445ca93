"""Dense-grid SDF evaluation for mesh extraction.

The reference's ``scripts/extract_mesh.py:97-126`` hands ``lambda x: field.forward_geonetwork(x)[:, 0]`` to
``utils/marching_cubes.py`` (:15-341), which materialises 512^3-point lattices on the host and evaluates them in 100 000-point
chunks.  Here the lattice is generated on the device chunk by chunk (sdfb200_lattice_points) and only the SDF head is evaluated
(the fused kernel's sdf-only mode when ``precision != "fp32"``), so a 512^3 block needs 0.5 GB for its result and nothing else.
Marching cubes itself (skimage / trimesh in the reference) stays outside the path.
"""
import ctypes as C
from typing import Callable, Sequence

import torch

from . import _lib


def sdf_fn(field, level: float = 0.0) -> Callable[[torch.Tensor], torch.Tensor]:
    """Drop-in for the ``sdf=`` callable of get_surface_sliding / get_surface_sliding_with_contraction: [N,3] -> [N]."""

    def fn(x: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            pts = _lib.f32c(x.reshape(-1, 3))
            s = field._run(pts, None, None, 1, ("sdf",), apply_contraction=False)["sdf"]
        return s - level if level != 0.0 else s

    return fn


def lattice_points(bbox_min: Sequence[float], bbox_max: Sequence[float], resolution, start: int, n: int, device) -> torch.Tensor:
    """[n,3] slice of the 'ij'-ordered np.linspace lattice (marching_cubes.py:49-56)."""
    lib = _lib.load()
    res = (resolution,) * 3 if isinstance(resolution, int) else tuple(int(r) for r in resolution)
    out = torch.empty(n, 3, device=device, dtype=torch.float32)
    mn = (C.c_double * 3)(*[float(v) for v in bbox_min])
    mx = (C.c_double * 3)(*[float(v) for v in bbox_max])
    rs = (C.c_int32 * 3)(*res)
    _lib.check(lib.sdfb200_lattice_points(mn, mx, rs, int(start), int(n), _lib.ptr(out), _lib.stream_ptr()), "sdfb200_lattice_points")
    return out


@torch.no_grad()
def evaluate_sdf_grid(field, resolution, bbox_min=(-1.0, -1.0, -1.0), bbox_max=(1.0, 1.0, 1.0), chunk: int = 1 << 22) -> torch.Tensor:
    """SDF on the dense lattice -> float32 tensor [rx, ry, rz] (what ``evaluate(points).reshape(N, N, N)`` is in the reference)."""
    res = (resolution,) * 3 if isinstance(resolution, int) else tuple(int(r) for r in resolution)
    total = res[0] * res[1] * res[2]
    dev = field.aabb.device
    out = torch.empty(total, device=dev, dtype=torch.float32)
    f = sdf_fn(field)
    for start in range(0, total, chunk):
        n = min(chunk, total - start)
        out[start:start + n] = f(lattice_points(bbox_min, bbox_max, res, start, n, dev))
    return out.view(*res)

"""Multi-GPU plumbing for the render path: rays are independent given the weights, so a frame's
ray list is cut into contiguous per-rank slices (instead of the reference's sequential 1024-ray chunk loop,
nerfstudio/models/base_model.py:165-189), every rank renders its slice with its own replica of the field, and the
slices are gathered on rank 0.  No data-path collective runs inside the hot path.  One process per GPU
(``torch.distributed``; NCCL on GPUs, gloo in the CPU tests)."""
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rays: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, end) slice of `n_rays` for `rank` (first `n_rays % world_size` ranks get one extra)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, extra = divmod(n_rays, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_ray_bundle(ray_bundle, rank: int, world_size: int):
    """Slice every per-ray tensor field of a (reference or local) RayBundle."""
    s, e = shard_bounds(ray_bundle.origins.shape[0], rank, world_size)
    return slice_ray_bundle(ray_bundle, s, e)


def slice_ray_bundle(ray_bundle, s: int, e: int):
    """rays [s, e) of a flat RayBundle (RayBundle.get_row_major_sliced_ray_bundle, cameras/rays.py:277-293)."""
    kw = {}
    for name in ("origins", "directions", "pixel_area", "directions_norm", "camera_indices", "nears", "fars", "times"):
        v = getattr(ray_bundle, name, None)
        kw[name] = None if v is None else v[s:e]
    return type(ray_bundle)(**{k: v for k, v in kw.items() if v is not None or k in ("directions_norm", "camera_indices", "nears", "fars")})


def flatten_ray_bundle(ray_bundle):
    """(flat bundle, image_shape | None).  The reference hands ``get_outputs_for_camera_ray_bundle`` a camera ray bundle of shape
    [H, W] (every field [H, W, k]; models/base_model.py:165-189 flattens it chunk by chunk with get_row_major_sliced_ray_bundle);
    a flat [N] bundle passes through unchanged."""
    o = ray_bundle.origins
    if o.dim() == 2:
        return ray_bundle, None
    image_shape = tuple(o.shape[:-1])
    kw = {}
    for name in ("origins", "directions", "pixel_area", "directions_norm", "camera_indices", "nears", "fars", "times"):
        v = getattr(ray_bundle, name, None)
        kw[name] = None if v is None else v.reshape(-1, v.shape[-1])
    return type(ray_bundle)(**{k: v for k, v in kw.items() if v is not None or k in ("directions_norm", "camera_indices", "nears", "fars")}), image_shape


def gather_outputs(outputs: Dict[str, torch.Tensor], n_rays: int, dst: int = 0) -> Optional[Dict[str, torch.Tensor]]:
    """Concatenate per-rank ray slices (in rank order) on `dst`; other ranks get None."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return outputs
    world, rank = dist.get_world_size(), dist.get_rank()
    result = {} if rank == dst else None
    sizes = [shard_bounds(n_rays, r, world) for r in range(world)]
    pad_to = max(e - s for s, e in sizes)
    for k in sorted(outputs):
        t = outputs[k].contiguous()
        if t.shape[0] < pad_to:  # collectives need equal shapes: pad the short slices, trim after the gather
            t = torch.cat([t, t.new_zeros((pad_to - t.shape[0], *t.shape[1:]))], dim=0)
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst)
        if rank == dst:
            result[k] = torch.cat([b[: e - s] for b, (s, e) in zip(bufs, sizes)], dim=0)
    return result


def max_over_ranks(value: float, device=None) -> float:
    """max-reduce a per-rank scalar (device timings are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

"""world_size-2 gloo test (CPU) of the ray-shard plumbing used by the multi-GPU render path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdfstudio_b200 import parallel
from sdfstudio_b200.rays import RayBundle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    full = RayBundle(origins=torch.rand(n_rays, 3, generator=g), directions=torch.rand(n_rays, 3, generator=g), pixel_area=torch.ones(n_rays, 1),
                     nears=torch.zeros(n_rays, 1), fars=torch.ones(n_rays, 1), camera_indices=torch.arange(n_rays).view(-1, 1))
    mine = parallel.shard_ray_bundle(full, rank, world)
    s, e = parallel.shard_bounds(n_rays, rank, world)
    assert mine.origins.shape[0] == e - s and torch.equal(mine.camera_indices[:, 0], torch.arange(s, e))
    # stand-in "render": a per-ray function of the ray only, so the gathered result must equal the unsharded one
    out = {"rgb": mine.origins * 2 + mine.directions, "depth": mine.origins.sum(-1, keepdim=True)}
    got = parallel.gather_outputs(out, n_rays, dst=0)
    t = parallel.max_over_ranks(float(rank + 1))
    assert t == float(world)
    if rank == 0:
        ok = torch.equal(got["rgb"], full.origins * 2 + full.directions) and torch.equal(got["depth"], full.origins.sum(-1, keepdim=True))
        q.put(bool(ok))
    else:
        assert got is None
    dist.destroy_process_group()


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 4096, 65537):
        for w in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_ray_shard_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


class _StubField(torch.nn.Module):  # SurfaceRenderer only stores it; get_outputs is replaced below
    pass


def _worker_image(rank, world, port, n_rays, q):
    """SurfaceRenderer.get_outputs_for_camera_ray_bundle(distributed=True): per-rank contiguous slices, chunked, gathered on rank 0."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdfstudio_b200 import NeuSSampler, SurfaceRenderer

    g = torch.Generator().manual_seed(0)
    full = RayBundle(origins=torch.rand(n_rays, 3, generator=g), directions=torch.rand(n_rays, 3, generator=g), pixel_area=torch.ones(n_rays, 1),
                     nears=torch.zeros(n_rays, 1), fars=torch.ones(n_rays, 1), camera_indices=torch.arange(n_rays).view(-1, 1))
    model = SurfaceRenderer(_StubField(), NeuSSampler(), kind="neus", eval_num_rays_per_chunk=97)
    calls = []

    def fake_outputs(bundle):  # stand-in for the CUDA path: a per-ray function of the ray only
        calls.append(bundle.origins.shape[0])
        return {"rgb": bundle.origins * 2 + bundle.directions, "depth": bundle.origins.sum(-1, keepdim=True),
                "normal": bundle.directions, "accumulation": bundle.origins[:, :1]}

    model.get_outputs = fake_outputs
    img = model.get_outputs_for_camera_ray_bundle(full, image_shape=(7, n_rays // 7), distributed=True)
    s, e = parallel.shard_bounds(n_rays, rank, world)
    assert sum(calls) == e - s and max(calls) <= 97
    if rank == 0:
        ok = torch.equal(img["rgb"].reshape(-1, 3), full.origins * 2 + full.directions) and img["depth"].shape == (7, n_rays // 7, 1)
        q.put(bool(ok))
    else:
        assert img is None
    dist.destroy_process_group()


def test_surface_renderer_image_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_image, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

"""Multi-GPU training check: DistributedDataParallel over SDFField, one process per GPU, NCCL gradient
all-reduce.  Each rank takes its contiguous ray shard (parallel.shard_ray_bundle); the averaged gradients must equal the
single-process gradients of the full batch, and an Adam step must leave all ranks with identical parameters.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_train_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdfstudio_b200 as sb  # noqa: E402
from sdfstudio_b200 import parallel, synthetic  # noqa: E402


def step_loss(field, bundle, target, S):
    with torch.no_grad():
        rs = sb.UniformSampler(num_samples=S, train_stratified=False).eval()(bundle)
    fo = field(rs, return_alphas=True)
    res = sb.render_from_alphas(fo[sb.FieldHeadNames.ALPHA], fo[sb.FieldHeadNames.RGB], fo[sb.FieldHeadNames.NORMAL], rs,
                                torch.ones(3, device=target.device), training=True)
    eik = ((fo[sb.FieldHeadNames.GRADIENT].norm(2, dim=-1) - 1) ** 2).mean()
    return (res["rgb"] - target).abs().mean() + 0.1 * eik


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl")
    R, S = 512 * world, 32
    cfg = sb.SDFFieldConfig(num_layers=2, num_layers_color=2, use_grid_feature=True, bias=0.5, beta_init=0.3, inside_outside=False,
                            grid_layout="torch", log2_hashmap_size=15)
    torch.manual_seed(0)
    field = sb.SDFField(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=49).to(dev)
    synthetic.perturb_field_(field, seed=7)
    field.train()
    o, d, cam, nears, fars = synthetic.dtu_like_rays(R, seed=5)
    bundle = sb.RayBundle(origins=o.to(dev), directions=d.to(dev), pixel_area=torch.ones(R, 1, device=dev), directions_norm=torch.ones(R, 1, device=dev),
                          camera_indices=cam.view(R, 1).to(dev), nears=nears.to(dev), fars=fars.to(dev))
    target = torch.rand(R, 3, generator=torch.Generator().manual_seed(9)).to(dev)

    # single-process reference gradients on the full batch
    field.zero_grad()
    step_loss(field, bundle, target, S).backward()
    ref = {k: p.grad.clone() for k, p in field.named_parameters() if p.grad is not None}

    out = {"world": world, "n_rays": R}
    if world > 1:
        ddp = torch.nn.parallel.DistributedDataParallel(field, device_ids=[local], find_unused_parameters=True)
        lo, hi = parallel.shard_bounds(R, rank, world)
        shard = parallel.shard_ray_bundle(bundle, rank, world)
        field.zero_grad()
        fo_loss = _ddp_loss(ddp, shard, target[lo:hi], S)
        fo_loss.backward()
        worst = 0.0
        for k, p in field.named_parameters():
            if p.grad is None or k not in ref:
                continue
            den = float(ref[k].abs().max())
            if den == 0:
                continue
            worst = max(worst, float((p.grad - ref[k]).abs().max()) / den)
        out["max_rel_grad_diff_vs_single_process"] = worst
        opt = torch.optim.Adam(field.parameters(), lr=1e-3)
        opt.step()
        flat = torch.cat([p.detach().reshape(-1).float() for p in field.parameters()])
        chk = torch.stack([flat.sum(), flat.abs().sum()])
        gathered = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(gathered, chk)
        out["params_identical_across_ranks"] = bool(all(torch.equal(g, gathered[0]) for g in gathered))
        ok = worst < 1e-3 and out["params_identical_across_ranks"]
        out["ok"] = ok
        if rank == 0:
            print(json.dumps(out))
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    print(json.dumps(out))


def _ddp_loss(ddp, shard, target, S):
    """same as step_loss, but the field call goes through the DDP wrapper so that its reducer hooks fire"""
    with torch.no_grad():
        rs = sb.UniformSampler(num_samples=S, train_stratified=False).eval()(shard)
    fo = ddp(rs, return_alphas=True)
    res = sb.render_from_alphas(fo[sb.FieldHeadNames.ALPHA], fo[sb.FieldHeadNames.RGB], fo[sb.FieldHeadNames.NORMAL], rs,
                                torch.ones(3, device=target.device), training=True)
    eik = ((fo[sb.FieldHeadNames.GRADIENT].norm(2, dim=-1) - 1) ** 2).mean()
    return (res["rgb"] - target).abs().mean() + 0.1 * eik


if __name__ == "__main__":
    main()

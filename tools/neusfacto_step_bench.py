"""Render-batch throughput of the actual neus-facto preset path (models/neus_facto.py:47-64, :138-170): ProposalNetworkSampler with
two HashMLPDensityField proposal networks (256 and 96 proposal samples per ray) -> 48 final samples -> SDFField -> alpha compositing.
Secondary measurement (bench.py holds the BASELINE metric, which fixes 128 uniform samples per ray); prints one JSON line.

    python tools/neusfacto_step_bench.py [--rays 4096] [--steps 20] [--warmup 5] [--precision bf16x3]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdfstudio_b200 as sb  # noqa: E402
from sdfstudio_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default="bf16x3")
    args = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
    cfg = sb.SDFFieldConfig(use_grid_feature=True, num_layers=2, num_layers_color=2, hidden_dim=256, bias=0.5, beta_init=0.3, inside_outside=False,
                            grid_layout="torch", precision=args.precision)
    field = synthetic.perturb_field_(sb.SDFField(cfg, aabb, num_images=49), 0).to(dev).eval()
    # neus_facto.py:52-59: two proposal networks, hidden 16, 5 levels, log2 T 17, max_res 64 / 256
    nets = []
    g = torch.Generator().manual_seed(1)
    for max_res in (64, 256):
        f = sb.HashMLPDensityField(aabb, num_layers=2, hidden_dim=16, num_levels=5, max_res=max_res, log2_hashmap_size=17).to(dev).eval()
        with torch.no_grad():
            nb = f.mlp_base
            nb.params[nb.n_net:] = ((torch.rand(nb.n_grid, generator=g) * 2 - 1) * 2.0).to(dev)
        nets.append(f)
    sampler = sb.ProposalNetworkSampler(num_proposal_samples_per_ray=(256, 96), num_nerf_samples_per_ray=48, num_proposal_network_iterations=2,
                                        use_uniform_sampler=True).eval()
    R = args.rays
    o, d, cam, nears, fars = synthetic.dtu_like_rays(R, 11)
    rb = sb.RayBundle(origins=o.to(dev), directions=d.to(dev), pixel_area=torch.ones(R, 1, device=dev), directions_norm=torch.ones(R, 1, device=dev),
                      camera_indices=cam.view(R, 1).to(dev), nears=nears.to(dev), fars=fars.to(dev))
    white = torch.ones(3, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fns = [n.density_fn for n in nets]

    @torch.no_grad()
    def step():
        rs, _, _ = sampler(rb, density_fns=fns)
        fo = field(rs, return_alphas=True)
        return sb.render_from_alphas(fo[sb.FieldHeadNames.ALPHA], fo[sb.FieldHeadNames.RGB], fo[sb.FieldHeadNames.NORMAL], rs, white, training=False)

    lib = sb._lib.load()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    total, launches0 = 0.0, lib.sdfb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.steps):
        flush.zero_()                          # L2 flush between timed iterations
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    ms = total / args.steps
    print(json.dumps({"metric": "rays/sec, neus-facto preset path (proposal 256+96 -> 48 samples)", "value": R / ms * 1e3, "unit": "rays/s", "rays": R,
                      "ms_per_step": ms, "proposal_evals_per_ray": 352, "final_samples_per_ray": 48, "precision": args.precision,
                      "gpu_launches_per_step": (lib.sdfb200_launch_count() - launches0) / args.steps, "l2_flush": "256 MiB write between steps"}))


if __name__ == "__main__":
    main()

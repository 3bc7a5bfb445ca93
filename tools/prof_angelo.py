"""torch.profiler table of one angelo-train-8192 step (bench.py --workload angelo-train-8192), single GPU."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import sdfstudio_b200 as sb
from sdfstudio_b200.synthetic import dtu_like_rays
import train_workload as tw
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
field = tw.make_angelo_field(dev, "bf16x3")
aabb = torch.tensor([[-1.0, -1, -1], [1, 1, 1]])
nets = [sb.HashMLPDensityField(aabb, num_layers=2, hidden_dim=16, num_levels=5, max_res=m, log2_hashmap_size=17).to(dev).eval() for m in (64, 256)]
fns = [n.density_fn for n in nets]
sampler = sb.ProposalNetworkSampler(num_proposal_samples_per_ray=(256, 96), num_nerf_samples_per_ray=48, num_proposal_network_iterations=2, use_uniform_sampler=False).train()
opt = torch.optim.Adam(field.parameters(), lr=5e-4, eps=1e-15)
R = 8192
o, d, cam, nears, fars = dtu_like_rays(R, 11)
rb = sb.RayBundle(origins=o.to(dev), directions=d.to(dev), pixel_area=torch.ones(R, 1, device=dev), directions_norm=torch.ones(R, 1, device=dev), camera_indices=cam.view(R, 1).to(dev), nears=nears.to(dev), fars=fars.to(dev))
target = torch.rand(R, 3, device=dev); white = torch.ones(3, device=dev)
def step():
    with torch.no_grad():
        rs, _, _ = sampler(rb, density_fns=fns)
    fo = field(rs, return_alphas=True)
    out = sb.render_from_alphas(fo[sb.FieldHeadNames.ALPHA], fo[sb.FieldHeadNames.RGB], fo[sb.FieldHeadNames.NORMAL], rs, white, training=True)
    eik = ((fo[sb.FieldHeadNames.GRADIENT].norm(2, dim=-1) - 1) ** 2).mean()
    loss = (out["rgb"] - target).abs().mean() + 0.1 * eik
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))

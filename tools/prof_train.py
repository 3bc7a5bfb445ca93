"""torch.profiler table of one neus-facto training step (8192 rays x 128 samples); see DESIGN.md section 3b.  argv[1]: tc | aten"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdfstudio_b200 as sb
from sdfstudio_b200 import synthetic
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
cfg = sb.SDFFieldConfig(use_grid_feature=True, num_layers=2, num_layers_color=2, hidden_dim=256, bias=0.5, beta_init=0.3, inside_outside=False, grid_layout="torch", precision="bf16x3",
                        train_gemm=sys.argv[1] if len(sys.argv) > 1 else "auto")
field = synthetic.perturb_field_(sb.SDFField(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=49), 0).to(dev).train()
sampler = sb.NeuSSampler(num_samples=64, num_samples_importance=64, num_samples_outside=0, num_upsample_steps=4).train()
opt = torch.optim.Adam(field.parameters(), lr=5e-4, eps=1e-15)
R = 8192
o, d, cam, nears, fars = synthetic.dtu_like_rays(R, 11)
rb = sb.RayBundle(origins=o.to(dev), directions=d.to(dev), pixel_area=torch.ones(R, 1, device=dev), directions_norm=torch.ones(R, 1, device=dev), camera_indices=cam.view(R, 1).to(dev), nears=nears.to(dev), fars=fars.to(dev))
target = torch.rand(R, 3, device=dev); white = torch.ones(3, device=dev)
def step():
    with torch.no_grad():
        rs = sampler(rb, sdf_fn=field.get_sdf)
    fo = field(rs, return_alphas=True)
    out = sb.render_from_alphas(fo[sb.FieldHeadNames.ALPHA], fo[sb.FieldHeadNames.RGB], fo[sb.FieldHeadNames.NORMAL], rs, white, training=True)
    eik = ((fo[sb.FieldHeadNames.GRADIENT].norm(2, dim=-1) - 1) ** 2).mean()
    loss = (out["rgb"] - target).abs().mean() + 0.1 * eik
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=90))

import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdfstudio_b200 as sb
from sdfstudio_b200 import synthetic
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
R, S = 4096, 128
o, d, cam, nears, fars = synthetic.dtu_like_rays(R, 11)
rb = sb.RayBundle(origins=o.to(dev), directions=d.to(dev), pixel_area=torch.ones(R, 1, device=dev), directions_norm=torch.ones(R, 1, device=dev), camera_indices=cam.view(R, 1).to(dev), nears=nears.to(dev), fars=fars.to(dev))
rs = sb.UniformSampler(num_samples=S).eval()(rb)
for prec in ("bf16x3",):
    cfg = sb.SDFFieldConfig(grid_layout="torch", precision=prec, num_layers=8, num_layers_color=4, hidden_dim=256, use_grid_feature=False, bias=0.8, beta_init=0.1, inside_outside=True)
    field = synthetic.perturb_field_(sb.SDFField(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=49), 0).to(dev).eval()
    with torch.no_grad():
        for _ in range(2): field(rs, return_alphas=True)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            field(rs, return_alphas=True); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))

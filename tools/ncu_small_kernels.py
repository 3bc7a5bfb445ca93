"""Driver for ncu captures of the non-tensor kernels of the path (hash-grid operator forward / backward at the headline and the angelo
table shapes, the error-bounded sampler's inner step, the compositing kernels).  Run under
    ncu --set full --clock-control none -k regex:'k_grid_encode|k_volsdf_step|k_render|k_weights' -c 12 -o gpurun_out/r02_small python tools/ncu_small_kernels.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sdfstudio_b200 as sb  # noqa: E402
from sdfstudio_b200.synthetic import dtu_like_rays  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
N = 4096 * 128
x = torch.rand(N, 3, device=dev)
# headline table: L16 F2 T2^19 torch layout (67 MB); angelo: L16 F8 T2^22 (2.1 GB)
for (F, log2t, base, mx) in ((2, 19, 16, 2048), (8, 22, 64, 4096)):
    enc = sb.HashEncoding(num_levels=16, min_res=base, max_res=mx, log2_hashmap_size=log2t, features_per_level=F).to(dev)
    xr = x.clone().requires_grad_(True)
    out = enc(xr)                                    # k_grid_encode
    g = torch.randn_like(out)
    out.backward(g)                                  # k_grid_encode_bwd (table scatter + dx)
    torch.cuda.synchronize()
    del enc, out, g, xr
    torch.cuda.empty_cache()

# error-bounded sampler + compositing on the volsdf-shaped field
cfg = sb.SDFFieldConfig(use_grid_feature=True, num_layers=2, num_layers_color=2, hidden_dim=256, bias=0.5, beta_init=0.1, inside_outside=False,
                        grid_layout="torch", precision="bf16x3")
field = sb.SDFField(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=49).to(dev).eval()
R = 4096
o, d, cam, nears, fars = dtu_like_rays(R, 3)
rb = sb.RayBundle(origins=o.to(dev), directions=d.to(dev), pixel_area=torch.ones(R, 1, device=dev), directions_norm=torch.ones(R, 1, device=dev),
                  camera_indices=cam.view(R, 1).to(dev), nears=nears.to(dev), fars=fars.to(dev))
with torch.no_grad():
    smp = sb.ErrorBoundedSampler(num_samples=64, num_samples_eval=128, num_samples_extra=32, eps=0.1, beta_iters=10, max_total_iters=5).eval()
    rs, _ = smp(rb, density_fn=field.laplace_density, sdf_fn=field.get_sdf)
    fo = field(rs)
    w, T = rs.get_weights_and_transmittance(fo[sb.FieldHeadNames.DENSITY])
    img = sb.render_all(w, fo[sb.FieldHeadNames.RGB], fo[sb.FieldHeadNames.NORMAL], rs, torch.ones(3, device=dev))
torch.cuda.synchronize()
print("done")

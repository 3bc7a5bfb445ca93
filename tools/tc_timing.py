"""Per-phase cycle breakdown of the fused tensor-core field kernel (CTA 0, first tiles).  SDFB200_TC_TIMING=1 is set here."""
import ctypes
import os
import sys

os.environ["SDFB200_TC_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import sdfstudio_b200 as sb  # noqa: E402
from sdfstudio_b200.synthetic import dtu_like_rays  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
dev = torch.device("cuda")
log2t = int(sys.argv[2]) if len(sys.argv) > 2 else 19
tdt = sys.argv[3] if len(sys.argv) > 3 else "fp32"
torch.manual_seed(0)
cfg = sb.SDFFieldConfig(use_grid_feature=True, num_layers=2, num_layers_color=2, hidden_dim=256, bias=0.5, beta_init=0.3, inside_outside=False,
                        log2_hashmap_size=log2t, grid_layout="torch", precision=prec, table_dtype=tdt)
from sdfstudio_b200.synthetic import perturb_field_  # noqa: E402
field = perturb_field_(sb.SDFField(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=49), 0).to(dev).eval()
o, d, cam, nears, fars = dtu_like_rays(4096, 1000)
rb = sb.RayBundle(origins=o.to(dev), directions=d.to(dev), pixel_area=torch.ones(4096, 1, device=dev), directions_norm=torch.ones(4096, 1, device=dev),
                  camera_indices=cam.view(-1, 1).to(dev), nears=nears.to(dev), fars=fars.to(dev))
rs = sb.UniformSampler(num_samples=128).eval()(rb)
with torch.no_grad():
    for _ in range(3):
        field(rs, return_alphas=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
sb._lib.check(sb._lib.load().sdfb200_debug_tc_timing(buf))
names = ["start", "wait G0", "E0", "wait G1", "E1", "slice2", "-", "wait B1", "EB1", "wait B0", "EB0", "wait C0", "EC0", "wait C1", "EC1+heads"]
for t in (1, 2, 5, 10):
    st = [buf[t * 32 + k] for k in range(16)]
    tot = st[15] - st[0]
    print(f"tile {t}: total {tot} cycles  " + "  ".join(f"{n} {st[i+1]-st[i]}" for i, n in enumerate(names)))

"""Per-phase cycle breakdown of the fused tensor-core field kernel (CTA 0, first tiles): epilogue thread, gather thread and MMA
issuer stamps.  Uses the timing build of the kernel inside libsdfb200_dbg.so (sdfstudio_b200/build.py)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sdfstudio_b200 import _lib  # noqa: E402

variant = sys.argv[5] if len(sys.argv) > 5 else None
_lib.LIB_PATH = os.path.join(_lib.HERE, f"libsdfb200_v{variant}.so" if variant else "libsdfb200_dbg.so")     # same ABI + the stamped kernel
import sdfstudio_b200 as sb  # noqa: E402
from sdfstudio_b200.synthetic import dtu_like_rays, perturb_field_  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
dev = torch.device("cuda")
log2t = int(sys.argv[2]) if len(sys.argv) > 2 else 19
tdt = sys.argv[3] if len(sys.argv) > 3 else "fp32"
fused = (sys.argv[4] if len(sys.argv) > 4 else "fused") == "fused"
torch.manual_seed(0)
cfg = sb.SDFFieldConfig(use_grid_feature=True, num_layers=2, num_layers_color=2, hidden_dim=256, bias=0.5, beta_init=0.3, inside_outside=False,
                        log2_hashmap_size=log2t, grid_layout="torch", precision=prec, table_dtype=tdt)
field = perturb_field_(sb.SDFField(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=49), 0).to(dev).eval()
o, d, cam, nears, fars = dtu_like_rays(4096, 1000)
rb = sb.RayBundle(origins=o.to(dev), directions=d.to(dev), pixel_area=torch.ones(4096, 1, device=dev), directions_norm=torch.ones(4096, 1, device=dev),
                  camera_indices=cam.view(-1, 1).to(dev), nears=nears.to(dev), fars=fars.to(dev))
rs = sb.UniformSampler(num_samples=128).eval()(rb)
with torch.no_grad():
    for _ in range(3):
        if fused:
            field.render(rs, torch.ones(3, device=dev))
        else:
            field(rs, return_alphas=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ms = []
with torch.no_grad():
    for _ in range(10):
        flush.zero_()
        e0.record()
        field.render(rs, torch.ones(3, device=dev)) if fused else field(rs, return_alphas=True)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
print(f"variant {variant}: field call {sorted(ms)[len(ms)//2]:.3f} ms (median of 10, L2 flushed)")
lib = _lib.load()
lib.sdfb200_debug_tc_timing.argtypes = [ctypes.c_void_p]
buf = (ctypes.c_longlong * 512)()
assert lib.sdfb200_debug_tc_timing(buf) == 0
names = ["start", "wait G0", "E0", "wait G1", "E1", "-", "-", "wait B1", "EB1", "wait B0", "EB0", "wait C0", "EC0", "wait C1", "EC1+heads"]
for t in (5, 10):
    st = [buf[t * 32 + k] for k in range(32)]
    tot = st[15] - st[0]
    print(f"tile {t}: epilogue total {tot} cycles  " + "  ".join(f"{n} {st[i+1]-st[i]}" for i, n in enumerate(names) if n != "-"))
    print(f"   gather: wait g0done {st[17]-st[16]}  colour cols {st[18]-st[17]}  encode next {st[19]-st[18]}   (loop start at +{st[16]-st[0]} of the epilogue's tile start)")
    print("   mma issue start (relative to epilogue tile start): " + "  ".join(f"L{L} +{st[20+L]-st[0]}" for L in range(7)) + f"  done issuing +{st[27]-st[0]}")

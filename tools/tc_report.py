"""Prints the error of each precision mode of the field against the CPU oracle (fp32 and fp64) on seeded cases."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_case, make_bundle, oracle64, rel_err  # noqa: E402
from oracle import render, samplers  # noqa: E402

import sdfstudio_b200 as sb  # noqa: E402


def main():
    H = sb.FieldHeadNames
    for name in ("neusfacto_c1", "neusfacto_c1_init"):
        for prec in ("fp32", "bf16x3", "bf16"):
            spec, kw, o, d, cam, nears, fars, oracle, field = build_case(name, precision=prec)
            rb = make_bundle(o, d, cam, nears, fars)
            rs = sb.UniformSampler(num_samples=kw["S"]).eval()(rb)
            out = field(rs, return_alphas=True, return_occupancy=True)
            torch.cuda.synchronize()
            eu = sb.rays.bins_of(rs).cpu().double()
            o64 = oracle64(spec, oracle.p, kw)
            e = o64.get_outputs(o.double(), d.double(), eu[:, :-1], eu[:, 1:] - eu[:, :-1], cam, return_alphas=True, return_occupancy=True)
            w = rs.get_weights_from_alphas(out[H.ALPHA])
            img = sb.render_all(w, out[H.RGB], out[H.NORMAL], rs, torch.ones(3, device="cuda"))
            ow, _ = samplers.weights_from_alphas(e["alphas"][..., 0])
            orgb = render.render_rgb(e["rgb"], ow[..., None], torch.ones(3, dtype=torch.float64))
            odep = render.render_depth(ow[..., None], eu[:, :-1, None], eu[:, 1:, None], "expected")
            gs = float(e["gradients"].abs().max())
            mse = float(((img["rgb"].cpu().double() - orgb) ** 2).mean())
            psnr = -10 * torch.log10(torch.tensor(max(mse, 1e-30)))
            sdf_only = field.get_sdf(rs)
            e_sdf_u = o64.get_sdf(o.double(), d.double(), eu[:, :-1])
            print(f"{name:18s} {prec:7s} sdf {rel_err(out[H.SDF], e['sdf']):.2e} get_sdf {rel_err(sdf_only[..., 0], e_sdf_u):.2e} grad/gs {float((out[H.GRADIENT].cpu().double()-e['gradients']).abs().max())/gs:.2e} "
                  f"rgb {rel_err(out[H.RGB], e['rgb'], 1e-2):.2e} alpha {rel_err(out[H.ALPHA], e['alphas'], 1e-2):.2e} dens {rel_err(out[H.DENSITY], e['density'], 1e-2):.2e} | "
                  f"render rgb {rel_err(img['rgb'], orgb, 1e-2):.2e} depth {rel_err(img['depth'], odep, 1e-2):.2e} psnr {float(psnr):.1f} dB")


if __name__ == "__main__":
    main()

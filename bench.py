#!/usr/bin/env python
"""bench.py -- rays/sec of the SDF volume-rendering hot path (BASELINE.json metric).

Workload (config.workload = "neus-facto-dtu65-4096x128", BASELINE.json configs[1]): DTU-scan65-shaped synthetic rays,
4096 rays x 128 samples per GPU per step, neus-facto SDFField (hash L=16 F=2 T=2^19, geo MLP 71-256-256-257, colour MLP
321-256-256-3, random-init weights).  One step = UniformSampler(128) -> SDFField.forward(return_alphas) ->
weights-from-alphas -> RGB/depth/normal/accumulation renderers, i.e. the render-only pass of
SurfaceModel.get_outputs (models/base_surface_model.py:292-365).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16x3|bf16]

value   : whole-job rays/s, inputs resident in HBM, CUDA-event time (max over ranks), L2 flushed between steps.
e2e     : same metric through the public module API with HOST (pinned) ray buffers: H2D of the rays and D2H of the
          rendered rgb/depth/normal/accumulation inside the timed region.
roofline: the field kernel(s) against the measured bf16 tensor peak (MEASURED_PEAKS.json).
cpu_baseline / --impl reference: the CPU oracle port (oracle/, a restatement of the reference's torch-CPU path) on a
          bounded sample of the same workload, all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_PER_GPU, S = 4096, 128
FLOP_PER_SAMPLE = 763904  # 381 952 MAC (geo 149 504 + grad 83 968 + colour 148 480)
WORKLOAD = "neus-facto-dtu65-4096x128"


def make_field(device, precision="fp32", seed=0, table_dtype="fp32"):
    """The product SDFField for the workload (neus-facto preset, method_configs.py:472-480 + README override
    inside_outside=False), random-init + perturbation so that hash + PE inputs matter."""
    import sdfstudio_b200 as sb
    from sdfstudio_b200.synthetic import perturb_field_

    torch.manual_seed(seed)
    cfg = sb.SDFFieldConfig(use_grid_feature=True, num_layers=2, num_layers_color=2, hidden_dim=256, bias=0.5, beta_init=0.3,
                            use_appearance_embedding=False, inside_outside=False, grid_layout="torch", precision=precision, table_dtype=table_dtype)
    field = sb.SDFField(cfg, torch.tensor([[-1.0, -1, -1], [1, 1, 1]]), num_images=49)
    perturb_field_(field, seed)
    return field.to(device).eval()


def oracle_of(field):
    """CPU oracle holding the SAME parameters as `field` (cpu_baseline / reference arm only)."""
    from oracle.field import FieldSpec, OracleField

    spec = FieldSpec(num_layers=2, num_layers_color=2, hidden_dim=256, use_grid_feature=True, grid_layout="torch")
    sd = {k: v.detach().cpu() for k, v in field.state_dict().items()}
    sd["hash_table"] = sd.pop("encoding.hash_table")
    return OracleField(spec, sd)


def read_traffic(precision):
    """dram__bytes_read.sum + dram__bytes_write.sum of the field kernel from the committed `ncu --set full` capture."""
    p = os.path.join(ROOT, "profiles", f"r01_ncu_field_tc_{precision}_summary.json")
    if not os.path.exists(p):
        return None
    try:
        with open(p) as fh:
            d = json.load(fh)
        def mb(k):
            v, u = float(d[k]["value"]), d[k]["unit"]
            return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[u]
        return mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    except Exception:  # noqa: BLE001
        return None


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"bf16_tflops": d.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_step(oracle, o, d, cam, nears, fars, S_):
    """The same step on the CPU through the oracle port (what the reference's torch-CPU path computes)."""
    from oracle import render, samplers

    b = samplers.spaced_sampler(nears, fars, S_, "uniform")
    out = oracle.get_outputs(o, d, b.starts, b.deltas, cam, return_alphas=True)
    w, _ = samplers.weights_from_alphas(out["alphas"][..., 0])
    w = w[..., None]
    rgb = render.render_rgb(out["rgb"], w, torch.ones(3))
    depth = render.render_depth(w, b.starts[..., None], b.ends[..., None], "expected")
    normal = render.render_semantics(out["normals"], w)
    acc = render.render_accumulation(w)
    return rgb, depth, normal, acc


def cpu_arm(rays_per_step, steps, warmup):
    """rays/s of the CPU oracle port on all host threads, bounded sample of the workload."""
    from sdfstudio_b200.synthetic import dtu_like_rays

    ncpu = os.cpu_count() or 1
    oracle = oracle_of(make_field("cpu"))
    o, d, cam, nears, fars = dtu_like_rays(rays_per_step, 4242)
    # torch-CPU does not scale monotonically with threads on small batches: pick the fastest thread count (<= all cores)
    best, cores = None, ncpu
    for t in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(t)
        oracle_step(oracle, o, d, cam, nears, fars, S)
        t0 = time.perf_counter()
        oracle_step(oracle, o, d, cam, nears, fars, S)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, cores = dt, t
    torch.set_num_threads(cores)
    for _ in range(warmup):
        oracle_step(oracle, o, d, cam, nears, fars, S)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_step(oracle, o, d, cam, nears, fars, S)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return rays_per_step / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rays = 256  # bounded sample: 256 rays x 128 samples per step (~0.3 s of CPU work per step)
    steps = max(1, min(args.steps, 20))
    value, dt, cores = cpu_arm(rays, steps, min(args.warmup, 3))
    line = {
        "impl": "reference", "metric": "rays/sec at 4096 rays x 128 samples", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 3), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": f"{rays} rays x {S} samples per step"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": f"{rays} rays x {S} samples per step, {steps} steps"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }  # fmt: skip
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SDFB200_PRECISION", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="A/B: separate field and compositing launches (per-sample heads through HBM)")
    ap.add_argument("--table-dtype", default="fp32", choices=["fp32", "fp16"], help="fp16 = gather from a half-precision copy (tiny-cuda-nn's storage)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    import sdfstudio_b200 as sb
    from sdfstudio_b200.synthetic import dtu_like_rays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path in the product)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    precision = args.precision
    if precision == "auto":
        precision = "bf16x3"   # tcgen05 path at parity-grade precision (bf16 split, fp32 accumulate); fp32 / bf16 via --precision
    field = make_field(dev, precision, table_dtype=args.table_dtype)
    sampler = sb.UniformSampler(num_samples=S).eval()
    white = torch.ones(3, device=dev)
    H = sb.FieldHeadNames

    # ray shard of this rank (independent rays: no data-path collective)
    o, d, cam, nears, fars = dtu_like_rays(R_PER_GPU, 1000 + rank)
    host = [t.pin_memory() for t in (o, d, nears, fars)]
    cam_d = cam.view(-1, 1).to(dev)
    pix = torch.ones(R_PER_GPU, 1, device=dev)
    dev_in = [t.to(dev) for t in host]

    ev = lambda: torch.cuda.Event(enable_timing=True)
    field_ms = []

    def step(o_, d_, n_, f_, time_field=False):
        rb = sb.RayBundle(origins=o_, directions=d_, pixel_area=pix, directions_norm=pix, camera_indices=cam_d, nears=n_, fars=f_)
        rs = sampler(rb)
        if time_field:
            e0, e1 = ev(), ev()
            e0.record()
        if args.unfused:
            out = field(rs, return_alphas=True)
            if time_field:
                e1.record()
                field_ms.append((e0, e1))
            return sb.render_from_alphas(out[H.ALPHA], out[H.RGB], out[H.NORMAL], rs, white)
        res = field.render(rs, white)          # field + compositing: one fused launch (+ the global depth clip)
        if time_field:
            e1.record()
            field_ms.append((e0, e1))
        return res

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step(*dev_in)
        barrier()
        launches0 = sb._lib.launch_count()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        # ---- device-resident timing ----
        pairs = []
        barrier()
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = ev(), ev()
            e0.record()
            step(*dev_in, time_field=True)
            e1.record()
            pairs.append((e0, e1))
        barrier()
        launches = sb._lib.launch_count() - launches0
        dev_ms = sum(a.elapsed_time(b) for a, b in pairs)
        fld_ms = sum(a.elapsed_time(b) for a, b in field_ms)
        # ---- end-to-end timing: host rays in, rendered images out ----
        out_host = {k: torch.empty(s, dtype=torch.float32).pin_memory() for k, s in (("rgb", (R_PER_GPU, 3)), ("depth", (R_PER_GPU, 1)),
                                                                                       ("normal", (R_PER_GPU, 3)), ("accumulation", (R_PER_GPU, 1)))}
        for _ in range(2):
            r = step(*[t.to(dev, non_blocking=True) for t in host])
        barrier()
        pairs2 = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = ev(), ev()
            e0.record()
            r = step(*[t.to(dev, non_blocking=True) for t in host])
            for k, t in out_host.items():
                t.copy_(r[k], non_blocking=True)
            e1.record()
            pairs2.append((e0, e1))
        barrier()
        e2e_ms = sum(a.elapsed_time(b) for a, b in pairs2)
        clk = clocks.stop() if rank == 0 else None

    times = torch.tensor([dev_ms, e2e_ms, fld_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, fld_ms = (float(x) for x in times)
    if rank == 0:
        peaks = read_peaks()
        rays_total = R_PER_GPU * world * args.steps
        value = rays_total / (dev_ms * 1e-3)
        e2e = rays_total / (e2e_ms * 1e-3)
        flop_per_launch = FLOP_PER_SAMPLE * R_PER_GPU * S
        achieved_tflops = flop_per_launch / (fld_ms / args.steps * 1e-3) / 1e12
        h2d = sum(t.numel() * t.element_size() for t in host)
        d2h = sum(t.numel() * t.element_size() for t in out_host.values())
        cpu = None
        if not args.no_cpu_baseline:
            cv, cdt, cores = cpu_arm(256, 8, 2)
            cpu = {"value": cv, "unit": "rays/s", "cores": cores, "kind": "port", "sample": f"256 rays x {S} samples per step, 8 steps ({cdt*1e3:.0f} ms/step)"}
        line = {
            "metric": "rays/sec at 4096 rays x 128 samples", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (fp32 accumulate)", "bf16": "bf16 (fp32 accumulate)"}[precision], "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_gpu": R_PER_GPU, "samples_per_ray": S, "sampler": "UniformSampler(128), eval", "field": f"neus-facto SDFField L16 F2 T2^19 MLP 2x256 (torch-layout table, {args.table_dtype})",
                       "precision": precision, "l2": "flushed between timed steps (256 MiB write)", "parallelism": f"ray-shard x{world}, no data-path collective"},
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": achieved_tflops / peaks["bf16_tflops_sustained"], "traffic": read_traffic(precision), "peak_source": peaks["source"] + " bf16_tflops_sustained",
                         "kernel": "k_field_tc (sdfb200_field_forward)" if precision != "fp32" else "sdfb200_field_forward (k_sgemm + elementwise kernels)", "ms_per_launch": fld_ms / args.steps,
                         "algorithmic_flop_per_launch": flop_per_launch},
            "cpu_baseline": cpu, "clocks": clk,
        }  # fmt: skip
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

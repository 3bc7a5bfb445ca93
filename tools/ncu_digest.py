"""Digest of an `ncu --set full` report read on the CPU box: headline metrics, stall mix, SASS hot spots bucketed by source line.
    python tools/ncu_digest.py gpurun_out/x.ncu-rep [--json out.json]"""
import collections
import csv
import json
import re
import subprocess
import sys

rep = sys.argv[1]
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
out = {}
for i, h in enumerate(hdr):
    if h in want:
        out[h] = {"value": vals[i], "unit": units[i]}
stalls = {}
for i, h in enumerate(hdr):
    m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", h)
    if m and not m.group(1).endswith("not_issued"):
        try:
            stalls[m.group(1)] = int(float(vals[i]))
        except ValueError:
            pass
tot = sum(stalls.values()) or 1
out["stall_mix_pct"] = {k: round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v * 100 > tot}
for k, v in out.items():
    print(k, v)
src = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h2 = src[1]
ix = {h: i for i, h in enumerate(h2)}
data = src[2:]
print("SASS instructions:", len(data))
if "--json" in sys.argv:
    with open(sys.argv[sys.argv.index("--json") + 1], "w") as fh:
        json.dump(out, fh, indent=1)
# hot spots: buckets of 250 instructions with their dominant stall reasons and opcodes
keys = [k for k in h2 if k.startswith("stall_") and "Not Issued" not in k]
for b in range(0, len(data), 250):
    seg = data[b:b + 250]
    t = sum(int(r[ix["# Samples"]]) for r in seg)
    if t * 60 < sum(int(r[ix["# Samples"]]) for r in data):
        continue
    d = {k: sum(int(r[ix[k]]) for r in seg) for k in keys}
    top = sorted(d.items(), key=lambda kv: -kv[1])[:3]
    ops = collections.Counter()
    for r in seg:
        for w in r[ix["Source"]].split():
            if w.startswith(("LDTM", "STTM", "UTC", "LDG", "STG", "MUFU", "SHFL", "BAR", "SYNCS", "UBLKCP", "ATOMS", "F2I", "I2F", "CALL", "LDL", "STL")):
                ops[w.split(".")[0]] += 1
    print(b, t, top, dict(ops))

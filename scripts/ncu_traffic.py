"""Turns an ncu per-launch metrics CSV (scripts/gpu_r2_final.sh) + the bench line of the same command into the record
bench.py's roofline uses: warp instructions, ALU / FMA-pipe instructions and DRAM bytes PER RAY of one kernel of one
workload, stamped with the commit and the files they were read from.  usage:
  ncu_traffic.py <workload> <kernel-key> <kernel-name-regex> <metrics.csv> <rays-per-render> <renders-captured> <source-note> [<separator-regex> <segment> [<max-launches>]]
With a separator regex the launch list is cut into segments at every launch whose name matches it (bench.py --workload c3 runs one
COUNTING launch before the timed launches of each ray batch) and only the launches of segment <segment> (0-based) are used."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload, key, name_re, csv_path, rays, renders, note = sys.argv[1:8]
sep_re, segment = (sys.argv[8], int(sys.argv[9])) if len(sys.argv) > 9 else (None, 0)
max_launches = int(sys.argv[10]) if len(sys.argv) > 10 else 0   # only the first launches of the segment (what follows is another input)
rays, renders = float(rays), float(renders)
rows = []
with open(csv_path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = list(csv.reader(lines))
hdr = r[0]
iname, imetric, ivalue, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
iid = hdr.index("ID")
acc, launches = {}, set()
pct, dur = {}, {}
seg, last_sep_id = -1 if sep_re else 0, None
for row in r[1:]:
    if len(row) <= ivalue:
        continue
    if sep_re and re.search(sep_re, row[iname]):
        if row[iid] != last_sep_id:
            seg += 1; last_sep_id = row[iid]
        continue
    if seg != segment or not re.search(name_re, row[iname]):
        continue
    if max_launches and row[iid] not in launches and len(launches) >= max_launches:
        continue
    launches.add(row[iid])
    v = float(row[ivalue].replace(",", ""))
    u = row[iunit]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    m = row[imetric]
    if m == "gpu__time_duration.sum":          # ("du-ratio-n" is not a ratio)
        dur[row[iid]] = v * scale
        acc[m] = acc.get(m, 0.0) + v * scale
    elif "pct" in m or m.endswith(".ratio"):
        pct.setdefault(m, []).append((row[iid], v))
    else:
        acc[m] = acc.get(m, 0.0) + v * scale
n = max(1, len(launches))
# utilisations as DURATION-weighted means over the launches (the short late-bounce launches of a render would otherwise count
# like the long first ones)
mean = {m: sum(dur.get(i, 0.0) * v for i, v in vs) / max(sum(dur.get(i, 0.0) for i, _ in vs), 1e-30) for m, vs in pct.items()}
lsu = mean.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 0.0)
issue = mean.get("sm__inst_issued.avg.pct_of_peak_sustained_active", 0.0)
alu = mean.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
# the limiter ncu names: the busier of instruction issue and the L1 data pipe (DRAM stays below 25 % everywhere)
bound = "l1-lsu" if lsu > issue else ("sm-issue (ALU pipe)" if alu > 0.8 * issue else "sm-issue")
if renders == 0:            # one launch = one pass over the batch (bench.py --workload c3)
    renders = float(n)
total_rays = rays * renders
rec = {
    "bound": bound,
    "warp_inst_per_ray": acc.get("smsp__inst_executed.sum", 0.0) / total_rays,
    "alu_pipe_inst_per_ray": acc.get("sm__inst_executed_pipe_alu.sum", 0.0) / total_rays,
    "fma_pipe_inst_per_ray": acc.get("sm__inst_executed_pipe_fma.sum", 0.0) / total_rays,
    "lsu_wavefronts_per_ray": acc.get("l1tex__data_pipe_lsu_wavefronts.sum", 0.0) / total_rays,
    "dram_bytes_per_ray": (acc.get("dram__bytes_read.sum", 0.0) + acc.get("dram__bytes_write.sum", 0.0)) / total_rays,
    "launches_captured": n, "rays_captured": total_rays,
    "gpu_time_under_ncu_ms": acc.get("gpu__time_duration.sum", 0.0) * 1e3,
    "pct_mean": mean,
    "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
    "source": note,
}
p = os.path.join(ROOT, "profiles", "traffic.json")
d = json.load(open(p))
d.setdefault(workload, {})[key] = rec
json.dump(d, open(p, "w"), indent=1)
print(workload, key, json.dumps({k: rec[k] for k in ("warp_inst_per_ray", "alu_pipe_inst_per_ray", "fma_pipe_inst_per_ray", "dram_bytes_per_ray", "launches_captured")}))

"""Per-source-line summary of ONE launch of an ncu report captured with --import-source on.
usage: python scripts/ncu_source_launch.py report.ncu-rep <launch index in the report> [top N] [sort: smp|inst]"""
import collections
import csv
import io
import subprocess
import sys

rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
sort = sys.argv[4] if len(sys.argv) > 4 else "inst"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
fn = hdr = cur = None
d = {}
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
        ii, it, ism = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    elif hdr and r[0].isdigit():
        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        e = d.setdefault((cur, int(r[0]), r[1].strip()[:110]), [0, 0, 0])
        e[0] += num(r[ii]); e[1] += num(r[it]); e[2] += num(r[ism])
print(fn[:150])
ti = sum(e[0] for e in d.values()) or 1
tt = sum(e[1] for e in d.values())
ts = sum(e[2] for e in d.values()) or 1
print("warp inst %.3e  thread inst %.3e  threads/inst %.2f  samples %d" % (ti, tt, tt / ti, ts))
bf = collections.defaultdict(lambda: [0, 0, 0])
for (f, l, s), e in d.items():
    for k in range(3):
        bf[f][k] += e[k]
for f, e in sorted(bf.items(), key=lambda kv: -kv[1][0]):
    print("  %-30s inst %5.1f%%  smp %5.1f%%  thr %5.1f" % (f, 100 * e[0] / ti, 100 * e[2] / ts, e[1] / max(e[0], 1)))
k = 2 if sort == "smp" else 0
for (f, l, s), e in sorted(d.items(), key=lambda kv: -kv[1][k])[:top]:
    print("%5.1f%% inst %5.1f%% smp thr %4.1f %s:%d %s" % (100 * e[0] / ti, 100 * e[2] / ts, e[1] / max(e[0], 1), f, l, s))

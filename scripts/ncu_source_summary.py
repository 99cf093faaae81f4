"""Per-source-line summary of an ncu report captured with --import-source on.
usage: python scripts/ncu_source_summary.py report.ncu-rep [kernel-substring] [top N]
Runs `ncu -i ... --page source --csv --print-source cuda,sass` and prints, per CUDA source line:
warp instructions executed, share, average active threads per instruction, stall samples."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    kernels = {}          # function -> {(file, line, src): [inst, thread_inst, samples]}
    cur_file, cur_fn, hdr = None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            cur_fn = r[1]
        elif r[0] == "Line No":
            hdr = r
            i_inst, i_thr, i_smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        elif hdr and r[0] not in ("", "-") and r[0].isdigit():
            def num(x):
                try:
                    return float(x)
                except ValueError:
                    return 0.0
            d = kernels.setdefault(cur_fn, {})
            key = (cur_file, int(r[0]), r[1].strip()[:90])
            e = d.setdefault(key, [0.0, 0.0, 0.0])
            e[0] += num(r[i_inst]); e[1] += num(r[i_thr]); e[2] += num(r[i_smp])
    for fn, d in kernels.items():
        if want not in fn:
            continue
        tot_i = sum(e[0] for e in d.values()) or 1.0
        tot_t = sum(e[1] for e in d.values())
        tot_s = sum(e[2] for e in d.values()) or 1.0
        print("==", fn[:110])
        print("   warp inst %.3e  thread inst %.3e  avg threads/inst %.2f  samples %d" % (tot_i, tot_t, tot_t / tot_i, tot_s))
        for (f, ln, src), e in sorted(d.items(), key=lambda kv: -kv[1][0])[:top]:
            print("   %5.1f%% inst %5.1f%% smp  thr %5.1f  %s:%d  %s" % (100 * e[0] / tot_i, 100 * e[2] / tot_s, e[1] / max(e[0], 1), f, ln, src))


if __name__ == "__main__":
    main()

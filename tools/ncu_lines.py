"""Per-source-line instruction / stall-sample summary from an ncu report (first launch).
usage: python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out, hdr, fname, seen_kernel = [], None, None, 0
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]
    if r and r[0] == "Function Name":
        pass
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        d = dict(zip(hdr, r))
        out.append((fname, int(r[0]), r[1], int(d["Instructions Executed"]), int(d["# Samples"]),
                    int(d["Thread Instructions Executed"])))
# only first launch: the listing repeats per launch; keep first occurrence of (file,line)
first = {}
for f, ln, src, ie, sm, te in out:
    first.setdefault((f, ln), (src, ie, sm, te))
tot_i = sum(v[1] for v in first.values()); tot_s = sum(v[2] for v in first.values())
print(f"total warp-instructions {tot_i}  samples {tot_s}")
for (f, ln), (src, ie, sm, te) in sorted(first.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{ln:4d} inst={ie:9d} ({100*ie/tot_i:4.1f}%) samp={sm:5d} ({100*sm/max(tot_s,1):4.1f}%) thr/inst={te/max(ie,1):4.1f} | {src[:90]}")

# per-function breakdown (line ranges of jss_device.cuh found by scanning the source)
import re
src = open("jssenv_b200/csrc/jss_device.cuh").read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:JSS_DEV|__global__)?\s*(?:\w[\w<> ,\*&:]*\s)?\s*(env_\w+|jss_\w+)\s*\(", l)
    if m and not l.startswith(" ") and not l.rstrip().endswith(";"):
        funcs.append((i, m.group(1)))
agg = {}
for (f, ln), (s_, ie, sm, te) in first.items():
    name = "other:" + f
    if f == "jss_device.cuh":
        name = "?"
        for start, fn in funcs:
            if start <= ln: name = fn
    a = agg.setdefault(name, [0, 0]); a[0] += ie; a[1] += sm
print()
for k, (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:32s} inst={ie:10d} ({100*ie/tot_i:4.1f}%)  per-env-step={ie/65536:7.1f}  samples={100*sm/max(tot_s,1):4.1f}%")

"""Key metrics + per-function instruction breakdown of the first launch in an ncu report.
usage: python tools/ncu_summary.py report.ncu-rep [n_envs]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'smsp__inst_executed.sum']
for k in keys:
    if k in hdr:
        i = hdr.index(k); print(f"{k:70s} {r[i]:>16s} {units[i]}")
stall = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") or h.startswith("smsp__average_warp_latency_issue_stalled")]
for v, h in sorted(stall, reverse=True)[:10]:
    print(f"  stall {h:75s} {v:8.3f}")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; fname = None; first = {}
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        d = dict(zip(hdr, r)); key = (fname, int(r[0]))
        if key not in first: first[key] = (r[1], int(d["Instructions Executed"]), int(d["# Samples"]), int(d["Thread Instructions Executed"]))
# function ranges from the source embedded in the report (--import-source on)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
funcs = []; cur = None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Name": cur = r[1].split("/")[-1]
    if cur == "jss_device.cuh" and len(r) == 2 and r[0].isdigit():
        m = re.match(r"^JSS_DEV\s+[\w<> ]+?[\s\*&]+(\w+)\s*\(", r[1]) or re.match(r"^(jss_env_kernel)\(", r[1])
        if m: funcs.append((int(r[0]), m.group(1)))
tot = sum(v[1] for v in first.values()); tots = sum(v[2] for v in first.values())
print(f"total warp-instructions {tot} = {tot / n_envs:.1f} per env-step; stall samples {tots}")
agg = {}
for (f, ln), (s_, ie, sm, te) in first.items():
    name = "other:" + f
    if f == "jss_device.cuh":
        name = "?"
        for st, fn in funcs:
            if st <= ln: name = fn
    a = agg.setdefault(name, [0, 0]); a[0] += ie; a[1] += sm
for k, (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if ie / n_envs >= 1: print(f"  {k:28s} inst/env-step={ie / n_envs:7.1f} ({100 * ie / tot:4.1f}%)  stall samples {100 * sm / max(tots, 1):4.1f}%")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
print("hot lines:")
for (f, ln), (s_, ie, sm, te) in sorted(first.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"  {f}:{ln:4d} inst={ie / n_envs:6.1f} samp={100 * sm / max(tots, 1):4.1f}% thr/inst={te / max(ie, 1):4.1f} | {s_[:100]}")

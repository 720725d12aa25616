"""Text summary of one `ncu --set full --import-source on` capture for profiles/:
    python tools/ncu_summary.py gpurun_out/<rep>.ncu-rep "<title>" [regions] > profiles/<name>.txt
Sections: launch + throughput figures, stall cycles per issued instruction, memory counters, hottest source lines,
optional source regions (file:lo-hi=name,...), and the instruction-fetch view (dynamic weight and fetch stalls per 1 KB of SASS)."""
import collections, csv, io, subprocess, sys

rep, title = sys.argv[1], sys.argv[2]
regions = sys.argv[3] if len(sys.argv) > 3 else ""
def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout
print(f"# {title}\n# source: {rep} (ncu --set full --clock-control none --import-source on)\n")
raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, vals = raw[0], raw[1], raw[2]
d = dict(zip(hdr, zip(units, vals)))
def show(keys):
    for k in keys:
        if k in d:
            print(f"{k:90s} {d[k][1]:>18s} {d[k][0]}")
print("## launch / throughput")
show(["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
      "launch__shared_mem_per_block_static", "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
      "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
      "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
      "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
      "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed",
      "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"])
print("\n## stall cycles per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active)")
st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v[1])) for k, v in d.items()
      if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for k, v in sorted(st, key=lambda kv: -kv[1]):
    if v >= 0.005:
        print(f"{k:28s} {v:8.3f}")

src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
cur_file, hd, lines, addr2src = None, None, [], {}
cur_line = None
for r in src:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No":
        hd = r; continue
    if hd is None or not r: continue
    def col(name):
        try: return float(r[hd.index(name)])
        except Exception: return 0.0
    if r[0] == "":
        if len(r) > 2 and r[2].startswith("0x"): addr2src[int(r[2], 16)] = (cur_file, cur_line)
        continue
    cur_line = int(r[0])
    lines.append((cur_file, cur_line, r[1].strip()[:86], col("Instructions Executed"), col("Thread Instructions Executed"), col("# Samples"),
                  col("stall_no_inst"), col("stall_wait"), col("stall_short_sb"), col("stall_barrier"), col("stall_long_sb")))
ti = sum(l[3] for l in lines) or 1.0; ts = sum(l[5] for l in lines) or 1.0; tn = sum(l[6] for l in lines) or 1.0
print(f"\n## hottest source lines (of {ti:.3e} warp-instructions, {ts:.0f} stall samples; share of instructions / lanes / share of samples / fetch / barrier)")
for l in sorted(lines, key=lambda l: -l[5])[:32]:
    print(f"{l[0]:22s}:{l[1]:4d} inst {100*l[3]/ti:5.2f}% lanes {l[4]/max(l[3],1):4.1f} samp {100*l[5]/ts:5.2f}% fetch {100*l[6]/tn:5.2f}% barrier {100*l[9]/ts:5.2f}% | {l[2]}")
if regions:
    regs = []
    for spec in regions.split(","):
        rng, name = spec.split("="); f, lh = rng.split(":"); lo, hi = lh.split("-"); regs.append((f, int(lo), int(hi), name))
    agg = collections.OrderedDict()
    for l in lines:
        name = "other: " + l[0]
        for f, lo, hi, nm in regs:
            if l[0] == f and lo <= l[1] <= hi: name = nm; break
        a = agg.setdefault(name, [0.0] * 7)
        for i in range(7): a[i] += l[3 + i] if i != 2 else l[5]
    print("\n## source regions: share of instructions, lanes, share of samples, share of fetch stalls, wait, short scoreboard, barrier (of samples)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print(f"{k:34s} inst {100*a[0]/ti:5.1f}% lanes {a[1]/max(a[0],1):4.1f} samp {100*a[2]/ts:5.1f}% fetch {100*a[3]/tn:5.1f}% wait {100*a[4]/ts:4.1f}% ssb {100*a[5]/ts:4.1f}% barrier {100*a[6]/ts:4.1f}%")

sass = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "sass"))))
hd, data = None, []
for r in sass:
    if r and r[0] == "Address": hd = r; continue
    if hd and len(r) == len(hd): data.append(r)
if data:
    ie, ni = hd.index("Instructions Executed"), hd.index("stall_no_inst")
    cnt = [float(r[ie] or 0) for r in data]; tot = sum(cnt) or 1.0
    ex = sum(1 for c in cnt if c > 0)
    print(f"\n## instruction-fetch view: {len(data)} SASS instructions = {len(data)*16/1024:.0f} KB, executed {ex} = {ex*16/1024:.0f} KB")
    order = sorted(cnt, reverse=True); run = 0.0; marks = [0.5, 0.8, 0.9, 0.95, 0.99]; mi = 0
    for i, c in enumerate(order):
        run += c
        while mi < len(marks) and run / tot >= marks[mi]:
            print(f"  {marks[mi]:.2f} of the dynamic instructions come from {i+1} static instructions = {(i+1)*16/1024:.1f} KB"); mi += 1
    tni = sum(float(r[ni] or 0) for r in data) or 1.0
    print("  1 KB blocks of SASS with >= 0.8 % of all instruction-fetch stalls: offset, executions per block instruction (relative to the hottest), fetch share, main source lines")
    mx = max(cnt) or 1.0
    for b in range(0, len(data), 64):
        seg = data[b:b + 64]; n = sum(float(r[ni] or 0) for r in seg)
        if n / tni < 0.008: continue
        c = sum(float(r[ie] or 0) for r in seg) / len(seg)
        srcs = collections.Counter(addr2src.get(int(r[0], 16)) for r in seg)
        top = ", ".join(f"{f}:{l}" for (f, l), k in [(k, v) for k, v in srcs.most_common(3) if k] )
        print(f"  +{b*16/1024:6.1f} KB  exec {c/mx:6.3f}  fetch {100*n/tni:5.2f}%  {top}")

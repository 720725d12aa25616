"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line and per function region.
usage: python tools/ncu_by_line.py export.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
lines = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or not r or r[0] == '': continue
    d = dict(zip(range(len(r)), r))
    def col(name):
        try: return float(r[hdr.index(name)])
        except Exception: return 0.0
    lines.append((cur_file, int(r[0]), r[1].strip()[:90], col('Instructions Executed'), col('Thread Instructions Executed'),
                  col('# Samples'), col('stall_no_inst'), col('stall_wait'), col('stall_short_sb'), col('stall_math'), col('L1 Wavefronts Shared Excessive')))
tot_i = sum(l[3] for l in lines); tot_s = sum(l[5] for l in lines); tot_ni = sum(l[6] for l in lines)
print(f"total inst {tot_i:.3e}  samples {tot_s:.0f}  no_inst {tot_ni:.0f}")
print("--- top lines by samples")
for l in sorted(lines, key=lambda l: -l[5])[:top]:
    print(f"{l[0]}:{l[1]:4d} inst {100*l[3]/tot_i:5.2f}% lanes {l[4]/max(l[3],1):4.1f} samp {100*l[5]/tot_s:5.2f}% noinst {100*l[6]/max(tot_ni,1):5.2f}% wait {l[7]:.0f} ssb {l[8]:.0f} math {l[9]:.0f} bankx {l[10]:.2e} | {l[2]}")

# ---- by region (line ranges of the kernel source at the profiled commit; pass as file:lo-hi=name,...)
if len(sys.argv) > 3:
    regs = []
    for spec in sys.argv[3].split(','):
        rng, name = spec.split('=')
        f, lh = rng.split(':')
        lo, hi = lh.split('-')
        regs.append((f, int(lo), int(hi), name))
    agg = collections.OrderedDict()
    for l in lines:
        name = 'other:' + l[0]
        for f, lo, hi, nm in regs:
            if l[0] == f and lo <= l[1] <= hi:
                name = nm; break
        a = agg.setdefault(name, [0.0] * 6)
        a[0] += l[3]; a[1] += l[4]; a[2] += l[5]; a[3] += l[6]; a[4] += l[7]; a[5] += l[8]
    print("--- regions: inst share, lanes, sample share, no_inst share, wait, short_sb")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print(f"{k:28s} inst {100*a[0]/tot_i:5.1f}% lanes {a[1]/max(a[0],1):4.1f} samp {100*a[2]/tot_s:5.1f}% noinst {100*a[3]/max(tot_ni,1):5.1f}% wait {100*a[4]/tot_s:4.1f}% ssb {100*a[5]/tot_s:4.1f}%")

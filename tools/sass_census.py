"""SASS census of the shipped library: per kernel and in total, how many bulk-copy / mbarrier / 256-bit-load / fp64 / shuffle /
barrier / local-memory instructions it holds (and that it holds no tensor-core or TMA-tensor instructions: the path has no
dense contraction).      python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "coregistrationgame_b200", "libficp_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
pat = {"UBLKCP (cp.async.bulk)": r"\bUBLKCP", "SYNCS (mbarrier)": r"\bSYNCS", "LDG.E.*.256 (256-bit global loads)": r"LDG\.E\S*\.256",
       "LDGSTS (cp.async)": r"\bLDGSTS", "DFMA/DADD/DMUL (fp64)": r"\b(DFMA|DADD|DMUL)\b", "DSETP": r"\bDSETP", "SHFL": r"\bSHFL",
       "BAR": r"\bBAR\.", "REDUX": r"\bREDUX", "ATOMS/ATOMG/RED": r"\b(ATOMS|ATOMG|RED)\b", "LDL/STL (local memory)": r"\b(LDL|STL)\b",
       "UTMALDG/UTMASTG (TMA tensor)": r"\bUTMA", "UTCMMA / tcgen05 / TMEM": r"UTC\w*MMA|LDTM|STTM", "HMMA/IMMA/DMMA": r"\b(HMMA|IMMA|DMMA)"}
rows, tot = [], collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    n = len(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", f, re.M))
    c = {k: len(re.findall(v, f)) for k, v in pat.items()}
    rows.append((name, n, c))
    tot.update(c)
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a; tools/sass_census.py)\n# {len(funcs)} kernels\n\n## totals")
for k in pat:
    print(f"{k:40s} {tot[k]}")
print("\n## per kernel: SASS instructions | UBLKCP | SYNCS | LDG.256 | fp64 arith | SHFL | BAR | local ld/st")
for (name, n, c), d in sorted(zip(rows, names), key=lambda r: -r[0][1]):
    d = re.sub(r"ficp::\(anonymous namespace\)::", "", d)[:110]
    print(f"{n:7d} | {c['UBLKCP (cp.async.bulk)']:3d} | {c['SYNCS (mbarrier)']:3d} | {c['LDG.E.*.256 (256-bit global loads)']:4d} | "
          f"{c['DFMA/DADD/DMUL (fp64)']:5d} | {c['SHFL']:4d} | {c['BAR']:3d} | {c['LDL/STL (local memory)']:3d} | {d}")

"""Aggregates ncu samples / executed warp instructions of lz77_v2.cu per phase (by 'PHASE_STAMP' / section markers).
usage: python tools/ncu_phases.py report.ncu-rep"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
src = open('compression_algorithms_b200/csrc/lz77_v2.cu').read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r'// -+ (P\d[^:]*):|// ---- (P4[ab])', l)
    if m: marks.append((i, (m.group(1) or m.group(2))))
def phase(ln):
    name = "pre"
    for i, nm in marks:
        if ln >= i: name = nm
    return name
import os
KFILTER = (["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else [])
out = subprocess.run(["ncu", "-i", rep] + KFILTER + ["--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if len(r) > 6 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    try: ln = int(r[0]); smp = int(r[hdr.index("# Samples")]); ins = int(r[hdr.index("Instructions Executed")])
    except ValueError: continue
    a = agg.setdefault(phase(ln), [0, 0]); a[0] += smp; a[1] += ins
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
for k, a in agg.items(): print("%-40s samples %5.1f%%  warp-instr %5.1f%% (%d)" % (k, 100.0*a[0]/ts, 100.0*a[1]/ti, a[1]))

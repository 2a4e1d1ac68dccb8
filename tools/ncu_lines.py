"""Aggregates ncu warp-stall samples per CUDA source line.
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
import os
KFILTER = (["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else [])
out = subprocess.run(["ncu", "-i", rep] + KFILTER + ["--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict(); stall = collections.defaultdict(lambda: collections.Counter())
fname = ""
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    try: ln = int(r[0]); smp = int(r[hdr.index("# Samples")]); ins = int(r[hdr.index("Instructions Executed")])
    except ValueError: continue
    key = (fname, ln)
    a = agg.setdefault(key, [0, 0, r[1]])
    a[0] += smp; a[1] += ins
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: stall[key][h] += int(r[i])
            except ValueError: pass
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print("total samples %d, warp instructions %d" % (tot, toti))
for key, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    s = ",".join("%s=%d" % (k[6:], v) for k, v in stall[key].most_common(3))
    print("%5.1f%% smp=%7d ins=%10d %s:%d  %s   [%s]" % (100.0 * a[0] / max(tot, 1), a[0], a[1], key[0], key[1], a[2].strip()[:90], s))

"""SASS evidence for every kernel of libb200comp.so: per-kernel instruction mix (the mnemonics that show
how a kernel uses the machine: shared-memory loads/stores/atomics, warp collectives, global accesses,
barriers) plus registers / shared memory from `cuobjdump -res-usage`, and the full SASS of the hot kernels.

    python tools/sass_summary.py            # writes profiles/r02_sass_summary.txt and profiles/sass/*.sass.gz
"""
import collections
import gzip
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "compression_algorithms_b200", "libb200comp.so")
OUT = os.path.join(ROOT, "profiles")
HOT = ("lz77_v2_kernel", "lz77_v4_kernel", "pdec_", "lz77_decode_kernel", "huff_encode_kernel", "huff_decode_kernel", "fse_encode_kernel",
       "fse_decode_kernel", "dfl_encode_kernel", "dfl_decode_kernel")
GROUPS = collections.OrderedDict([
    ("LDS", r"^LDS"), ("STS", r"^STS"), ("ATOMS", r"^ATOMS"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("ATOMG/RED", r"^(ATOMG|RED|ATOM)\b"),
    ("SHFL", r"^SHFL"), ("VOTE", r"^VOTE"), ("MATCH", r"^MATCH"), ("REDUX", r"^REDUX"), ("BAR", r"^BAR"), ("WARPSYNC", r"^WARPSYNC"),
    ("POPC/FLO/BREV", r"^(POPC|FLO|BREV)"), ("SHF/LOP3/IADD3/IMAD", r"^(SHF|LOP3|IADD3|IMAD|LEA)"), ("VIMNMX/VABSDIFF", r"^(VIMNMX|VABSDIFF|VIADD)"),
    ("BRA/BSSY/BSYNC", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET)"), ("TMA/UTMA", r"^(UTMA|UBLKCP|SYNCS)"), ("tcgen05/UTC", r"^(UTC|TCGEN)"),
])


def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip()
    except Exception:
        return n


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    kernels = collections.OrderedDict()
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name is not None:
            kernels[name].append(line)
    os.makedirs(os.path.join(OUT, "sass"), exist_ok=True)
    with open(os.path.join(OUT, "r02_sass_summary.txt"), "w") as f:
        f.write("# SASS summary of libb200comp.so (sm_100a, nvcc 12.9, -O3 -lineinfo); regenerate with tools/sass_summary.py\n")
        f.write("# counts are static instructions; none of these kernels uses TMA or tcgen05 (byte/bit work, no dense contraction)\n")
        for k, lines in kernels.items():
            ins = []
            for ln in lines:
                m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
                if m:
                    ins.append(m.group(1))
            d = demangle(k)
            d2 = d.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "").replace("(unsigned int)", "")
            short = re.sub(r"\(.*", "", d2)
            f.write("\n%s\n  %s\n  static instructions %d\n" % (short, usage.get(k, "(no res-usage line)"), len(ins)))
            mix = []
            for g, pat in GROUPS.items():
                c = sum(1 for i in ins if re.match(pat, i))
                if c:
                    mix.append("%s %d" % (g, c))
            f.write("  " + ", ".join(mix) + "\n")
            if any(h in d for h in HOT):
                fn = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_")[:80] + ".sass.gz"
                with gzip.open(os.path.join(OUT, "sass", fn), "wt") as g:
                    g.write("// %s\n" % d)
                    g.write("\n".join(lines))
    print("wrote", os.path.join(OUT, "r02_sass_summary.txt"), len(kernels), "kernels")


if __name__ == "__main__":
    main()

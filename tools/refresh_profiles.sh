#!/bin/bash
# Regenerates profiles/ from the files of the last evidence run under gpurun_out/ (suffix $1, e.g. r01d, and bench file $2).
set -e
cd "$(dirname "$0")/.."
S=$1; B=$2; R0=${3:-r02}   # R0: round prefix of the files written to profiles/
cp gpurun_out/$B profiles/bench_${R0}_final.json
[ -f gpurun_out/${S}_launches_deflate_100MB.csv ] && cp gpurun_out/${S}_launches_deflate_100MB.csv profiles/${R0}_launches_deflate_100MB.csv
[ -f gpurun_out/${S}_launches_all_codecs_100MB.csv ] && cp gpurun_out/${S}_launches_all_codecs_100MB.csv profiles/${R0}_launches_all_codecs_100MB.csv
R=gpurun_out/prof_lz77_${S}.ncu-rep
ncu -i $R -k regex:lz77_v2 --page details > profiles/${R0}_lz77_v2_ncu_details.txt 2>/dev/null
ncu -i $R -k regex:lz77_v2 --page raw --csv > profiles/${R0}_lz77_v2_ncu_raw.csv 2>/dev/null
ncu -i $R -k regex:lz77_decode_units --page details > profiles/${R0}_lz77_decode_units_ncu_details.txt 2>/dev/null
NCU_KERNEL=lz77_decode_units python tools/ncu_lines.py $R 25 > profiles/${R0}_lz77_decode_units_ncu_hot_lines.txt
NCU_KERNEL=lz77_v2 python tools/ncu_phases.py $R > profiles/${R0}_lz77_v2_ncu_phases.txt
NCU_KERNEL=lz77_v2 python tools/ncu_lines.py $R 40 > profiles/${R0}_lz77_v2_ncu_hot_lines.txt
(python tools/ncu_kernel_summary.py $R; echo; [ -f gpurun_out/prof_dfl_${S}.ncu-rep ] && { echo "# ---- deflate token entropy stage"; python tools/ncu_kernel_summary.py gpurun_out/prof_dfl_${S}.ncu-rep | tail -n +2; }; echo; echo "# ---- Huffman / FSE and the compaction kernels (earlier capture of this round; its lz77_v2 / lz77_decode / dfl rows are superseded by the ones above)"; python tools/ncu_kernel_summary.py gpurun_out/prof_all_codecs.ncu-rep | tail -n +2) > profiles/${R0}_all_kernels_ncu_summary.txt
[ -f gpurun_out/prof_lz77v4_${S}.ncu-rep ] && ncu -i gpurun_out/prof_lz77v4_${S}.ncu-rep --page raw --csv > profiles/${R0}_lz77_v4_ncu_raw.csv 2>/dev/null
R0=$R0 python - <<'PY'
import csv, json
import os
R0=os.environ.get('R0','r02')
rows=list(csv.reader(open('profiles/%s_lz77_v2_ncu_raw.csv' % R0)))
h,u,r=rows[0],rows[1],rows[2]
def g(n):
    i=h.index(n); return float(r[i].replace(',',''))*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u[i]]
rd,wr=g('dram__bytes_read.sum'),g('dram__bytes_write.sum')
out={"_source":"profiles/%s_lz77_v2_ncu_raw.csv (ncu --set full, one launch of lz77_v2_kernel<1> over a 100 000 000-byte enwik-shaped shard, 1 526 blocks of 64 KiB) and profiles/%s_lz77_v4_ncu_raw.csv (one launch of lz77_v4_kernel over 296 blocks of 64 KiB)" % (R0, R0),
 "lz77_v2_kernel<1>":{"input_bytes":100000000,"dram_bytes_read":int(rd),"dram_bytes_write":int(wr),"dram_bytes_per_launch":int(rd+wr),"dram_bytes_per_input_byte":round((rd+wr)/1e8,3)}}
try:
    rows=list(csv.reader(open('profiles/%s_lz77_v4_ncu_raw.csv' % R0)))
    h,u,r=rows[0],rows[1],rows[2]
    rd4,wr4=g('dram__bytes_read.sum'),g('dram__bytes_write.sum')
    nb4=296*65536
    out["lz77_v4_kernel"]={"input_bytes":nb4,"dram_bytes_read":int(rd4),"dram_bytes_write":int(wr4),"dram_bytes_per_launch":int(rd4+wr4),"dram_bytes_per_input_byte":round((rd4+wr4)/nb4,3)}
except Exception as e:
    print("no v4 raw page:", e)
json.dump(out,open('profiles/roofline_traffic.json','w'),indent=1)
PY
for f in v4_phase_cycles v4_vs_default decoder_sweep dropin_lz77 fse_segment_sweep phase_cycles v3_phase_cycles; do [ -f gpurun_out/${S}_$f.txt ] && cp gpurun_out/${S}_$f.txt profiles/${R0}_$f.txt; done
[ -f gpurun_out/${S}_reference_arm.json ] && cp gpurun_out/${S}_reference_arm.json profiles/${R0}_reference_arm.json
if [ -f gpurun_out/prof_lz77v4_${S}.ncu-rep ]; then
  ncu -i gpurun_out/prof_lz77v4_${S}.ncu-rep --page details > profiles/${R0}_lz77_v4_ncu_details.txt 2>/dev/null
  NCU_KERNEL=lz77_v4 python tools/ncu_lines.py gpurun_out/prof_lz77v4_${S}.ncu-rep 40 > profiles/${R0}_lz77_v4_ncu_hot_lines.txt
  (echo "# ---- lz77_v4_kernel (the default for text-like input), 296 blocks of 64 KiB"; python tools/ncu_kernel_summary.py gpurun_out/prof_lz77v4_${S}.ncu-rep | tail -n +2) >> profiles/${R0}_all_kernels_ncu_summary.txt
fi
if [ -f gpurun_out/prof_pdec_${S}.ncu-rep ]; then
  (echo; echo "# ---- token-parallel LZ77 decoder (lz77_pdec.cu), deflate variant, 256 MiB in 1 MiB blocks"; python tools/ncu_kernel_summary.py gpurun_out/prof_pdec_${S}.ncu-rep | tail -n +2) >> profiles/${R0}_all_kernels_ncu_summary.txt
fi
python tools/sass_summary.py > /dev/null

#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): the bench line, the reference arm, the ncu launch lists and the ncu --set full
# captures that tools/refresh_profiles.sh turns into profiles/. Usage: bash tools/evidence_run.sh <suffix>
S=${1:-r02x}
PART=${2:-all}     # a: bench + statistics + launch lists + the LZ77 captures; b: the entropy-coder captures (gpurun brings back at most 64 MiB per call)
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
if [ "$PART" != "b" ]; then
python bench.py > $O/${S}_bench.json 2> $O/${S}_bench.err || { echo "bench failed"; tail -5 $O/${S}_bench.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $O/${S}_reference_arm.json 2> $O/${S}_reference_arm.err
B200_LZ_V4=0 python tools/lz_stats.py 0 4 > $O/${S}_phase_cycles.txt 2>&1
B200_LZ_V3=1 python tools/lz_stats3.py 0 4 1 > $O/${S}_v3_phase_cycles.txt 2>&1
B200_LZ_V4=1 python tools/lz_stats4.py 0 4 > $O/${S}_v4_phase_cycles.txt 2>&1
python tools/v4_compare.py 200 > $O/${S}_v4_vs_default.txt 2>&1
python tools/pdec_sweep.py 256 > $O/${S}_decoder_sweep.txt 2>&1
python tools/dropin_time.py 32 > $O/${S}_dropin_lz77.txt 2>&1
python tools/fse_seg_sweep.py > $O/${S}_fse_segment_sweep.txt 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/${S}_launches_deflate_100MB.csv python bench.py --bytes 100000000 --no-cpu --no-detail --steps 2 --warmup 1 > $O/${S}_ncu1.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/${S}_launches_all_codecs_100MB.csv python tools/codec_run.py 100000000 1 > $O/${S}_ncu2.log 2>&1
B200_LZ_V4=0 $NCU --set full --import-source on -k regex:lz77 -c 8 -f -o $O/prof_lz77_${S} python bench.py --bytes 100000000 --no-cpu --no-detail --steps 1 --warmup 0 > $O/${S}_ncu3.log 2>&1
B200_LZ_V4=1 $NCU --set full --import-source on -k regex:lz77_v4 -c 1 -f -o $O/prof_lz77v4_${S} python tools/lz_run.py 296 > $O/${S}_ncu3b.log 2>&1
$NCU --set full --import-source on -k regex:pdec -c 12 -f -o $O/prof_pdec_${S} python tools/pdec_one.py 1 > $O/${S}_ncu3c.log 2>&1
fi
if [ "$PART" != "a" ]; then
$NCU --set full --import-source on -k regex:dfl -c 8 -f -o $O/prof_dfl_${S} python tools/codec_run.py 100000000 1 > $O/${S}_ncu4.log 2>&1
$NCU --set full -k regex:'huff|fse|hist|gather|offsets' -c 40 -f -o $O/prof_all_codecs python tools/codec_run.py 100000000 1 > $O/${S}_ncu5.log 2>&1
fi
ls -la $O | tail -20; du -sh $O

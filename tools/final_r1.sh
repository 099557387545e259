#!/bin/bash
# tools/final_r1.sh -- the round's measurement batch on one B200 (run through gpurun from the repo root); outputs in gpurun_out/
set -u
O=gpurun_out
python bench.py > $O/r1f_1gpu.json 2> $O/r1f_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 4 --warmup 1 > $O/r1f_ref.json 2> $O/r1f_ref.err; echo "ref rc=$?"
for s in 1 2 4; do ./tools/build/gather_probe 15 $s; done > $O/gather_probe_r1.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_list.json 2> $O/ncu_list.err; echo "ncu list rc=$?"
ncu --set full --clock-control none -k regex:'row_resident|row_group' --launch-skip 66 --launch-count 12 -o /tmp/prof_r1f_rows -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_rows.json 2> $O/ncu_rows.err; echo "ncu rows rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'heavy_accumulate' --launch-skip 110 --launch-count 3 -o $O/prof_r1f_heavy -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_heavy.json 2> $O/ncu_heavy.err; echo "ncu heavy rc=$?"
python tools/ncu_summary.py /tmp/prof_r1f_rows.ncu-rep > $O/ncu_r1f_rows.txt 2>&1
python tools/ncu_summary.py $O/prof_r1f_heavy.ncu-rep > $O/ncu_r1f_heavy.txt 2>&1
ls -la /tmp/*.ncu-rep $O/*.ncu-rep
# gpurun merges at most 64 MiB back: keep the big report only if it fits
sz=$(stat -c %s /tmp/prof_r1f_rows.ncu-rep); if [ "$sz" -lt 30000000 ]; then cp /tmp/prof_r1f_rows.ncu-rep $O/; fi
du -sh $O

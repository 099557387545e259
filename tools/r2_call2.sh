#!/bin/bash
# tools/r2_call2.sh -- 1-GPU call: GPU suite, probes, bench (new heavy pipeline: fused solve + 64-byte pair gathers, FFMA2 build),
# A/B against pair_gather=0, launch list + ncu --set full of the top kernel and the two biggest resident bins.
set -u
O=gpurun_out; mkdir -p $O
( time timeout 2400 python -m pytest tests -m gpu -q -x ) > $O/c2_pytest.log 2>&1; echo "pytest rc=$? $(grep -E 'passed|failed' $O/c2_pytest.log | tail -1)"
timeout 300 tools/bin/smem_gather_probe > $O/c2_smem_probe.txt 2>&1; timeout 300 tools/bin/smem_gather_probe 3072 1 >> $O/c2_smem_probe.txt 2>&1; echo "smem_probe rc=$?"; cat $O/c2_smem_probe.txt
timeout 900 python bench.py > $O/c2_bench.json 2> $O/c2_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --options pair_gather=0 --no-e2e --no-cpu-baseline --no-full-point --no-cli > $O/c2_bench_nopair.json 2> $O/c2_bench_nopair.err; echo "bench nopair rc=$?"
python - <<'E'
import json
for f in ("gpurun_out/c2_bench.json", "gpurun_out/c2_bench_nopair.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/sweep %.3f" % d["ms_per_step"], d["phases_ms"], "top us %.1f" % d["roofline"]["us_per_launch"], "frac", round(d["roofline"]["frac"], 3), d["roofline"]["bound"])
        if d.get("e2e"): print("  e2e %.1f G" % (d["e2e"]["value"] / 1e9), d["e2e"]["breakdown_rank0"])
    except Exception as e:
        print(f, "unreadable", e)
E
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-full-point --no-cli"
$B > $O/c2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r2.csv $B > $O/c2_ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'heavy_accumulate_kernel<2, 2, 2, 64, 0, 1>|row_resident_kernel<6, 1, 1>|row_group_kernel<6, 16, 1>|row_resident_kernel<3, 4, 1>' --launch-skip 45 --launch-count 15 -o $O/prof_r2_top -f $B > $O/c2_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O/*.ncu-rep; du -sh $O

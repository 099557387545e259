#!/bin/bash
# tools/r2_call1.sh -- first 1-GPU call of round 2: whole GPU suite, probes, FFMA2 A/B, default bench, launch list.
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/c1_smi.txt 2>&1
( time timeout 2400 python -m pytest tests -m gpu -q -x ) > $O/c1_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/c1_pytest.log)"
timeout 300 tools/bin/gather_probe > $O/c1_gather_probe.txt 2>&1; echo "gather_probe rc=$?"
timeout 300 tools/bin/smem_gather_probe > $O/c1_smem_probe.txt 2>&1; echo "smem_probe rc=$?"
timeout 300 tools/bin/smem_gather_probe 3072 1 >> $O/c1_smem_probe.txt 2>&1
timeout 900 bash tools/r2_ffma2.sh > $O/c1_ffma2.txt 2>&1; echo "ffma2 rc=$?"; tail -8 $O/c1_ffma2.txt
timeout 900 python bench.py > $O/c1_bench.json 2> $O/c1_bench.err; echo "bench rc=$?"
python - <<'E'
import json
try:
    d = json.loads(open("gpurun_out/c1_bench.json").read().strip().splitlines()[-1])
    print("ms/sweep %.3f" % d["ms_per_step"], d["phases_ms"]); print("roofline", json.dumps(d["roofline"])[:1500]); print("probes", d["probes"]); print("e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "fp", d["full_config_point"])
except Exception as e:
    print("bench unreadable", e)
E
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/c1_bench_ref.json 2> $O/c1_bench_ref.err; echo "bench ref rc=$?"; tail -c 1500 $O/c1_bench_ref.json

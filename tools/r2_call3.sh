#!/bin/bash
# tools/r2_call3.sh -- 2-GPU call: multi-GPU parity under pytest (all option sets), then the Netflix-shaped bench at N=2 with the
# set_train stage trace, default options and the round-1 path (host planner + cudaMalloc layout) for comparison.
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 1800 python -m pytest tests/test_zz_multi_gpu.py -q -x ) > $O/c3_pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$? $(grep -E 'passed|failed' $O/c3_pytest_mgpu.log | tail -1)"
timeout 900 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 --options trace=1 > $O/c3_bench2.json 2> $O/c3_bench2.err; echo "bench2 rc=$?"
timeout 900 $TR --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 --no-parity --options trace=1,device_plan=0,mgpu_pool=0 > $O/c3_bench2_r1path.json 2> $O/c3_bench2_r1path.err; echo "bench2 r1 path rc=$?"
grep -h "sbmf trace" $O/c3_bench2.err | tail -30
echo ---- r1 path
grep -h "sbmf trace" $O/c3_bench2_r1path.err | tail -30
python - <<'E'
import json
for f in ("gpurun_out/c3_bench2.json", "gpurun_out/c3_bench2_r1path.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f G/s" % (d["value"] / 1e9), "ms/sweep %.3f" % d["ms_per_step"], d["phases_ms"]); print("  e2e %.1f G/s" % (d["e2e"]["value"] / 1e9), d["e2e"]["breakdown_rank0"], d["e2e"].get("set_train_s_max_over_ranks"), "parity", d.get("parity"))
    except Exception as e:
        print(f, "unreadable:", e)
E

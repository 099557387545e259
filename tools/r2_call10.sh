#!/bin/bash
# tools/r2_call10.sh -- 2 GPUs: multi-GPU parity under pytest with the relabelled layout (chunked positions), then the bench at N=2
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( time timeout 1800 python -m pytest tests/test_zz_multi_gpu.py -q -x ) > $O/c10_pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$? $(grep -E 'passed|failed' $O/c10_pytest_mgpu.log | tail -1)"
for o in "" "relabel=0"; do
  timeout 900 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 --options "$o" > "$O/c10_bench2_$o.json" 2> "$O/c10_bench2_$o.err"; echo "bench2 '$o' rc=$?"
  python - "$O/c10_bench2_$o.json" <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f G/s" % (d["value"] / 1e9), "ms/sweep %.3f" % d["ms_per_step"], d["phases_ms"]); print("  e2e %.1f G/s" % (d["e2e"]["value"] / 1e9), d["e2e"]["breakdown_rank0"], d["e2e"].get("set_train_s_max_over_ranks"), "parity", d.get("parity"))
except Exception as e:
    print("unreadable:", e)
E
done

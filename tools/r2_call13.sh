#!/bin/bash
# tools/r2_call13.sh -- N GPUs (default 2): fused forward exchange (option fuse_exchange): multi-GPU parity under pytest, then A/B at N
set -u
O=gpurun_out; mkdir -p $O
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${SKIP_PYTEST:-0}" != 1 ]; then
( time timeout 1800 python -m pytest tests/test_zz_multi_gpu.py -q -x ) > $O/c13_pytest_mgpu.log 2>&1; echo "pytest mgpu rc=$? $(grep -E 'passed|failed' $O/c13_pytest_mgpu.log | tail -1)"
fi
p=29541
for o in "" "fuse_exchange=0"; do
  p=$((p+1))
  timeout 900 $TR --master-port $p bench.py --gpus $N --steps 20 --warmup 3 --options "$o" > "$O/c13_bench${N}_$o.json" 2> "$O/c13_bench${N}_$o.err"; echo "bench$N '$o' rc=$?"
  python - "$O/c13_bench${N}_$o.json" <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.1f G/s" % (d["value"] / 1e9), "ms/sweep %.3f" % d["ms_per_step"], d["phases_ms"]); print("  e2e %.1f G/s" % (d["e2e"]["value"] / 1e9), d["e2e"]["breakdown_rank0"], d["e2e"].get("set_train_s_max_over_ranks"), "parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_diff"))
except Exception as e:
    print("unreadable:", e)
E
done

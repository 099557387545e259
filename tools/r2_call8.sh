#!/bin/bash
# tools/r2_call8.sh -- 8-GPU call: Netflix-shaped K=100 (default and fused row updates), K=50, K=200, and BASELINE.json configs[4]
# (10M x 1M, 900M train ratings, K=200) sharded over 8 B200.  One bench.py line per run in gpurun_out/c8_*.json.
set -u
O=gpurun_out; mkdir -p $O
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # name port args...
  local name=$1 port=$2; shift 2
  timeout 900 $TR --master-port $port bench.py --gpus $N "$@" > $O/c${N}_$name.json 2> $O/c${N}_$name.err; echo "$name rc=$?"
  python - $O/c${N}_$name.json <<'E'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("   value %.1f G/s  %.3f ms/sweep  %s" % (d["value"] / 1e9, d["ms_per_step"], d["phases_ms"]))
    print("   e2e %s  set_train max %s  parity %s  top us %.1f" % (("%.1f G/s" % (e["value"] / 1e9)) if e else None, e.get("set_train_s_max_over_ranks"), (d.get("parity") or {}).get("ok"), d["roofline"]["us_per_launch"]))
except Exception as ex:
    print("   unreadable:", ex)
E
}
run netflix_k100 29531 --steps 20 --warmup 3
run netflix_k100_norelabel 29532 --steps 20 --warmup 3 --no-e2e --no-parity --options relabel=0
run netflix_k50 29533 --steps 20 --warmup 3 --workload netflix_k50 --no-parity
run netflix_k200 29534 --steps 20 --warmup 3 --workload netflix_k200 --no-parity
if [ "$N" = "8" ]; then
  run scaled_1b_k200 29535 --steps 5 --warmup 3 --workload scaled_1b_k200 --no-e2e --no-parity
  tail -3 $O/c8_scaled_1b_k200.err
fi
nvidia-smi --query-gpu=index,memory.used,clocks.sm --format=csv | head -9
free -g | head -2

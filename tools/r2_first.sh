#!/bin/bash
# tools/r2_first.sh -- first multi-GPU call of the next round (gpurun --gpus 2 -- 'bash tools/r2_first.sh'); outputs in gpurun_out/.
# Validates the two opt-in set_train paths written at the end of round 1 without a multi-GPU box:
#   SBMF_DEVICE_PLAN=1  exchange plan computed on the device instead of the host
#   SBMF_MGPU_POOL=1    rating-sized layout arrays of a multi-GPU handle from the stream-ordered pool
# 1) parity: G-GPU chain vs oracle and vs 1-GPU (tools/mgpu_check.py), default and with both flags
# 2) cost: SBMF_TRACE=1 stage times of set_train at Netflix size, default and with the flags, plus the e2e line
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 tools/mgpu_check.py > $O/r2_mgpu_default.log 2>&1; echo "mgpu_check default rc=$?"
SBMF_DEVICE_PLAN=1 $TR --master-port 29512 tools/mgpu_check.py > $O/r2_mgpu_devplan.log 2>&1; echo "mgpu_check device plan rc=$?"
SBMF_DEVICE_PLAN=1 SBMF_MGPU_POOL=1 $TR --master-port 29513 tools/mgpu_check.py > $O/r2_mgpu_devplan_pool.log 2>&1; echo "mgpu_check device plan + pool rc=$?"
SBMF_TRACE=1 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench2_default.json 2> $O/r2_bench2_default.err; echo "bench default rc=$?"
SBMF_TRACE=1 SBMF_DEVICE_PLAN=1 SBMF_MGPU_POOL=1 $TR --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench2_flags.json 2> $O/r2_bench2_flags.err; echo "bench flags rc=$?"
grep -h "sbmf trace" $O/r2_bench2_default.err | tail -24
echo ----
grep -h "sbmf trace" $O/r2_bench2_flags.err | tail -24
python - <<'E'
import json
for f in ("gpurun_out/r2_bench2_default.json", "gpurun_out/r2_bench2_flags.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.1f G/s" % (d["value"] / 1e9), "e2e %.1f G/s" % (d["e2e"]["value"] / 1e9), d["e2e"]["breakdown_rank0"])
    except Exception as e:
        print(f, "unreadable:", e)
E

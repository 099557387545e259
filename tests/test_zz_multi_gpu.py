"""Multi-GPU parity inside the pytest -m gpu suite (skipped on boxes with fewer than 2 GPUs): G-GPU zero-noise chains against the
oracle (<= 1e-4) and G-GPU live chains against the 1-GPU chain of the same seed, on ML-100K and on a matrix with streamed rows on
either side -- tools/mgpu_check.py under torchrun, for the default configuration and for every multi-GPU option
(host planner, cudaMalloc layout, NCCL exchanges instead of peer pushes)."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = len(glob.glob("/dev/nvidia[0-9]*"))

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(NGPU < 2, reason=f"needs >= 2 GPUs on the box (found {NGPU})")]

OPTION_SETS = ["", "device_plan=0,mgpu_pool=0", "peer=0", "device_plan=0,peer=0", "fuse_exchange=1", "relabel=0"]


@pytest.mark.parametrize("world", [2] + ([4] if NGPU >= 4 else []) + ([8] if NGPU >= 8 else []))
@pytest.mark.parametrize("opts", OPTION_SETS, ids=lambda o: o or "default")
def test_multi_gpu_chain_equals_single_gpu_and_oracle(world, opts):
    if world > 2 and opts not in ("", "peer=0"):
        pytest.skip("option matrix is covered at world 2")
    port = 29600 + 10 * world + OPTION_SETS.index(opts)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py"), opts]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "MGPU_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]

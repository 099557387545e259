#!/usr/bin/env python3
"""tools/mgpu_check.py -- multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
Every rank: G-GPU zero-noise sweeps vs the CPU oracle (<= 1e-4), G-GPU vs 1-GPU live chain on the same seed (same Philox
draws by construction), with rebuild_every 1 and 3 (the latter exercises the reverse CSC->CSR all-to-all).
Optional argument: "name=value,name=value" options for the multi-GPU handles (sbmf_cuda_set_option), e.g. "device_plan=0,peer=0".
tests/test_zz_multi_gpu.py runs it under pytest on boxes with >= 2 GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py as orc  # noqa: E402
import sbmf  # noqa: E402
from test_parity_gpu import skewed_case  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(np.max(np.abs(b)), 1e-30))


def main():
    opts = {}
    if len(sys.argv) > 1 and sys.argv[1]:
        opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[1].split(",")}
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(sbmf.nccl_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, 0)
    nid = bytes(idt.numpy().tobytes())

    ml = np.load(os.path.join(ROOT, "tests", "golden", "ml100k.npz"))
    cases = {"ml100k": {k: ml[k].astype(np.uint32 if "rating" not in k else np.float32) for k in ml.files}, "skewed": skewed_case()}
    cases["ml100k"].update(num_users=943, num_items=1682)
    sk = cases["skewed"]
    cases["skewed_T"] = dict(sk, train_user=sk["train_item"], train_item=sk["train_user"], test_user=sk["test_item"], test_item=sk["test_user"],
                             num_users=sk["num_items"], num_items=sk["num_users"])
    worst = 0.0
    first = True
    for name, d in cases.items():
        K = 20
        rs = np.random.RandomState(7)
        U0 = (0.1 * rs.standard_normal((d["num_users"], K))).astype(np.float32)
        V0 = (0.1 * rs.standard_normal((K, d["num_items"]))).astype(np.float32)
        for every in (1, 3):
            # a communicator per handle: a fresh id each time
            idt = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                idt = torch.frombuffer(bytearray(sbmf.nccl_unique_id()), dtype=torch.uint8).clone()
            dist.broadcast(idt, 0)
            m = sbmf.SbmfModel(K=K, sample_mode=2, device=local, rank=rank, world_size=world, nccl_id=idt.numpy().tobytes(), rebuild_every=every, options=opts)
            m.set_train(d["train_user"], d["train_item"], d["train_rating"], d["num_users"], d["num_items"])
            m.set_test(d["test_user"], d["test_item"], d["test_rating"])
            m.init_factors(U0, V0)
            m.sweep(10)
            gs = m.get_state(with_E=False)
            r_g = m.rmse_history(0, 10)[0]
            pred = m.get_pred()
            o = orc.Oracle(d["train_user"], d["train_item"], d["train_rating"], d["test_user"], d["test_item"], d["test_rating"],
                           d["num_users"], d["num_items"], K, noise=orc.NOISE_ZERO)
            o.init_factors(U0.astype(np.float64), V0.astype(np.float64))
            r_o, _ = o.sweep(10)
            os_ = o.state()
            errs = {k: rel(gs[k], os_[k]) for k in ("U", "V", "b_i", "b_j", "sigma_u", "mu_u", "sigma_v", "mu_v", "mu_b_i", "sigma_b_i")}
            errs["rmse"] = float(np.max(np.abs(r_g - r_o)))
            errs["pred"] = float(np.max(np.abs(pred - o.pred_mean())))
            errs["alpha"] = abs(gs["alpha"] - os_["alpha"]) / os_["alpha"]
            w = max(errs.values())
            worst = max(worst, w)
            if rank == 0:
                print(f"[{name} rebuild_every={every}] world={world} zero-noise vs oracle: max err {w:.2e}", flush=True)
            assert w <= 1e-4, errs
            m.close()
        # live chain: G GPUs vs 1 GPU, same seed
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(sbmf.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, 0)
        mg = sbmf.SbmfModel(K=K, sample_mode=0, seed=11, device=local, rank=rank, world_size=world, nccl_id=idt.numpy().tobytes(), options=opts)
        m1 = sbmf.SbmfModel(K=K, sample_mode=0, seed=11, device=local)
        mg.set_timing_enabled(0)     # multi-GPU chain replays the captured CUDA graph (with the NCCL calls inside) from sweep 2 on
        for m in (mg, m1):
            m.set_train(d["train_user"], d["train_item"], d["train_rating"], d["num_users"], d["num_items"])
            m.set_test(d["test_user"], d["test_item"], d["test_rating"])
            m.init_factors()
            m.sweep(5)
        a, b = mg.get_state(with_E=False), m1.get_state(with_E=False)
        e = max(rel(a[k], b[k].astype(np.float64)) for k in ("U", "V", "b_i", "b_j"))
        e = max(e, float(np.max(np.abs(mg.rmse_history(0, 5)[0] - m1.rmse_history(0, 5)[0]))))
        if rank == 0:
            print(f"[{name}] live chain {world} GPUs vs 1 GPU: max rel diff {e:.2e}", flush=True)
        assert e <= 1e-4, e
        mg.close(); m1.close()
    flag = torch.tensor([1])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    if rank == 0:
        print("MGPU_CHECK_OK worst", worst, "options", opts, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU tests of the multi-GPU host logic (csrc/plan.cpp through the C ABI; no GPU): nnz-balanced shard bounds and the residual
all-to-all plan, including a real 2-process exchange over torch.distributed/gloo that must reproduce the single-GPU permute."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_py as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200")


@pytest.fixture(scope="module")
def sbmf_mod():
    subprocess.run(["make", "-C", PKG], check=True, capture_output=True)
    import sbmf
    return sbmf


def layout_of(d):
    o = orc.Oracle(d["train_user"], d["train_item"], d["train_rating"], d["test_user"], d["test_item"], d["test_rating"],
                   d["num_users"], d["num_items"], 4)
    return o.layout()


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_plan_shards_balanced_and_complete(world, ml100k, sbmf_mod):
    L = layout_of(ml100k)
    for ptr in (L["row_ptr"], L["col_ptr"]):
        b = sbmf_mod.plan_shards(ptr, world)
        assert b[0] == 0 and b[-1] == ptr.size - 1 and np.all(np.diff(b.astype(np.int64)) >= 0)
        nnz = np.diff(ptr[b.astype(np.int64)])
        assert nnz.sum() == ptr[-1]
        maxrow = np.diff(ptr).max()
        assert nnz.max() <= ptr[-1] / world + maxrow     # no shard exceeds the ideal by more than one row


def test_plan_shards_degenerate(sbmf_mod):
    ptr = np.array([0, 0, 0, 10, 10], np.int64)          # one non-empty row, 4 ranks
    b = sbmf_mod.plan_shards(ptr, 4)
    assert b[0] == 0 and b[-1] == 4 and np.diff(ptr[b.astype(np.int64)]).sum() == 10
    b = sbmf_mod.plan_shards(np.zeros(6, np.int64), 3)   # empty matrix
    assert b[0] == 0 and b[-1] == 5


@pytest.mark.parametrize("case", ["ml100k", "tiny"])
@pytest.mark.parametrize("world", [2, 3])
def test_plan_exchange_reproduces_permute(case, world, ml100k, tiny, sbmf_mod):
    """Simulate all ranks in one process: pack / 'send' / unpack must equal e_csc = e_csr[perm]."""
    d = ml100k if case == "ml100k" else tiny
    L = layout_of(d)
    N = L["perm"].size
    perm = L["perm"].astype(np.uint32)
    ub, ib = sbmf_mod.plan_shards(L["row_ptr"], world), sbmf_mod.plan_shards(L["col_ptr"], world)
    cb, tb = L["row_ptr"][ub.astype(np.int64)], L["col_ptr"][ib.astype(np.int64)]
    e_csr = np.random.RandomState(0).standard_normal(N).astype(np.float32)
    plans = [sbmf_mod.plan_exchange(perm, world, r, cb, tb) for r in range(world)]
    for r in range(world):
        assert plans[r][1].sum() == cb[r + 1] - cb[r] and plans[r][3].sum() == tb[r + 1] - tb[r]
        assert sorted(plans[r][0].tolist()) == list(range(int(cb[r + 1] - cb[r])))       # every local slot sent exactly once
    for r in range(world):           # receiver r
        recv = []
        for q in range(world):       # from sender q: q's send buffer segment for destination r
            send_idx, sc = plans[q][0], plans[q][1]
            off = int(sc[:r].sum())
            seg = e_csr[cb[q]:cb[q + 1]][send_idx[off:off + int(sc[r])]]
            assert seg.size == plans[r][3][q]
            recv.append(seg)
        recvbuf = np.concatenate(recv) if recv else np.empty(0, np.float32)
        e_csc_local = recvbuf[plans[r][2]]
        assert np.array_equal(e_csc_local, e_csr[perm][tb[r]:tb[r + 1]])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ml100k", "tiny"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_device_planner_equals_host_planner(case, world, ml100k, tiny, sbmf_mod):
    """The exchange plan computed on the device (SBMF_DEVICE_PLAN=1 path of set_train) is bit-identical to plan.cpp's, for every
    rank: send order, receive positions and the traffic matrix.  Runs on ONE GPU (the planner takes the bounds as arguments)."""
    d = ml100k if case == "ml100k" else tiny
    L = layout_of(d)
    perm = L["perm"].astype(np.uint32)
    ub, ib = sbmf_mod.plan_shards(L["row_ptr"], world), sbmf_mod.plan_shards(L["col_ptr"], world)
    cb, tb = L["row_ptr"][ub.astype(np.int64)], L["col_ptr"][ib.astype(np.int64)]
    for r in range(world):
        send_idx, sc, recv_pos, rc = sbmf_mod.plan_exchange(perm, world, r, cb, tb)
        d_send, d_recv, pc = sbmf_mod.plan_exchange_device(perm, world, r, cb, tb)
        assert np.array_equal(d_send, send_idx), (world, r)
        assert np.array_equal(d_recv, recv_pos), (world, r)
        assert np.array_equal(pc[r, :], sc) and np.array_equal(pc[:, r], rc), (world, r)
        assert pc.sum() == perm.size


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["SBMF_PKG"]); sys.path.insert(0, os.environ["SBMF_TESTS"])
import sbmf, oracle_py as orc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
d = np.load(os.path.join(os.environ["SBMF_TESTS"], "golden", "ml100k.npz"))
tu, ti, tr = d["train_user"].astype(np.uint32), d["train_item"].astype(np.uint32), d["train_rating"].astype(np.float64)
o = orc.Oracle(tu, ti, tr, tu[:1], ti[:1], tr[:1], 943, 1682, 4)
L = o.layout()
perm = L["perm"].astype(np.uint32)
ub, ib = sbmf.plan_shards(L["row_ptr"], world), sbmf.plan_shards(L["col_ptr"], world)
cb, tb = L["row_ptr"][ub.astype(np.int64)], L["col_ptr"][ib.astype(np.int64)]
send_idx, sc, recv_pos, rc = sbmf.plan_exchange(perm, world, rank, cb, tb)
e_csr = np.random.RandomState(0).standard_normal(perm.size).astype(np.float32)     # same on every rank
mine = e_csr[cb[rank]:cb[rank + 1]]
sendbuf = torch.from_numpy(mine[send_idx].copy())
recvbuf = torch.empty(int(rc.sum()), dtype=torch.float32)
ins = list(torch.split(sendbuf, [int(x) for x in sc]))
outs = list(torch.split(recvbuf, [int(x) for x in rc]))
reqs = []
for q in range(world):           # all-to-all-v from send/recv pairs (gloo has no all_to_all)
    if q == rank:
        outs[q].copy_(ins[q])
        continue
    reqs.append(dist.isend(ins[q].contiguous(), q))
    reqs.append(dist.irecv(outs[q], q))
for rq in reqs:
    rq.wait()
got = torch.cat(outs).numpy()[recv_pos]
want = e_csr[perm][tb[rank]:tb[rank + 1]]
ok = np.array_equal(got, want)
flag = torch.tensor([1 if ok else 0]); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 3)
'''


def test_two_process_gloo_exchange(tmp_path, sbmf_mod):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SBMF_PKG=PKG, SBMF_TESTS=os.path.join(ROOT, "tests"), MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    rcs = [p.wait(timeout=240) for p in procs]
    assert rcs == [0, 0], rcs


# ------------------------------------------------------------------------------------------ sparse host workload generator
def test_synth_host_generator_properties(sbmf_mod):
    """csrc/synth_host.cpp (bench input for the 10M x 1M / 1B configuration): distinct pairs sorted by (user, item), the requested
    number of ratings in expectation, Zipf-shaped marginals, test split ~ test_frac, half-star ratings in [0.5, 5], and the same
    matrix whatever the number of host threads."""
    sbmf = sbmf_mod
    I, J, N = 20000, 3000, 600000
    a = sbmf.synth_generate_host(I, J, N, seed=7, threads=1)
    b = sbmf.synth_generate_host(I, J, N, seed=7, threads=4)
    for k in ("train_user", "train_item", "train_rating", "test_user", "test_item", "test_rating"):
        assert np.array_equal(a[k], b[k]), k
    ntr, nte = a["train_user"].size, a["test_user"].size
    assert abs(ntr + nte - N) < 5 * np.sqrt(N)
    assert abs(nte / (ntr + nte) - 0.1) < 0.005
    for p in ("train", "test"):
        key = a[p + "_user"].astype(np.int64) * J + a[p + "_item"]
        assert np.all(key[1:] > key[:-1])                      # sorted by (user, item), no duplicate pair
        assert a[p + "_user"].max() < I and a[p + "_item"].max() < J
        r = a[p + "_rating"]
        assert r.min() >= 0.5 and r.max() <= 5.0 and np.all(r * 2 == np.round(r * 2))
    assert np.intersect1d(a["train_user"].astype(np.int64) * J + a["train_item"], a["test_user"].astype(np.int64) * J + a["test_item"]).size == 0
    # marginals: degree of the rank-r item ~ r^-1, of the rank-r user ~ r^-0.8 (both flattened at the top by p = min(1, .) and by
    # sorting noisy degrees, hence the one-sided tolerances)
    dj = np.sort(np.bincount(np.concatenate([a["train_item"], a["test_item"]]), minlength=J))[::-1].astype(np.float64)
    du = np.sort(np.bincount(np.concatenate([a["train_user"], a["test_user"]]), minlength=I))[::-1].astype(np.float64)
    sj = np.polyfit(np.log(np.arange(30, 1000) + 1.0), np.log(dj[30:1000]), 1)[0]
    su = np.polyfit(np.log(np.arange(30, 5000) + 1.0), np.log(du[30:5000]), 1)[0]
    assert -1.05 < sj < -0.85 and -0.85 < su < -0.6, (sj, su)
    c = sbmf.synth_generate_host(I, J, N, seed=8, threads=4)
    assert not np.array_equal(a["train_item"][:1000], c["train_item"][:1000])
    with pytest.raises(sbmf.SbmfError):
        sbmf.synth_generate_host(10, 10, 90)                    # denser than half of the pair grid

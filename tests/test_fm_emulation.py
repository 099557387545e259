"""CPU: the general FM Gibbs path of csrc/fm.cu without a GPU.

* tools/fm_emulate.cu executes the kernels' formulas (csrc/fm_math.cuh, shared with the kernels), precisions (fp32 caches and
  parameters), run partition and two-phase run schedule sequentially on the host; after 10 zero-noise iterations it must agree
  with the fp64 restatement of libFM (which draws attribute after attribute) to 1e-4 relative -- the parity bar of the GPU tests.
* sbmf_fm_plan_runs (host planner inside libsbmf_cuda.so): runs are contiguous, conflict-free and maximal on random matrices."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import fm_oracle_py as fmo
from test_fm_oracle import load_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@pytest.fixture(scope="module")
def emu():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = os.path.join(ROOT, "tools", "build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libfm_emulate.so")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(PKG, "csrc"),
                    "-o", so, os.path.join(ROOT, "tools", "fm_emulate.cu")], check=True, capture_output=True)
    L = C.CDLL(so)
    L.fm_emulate.restype = C.c_int
    L.fm_emulate.argtypes = ([C.c_uint32] + [C.c_void_p] * 4 + [C.c_uint32] + [C.c_void_p] * 4 + [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                             C.c_int, C.c_double, C.c_double, C.c_double, C.c_uint32] + [C.c_void_p] * 7)
    return L


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def emulate(L, tr, te, p, G, group, K, iters, w, v, k0=1, k1=1, multilevel=1, reg=(0.0, 0.0, 0.0)):
    n, nt = tr["y"].size, te["y"].size
    w, v = w.astype(np.float32).copy(), np.ascontiguousarray(v.astype(np.float32))
    hyp = np.zeros(2 * G + 2 * G * K + 2)
    e, rtr, rte, runs = np.zeros(n, dtype=np.float32), np.zeros(iters), np.zeros(iters), np.zeros(p + 1, dtype=np.uint32)
    nr = L.fm_emulate(n, _p(tr["row_ptr"]), _p(tr["attr"]), _p(tr["x"]), _p(tr["y"]), nt, _p(te["row_ptr"]), _p(te["attr"]), _p(te["x"]), _p(te["y"]),
                      p, G, _p(group), K, k0, k1, multilevel, reg[0], reg[1], reg[2], iters, _p(w), _p(v), _p(hyp), _p(e), _p(rtr), _p(rte), _p(runs))
    return {"w": w, "v": v, "w_mu": hyp[:G], "w_lambda": hyp[G:2 * G], "v_mu": hyp[2 * G:2 * G + G * K].reshape(G, K),
            "v_lambda": hyp[2 * G + G * K:2 * G + 2 * G * K].reshape(G, K), "w0": hyp[-2], "alpha": hyp[-1], "e": e, "rmse_train": rtr, "rmse_test": rte,
            "run_begin": runs[:nr + 1]}


@pytest.mark.parametrize("name,K,mode", [("tiny_libfm", 4, "mcmc"), ("fm_general", 3, "mcmc"), ("fm_general", 5, "als"), ("fm_general", 0, "mcmc")])
def test_emulated_schedule_matches_the_sequential_checker(emu, name, K, mode):
    tr, te, group = load_fixture(name)
    p = fmo.num_attributes(tr, te)
    G = 1 if group is None else int(group.max()) + 1
    kw = dict(do_sample=0, do_multilevel=0, reg=(0.25, 1.0, 4.0)) if mode == "als" else {}
    rs = np.random.RandomState(5)
    w0 = (0.1 * rs.standard_normal(p)).astype(np.float32)
    v0 = (0.1 * rs.standard_normal((K, p))).astype(np.float32)
    o = fmo.FmOracle(tr, te, K, num_attr=p, attr_group=group, noise=fmo.NOISE_ZERO, **kw)
    o.init(w0.astype(np.float64), v0.astype(np.float64))
    rtr, rte = o.learn(10)
    so = o.state()
    got = emulate(emu, tr, te, p, G, group, K, 10, w0, v0, multilevel=0 if mode == "als" else 1, reg=kw.get("reg", (0.0, 0.0, 0.0)))
    for k in ("w", "w_mu", "w_lambda", "e") + (("v", "v_mu", "v_lambda") if K else ()):
        assert rel(got[k], so[k]) <= 1e-4, (k, rel(got[k], so[k]))
    assert abs(got["w0"] - so["w0"]) <= 1e-4 * max(abs(so["w0"]), 1.0) and abs(got["alpha"] - so["alpha"]) <= 1e-4 * so["alpha"]
    assert np.max(np.abs(got["rmse_train"] - rtr)) <= 1e-5 and np.max(np.abs(got["rmse_test"] - rte)) <= 1e-5
    # the runs: users | items (+ the data-free attributes behind them) for the MF fixture; co-occurring genres and the two dense
    # attributes are runs of their own in the general one
    rb = got["run_begin"].tolist()
    if name == "tiny_libfm":
        assert rb == [0, 50, p]
    else:
        assert rb[:3] == [0, 30, 70] and rb[-1] == p and len(rb) > 4
    o.close()


def test_emulated_schedule_on_long_columns(emu):
    """columns of 40,000 / 54,000 cases: the column sums must be fp64 (fm_math.cuh ColSums) -- an fp32 running sum is off by 1e-2"""
    from fm_gpu_cases import long_column_matrix
    tr, te = long_column_matrix()
    p, K = 2004, 3
    group = np.concatenate([np.zeros(2), np.ones(2000), np.full(2, 2)]).astype(np.uint32)
    rs = np.random.RandomState(5)
    w0 = (0.1 * rs.standard_normal(p)).astype(np.float32)
    v0 = (0.1 * rs.standard_normal((K, p))).astype(np.float32)
    o = fmo.FmOracle(tr, te, K, num_attr=p, attr_group=group, noise=fmo.NOISE_ZERO)
    o.init(w0.astype(np.float64), v0.astype(np.float64))
    o.learn(5)
    so = o.state()
    got = emulate(emu, tr, te, p, 3, group, K, 5, w0, v0)
    for k in ("w", "v", "w_mu", "w_lambda", "v_mu", "v_lambda", "e"):
        assert rel(got[k], so[k]) <= 1e-4, (k, rel(got[k], so[k]))
    assert got["run_begin"].tolist() == [0, 2, 2002, 2004]
    o.close()


def brute_force_runs(rows, p):
    """greedy scan with an explicit per-case marker (what libFM's order allows): a new run starts at the first attribute that
    shares a case with an attribute already in the run"""
    cols = [[] for _ in range(p)]
    for c, r in enumerate(rows):
        for a in r:
            cols[a].append(c)
    run_of_case, begins, cur = {}, [0], 0
    for j in range(p):
        if any(run_of_case.get(c) == cur for c in cols[j]):
            begins.append(j)
            cur += 1
        for c in cols[j]:
            run_of_case[c] = cur
    return begins + [p]


def test_planner_runs_are_contiguous_conflict_free_and_maximal():
    subprocess.run(["make", "-C", PKG], check=True, capture_output=True)
    lib = C.CDLL(os.path.join(PKG, "lib", "libsbmf_cuda.so"))
    lib.sbmf_fm_plan_runs.restype = C.c_uint32
    lib.sbmf_fm_plan_runs.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    rs = np.random.RandomState(11)
    for trial in range(60):
        p = int(rs.randint(1, 40))
        n = int(rs.randint(1, 60))
        rows = [sorted(rs.choice(p, size=rs.randint(0, min(p, 5) + 1), replace=False).tolist()) for _ in range(n)]
        nxt = np.full(p, 0xFFFFFFFF, dtype=np.uint32)
        for r in rows:
            for a, b in zip(r[:-1], r[1:]):
                nxt[a] = min(nxt[a], b)
        out = np.zeros(p + 1, dtype=np.uint32)
        nr = lib.sbmf_fm_plan_runs(p, _p(nxt), _p(out))
        assert out[:nr + 1].tolist() == brute_force_runs(rows, p), (trial, rows)

"""tests/fm_gpu_cases.py -- the GPU parity cases of the general FM Gibbs path (csrc/fm.cu, include/sbmf_fm_cuda.h), each runnable on
its own: `python tests/fm_gpu_cases.py <case>` exits 0 on success.  tests/test_z_fm_parity_gpu.py runs every case in a process of
its own, so that a fault in this (newest) path cannot take the rest of the GPU suite with it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200"), os.path.join(ROOT, "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

import fm_oracle_py as fmo   # noqa: E402
import sbmf                  # noqa: E402
from test_fm_oracle import load_fixture   # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)) if b.size else 0.0


def pair(tr, te, group, K, p=None, **kw):
    """device model + fp64 checker on the same data; kw: k0, k1, do_sample, do_multilevel, reg"""
    p = p or fmo.num_attributes(tr, te)
    reg = kw.pop("reg", (0.0, 0.0, 0.0))
    zero = kw.pop("zero", True)
    m = sbmf.FmModel(p, K, attr_group=group, sample_mode=sbmf.FM_SAMPLE_ZERO if zero else sbmf.FM_SAMPLE_LIVE, reg0=reg[0], regw=reg[1], regv=reg[2],
                     **kw)
    m.set_train(tr)
    m.set_test(te)
    o = fmo.FmOracle(tr, te, K, num_attr=p, attr_group=group, noise=fmo.NOISE_ZERO if zero else fmo.NOISE_RAND, reg=reg,
                     **{k: v for k, v in kw.items() if k in ("k0", "k1", "do_sample", "do_multilevel")})
    return m, o, p


def check_zero_noise(tr, te, group, K, iters=10, tol=1e-4, **kw):
    m, o, p = pair(tr, te, group, K, **kw)
    rs = np.random.RandomState(5)
    w0 = (0.1 * rs.standard_normal(p)).astype(np.float32)
    v0 = (0.1 * rs.standard_normal((K, p))).astype(np.float32)
    m.init(w0, v0)
    o.init(w0.astype(np.float64), v0.astype(np.float64))
    e0 = rel(m.get_state()["e"], o.state()["e"])
    assert e0 <= 1e-5, ("initial prediction pass", e0)
    m.learn(iters)
    rtr, rte = o.learn(iters)
    g, s = m.get_state(), o.state()
    assert g["iterations"] == iters
    for k in ("w", "w_mu", "w_lambda", "e", "pred_sum") + (("v", "v_mu", "v_lambda") if K else ()):
        assert rel(g[k], s[k]) <= tol, (k, rel(g[k], s[k]))
    assert abs(g["w0"] - s["w0"]) <= tol * max(abs(s["w0"]), 1.0) and abs(g["alpha"] - s["alpha"]) <= tol * s["alpha"], (g["w0"], s["w0"], g["alpha"], s["alpha"])
    gtr, gte = m.rmse_history(0, iters)
    assert np.max(np.abs(gtr - rtr)) <= 1e-5 and np.max(np.abs(gte - rte)) <= 1e-5, (gtr, rtr, gte, rte)
    # fm_learn_mcmc::predict, [G]:355-379
    lo, hi = float(tr["y"].min()), float(tr["y"].max())
    want = np.clip(s["pred_sum"] / iters, lo, hi) if kw.get("do_sample", 1) else None
    if want is not None:
        assert np.max(np.abs(m.predict() - want)) <= 1e-4
    m.close()
    o.close()


def case_columns():
    """the transposed design matrix built on the device == libFM's create_data_t (Data.h:472-528), bit for bit; the runs"""
    for name in ("tiny_libfm", "fm_general"):
        tr, te, group = load_fixture(name)
        m, o, p = pair(tr, te, group, 2)
        got, want = m.get_columns(), o.columns()
        for k in ("col_ptr", "case", "x"):
            assert np.array_equal(got[k], want[k]), (name, k)
        rb = m.get_runs().tolist()
        assert rb == ([0, 50, p] if name == "tiny_libfm" else [0, 30, 70, 71, 72, 73, 74, 75, 76, 77, p]), rb
        m.close()
        o.close()


def case_zero_mf():
    tr, te, group = load_fixture("tiny_libfm")
    check_zero_noise(tr, te, group, 4)


def case_zero_general():
    tr, te, group = load_fixture("fm_general")
    check_zero_noise(tr, te, group, 3)


def case_zero_general_k20():
    tr, te, group = load_fixture("fm_general")
    check_zero_noise(tr, te, group, 20)          # 32 lanes per case in the prediction pass


def case_zero_als():
    tr, te, group = load_fixture("fm_general")
    check_zero_noise(tr, te, group, 5, do_sample=0, do_multilevel=0, reg=(0.25, 1.0, 4.0))


def case_zero_variants():
    tr, te, group = load_fixture("fm_general")
    check_zero_noise(tr, te, group, 0)           # linear model
    check_zero_noise(tr, te, group, 3, k0=0)
    check_zero_noise(tr, te, group, 3, k1=0)
    check_zero_noise(tr, te, None, 3)            # one group


def long_column_matrix(n=60000, seed=3):
    """columns of every tier: attribute 0 in 3 of 4 cases (sliced, > 16384 entries), attribute 1 in the rest (CTA tier or sliced),
    a one-hot block of 2000 short columns, one dense real-valued attribute (a run of its own, sliced)"""
    rs = np.random.RandomState(seed)
    a0 = np.where(rs.rand(n) < 0.75, 0, 1).astype(np.uint32)
    a1 = (2 + rs.randint(0, 2000, size=n)).astype(np.uint32)
    xd = (np.round(rs.standard_normal(n) * 8) / 8).astype(np.float32)
    y = np.clip(np.round((3.0 + 0.5 * (a0 == 0) + 0.3 * np.sin(a1) + 0.4 * xd + 0.3 * rs.standard_normal(n)) * 2) / 2, 0.5, 5.0).astype(np.float32)
    attr = np.stack([a0, a1, np.full(n, 2002, dtype=np.uint32)], axis=1).reshape(-1)
    x = np.stack([np.ones(n, np.float32), np.ones(n, np.float32), xd], axis=1).reshape(-1)
    row_ptr = (3 * np.arange(n + 1)).astype(np.int64)
    cut = int(n * 0.9)
    tr = {"row_ptr": row_ptr[:cut + 1], "attr": attr[:3 * cut], "x": x[:3 * cut], "y": y[:cut]}
    te = {"row_ptr": row_ptr[cut:] - row_ptr[cut], "attr": attr[3 * cut:], "x": x[3 * cut:], "y": y[cut:]}
    return tr, te


def case_long_columns():
    tr, te = long_column_matrix()
    group = np.concatenate([np.zeros(2), np.ones(2000), np.full(2, 2)]).astype(np.uint32)
    check_zero_noise(tr, te, group, 3, iters=5, p=2004)


def case_live():
    """live sampling: draws are Philox on the device and glibc rand() in libFM, so chains are compared as distributions on
    ML-100K cast as a factorization machine (one-hot user + one-hot item): mean test-RMSE trajectory over seeds within 0.003
    from iteration 2 on (and within 4 standard errors before); bit-identical repeats for one seed"""
    d = np.load(os.path.join(GOLDEN, "ml100k.npz"))
    I = int(max(d["train_user"].max(), d["test_user"].max())) + 1

    def fm(u, i, r):
        n = u.size
        return {"row_ptr": (2 * np.arange(n + 1)).astype(np.int64), "attr": np.stack([u, I + i], axis=1).reshape(-1).astype(np.uint32),
                "x": np.ones(2 * n, dtype=np.float32), "y": r.astype(np.float32)}
    tr, te = fm(d["train_user"], d["train_item"], d["train_rating"]), fm(d["test_user"], d["test_item"], d["test_rating"])
    p = fmo.num_attributes(tr, te)
    group = (np.arange(p) >= I).astype(np.uint32)
    K, T = 8, 12
    dev = []
    for seed in range(16):
        m = sbmf.FmModel(p, K, attr_group=group, seed=100 + seed)
        m.set_train(tr)
        m.set_test(te)
        m.init()
        m.learn(T)
        dev.append(m.rmse_history(0, T)[1])
        if seed == 0:
            m.init()
            m.learn(T)
            assert np.array_equal(dev[0], m.rmse_history(0, T)[1]), "same seed, different chain"
        m.close()
    ref = []
    for seed in range(24):
        o = fmo.FmOracle(tr, te, K, num_attr=p, attr_group=group)
        o.srand(1000 + seed)
        o.init()
        ref.append(o.learn(T)[1])
        o.close()
    dev, ref = np.array(dev), np.array(ref)
    diff = np.abs(dev.mean(0) - ref.mean(0))
    se = np.sqrt(dev.var(0, ddof=1) / dev.shape[0] + ref.var(0, ddof=1) / ref.shape[0])
    assert np.all(diff[2:] <= 0.003), (diff, dev.mean(0), ref.mean(0))
    assert np.all(diff[:2] <= 4 * se[:2] + 0.003), (diff, se)
    assert np.all(dev.std(0)[2:] <= 3 * ref.std(0)[2:] + 1e-3)


def case_live_small():
    """live sampling on the small general fixture (also what the CPU execution of the kernels runs): the Philox-driven chains have
    the mean test-RMSE trajectory of libFM's rand()-driven ones within 4 standard errors (+0.003), and one seed repeats bit for bit"""
    tr, te, group = load_fixture("fm_general")
    p = fmo.num_attributes(tr, te)
    K, T, nd, nr = 3, 6, int(os.environ.get("FM_LIVE_SEEDS", "16")), 64
    dev = []
    for seed in range(nd):
        m = sbmf.FmModel(p, K, attr_group=group, seed=200 + seed)
        m.set_train(tr)
        m.set_test(te)
        m.init()
        m.learn(T)
        dev.append(m.rmse_history(0, T)[1])
        if seed == 0:
            m.init()
            m.learn(T)
            assert np.array_equal(dev[0], m.rmse_history(0, T)[1]), "same seed, different chain"
        m.close()
    ref = []
    for seed in range(nr):
        o = fmo.FmOracle(tr, te, K, num_attr=p, attr_group=group)
        o.srand(3000 + seed)
        o.init()
        ref.append(o.learn(T)[1])
        o.close()
    dev, ref = np.array(dev), np.array(ref)
    assert np.std(dev[:, -1]) > 1e-4, "no noise in live mode?"
    diff = np.abs(dev.mean(0) - ref.mean(0))
    se = np.sqrt(dev.var(0, ddof=1) / dev.shape[0] + ref.var(0, ddof=1) / ref.shape[0])
    assert np.all(diff <= 4 * se + 0.003), (diff, se, dev.mean(0), ref.mean(0))
    assert np.all(dev.std(0) <= 3 * ref.std(0) + 1e-3) and np.all(dev.std(0) >= ref.std(0) / 3 - 1e-3), (dev.std(0), ref.std(0))


def case_errors():
    tr, te, group = load_fixture("fm_general")
    p = fmo.num_attributes(tr, te)
    m = sbmf.FmModel(p, 3, attr_group=group)
    for call in (lambda: m.learn(1), lambda: m.init()):
        try:
            call()
            raise AssertionError("expected SBMF_ERR_STATE")
        except sbmf.SbmfError as e:
            assert e.code == -4, e
    bad = dict(tr)
    bad["attr"] = tr["attr"].copy()
    bad["attr"][0] = p
    try:
        m.set_train(bad)
        raise AssertionError("attribute id out of range accepted")
    except sbmf.SbmfError as e:
        assert e.code == -1, e
    dup = dict(tr)
    dup["attr"] = tr["attr"].copy()
    dup["attr"][1] = dup["attr"][0]                 # the first case lists one attribute twice
    try:
        m.set_train(dup)
        raise AssertionError("duplicate attribute in a case accepted")
    except sbmf.SbmfError as e:
        assert e.code == -1 and "twice" in str(e), e
    m.set_train(tr)
    m.init()
    m.learn(2)                                      # no test set: the test RMSE is NaN like libFM's 0 / 0, the train RMSE is not
    a, b = m.rmse_history(0, 2)
    assert np.all(np.isfinite(a)) and np.all(np.isnan(b))
    m.close()


def _preload_env(**extra):
    """child processes that LINK libsbmf_cuda.so (the host CLI, libFM with the CUDA learner) follow this process onto the CPU
    execution of the kernels when the tests run there (SBMF_FM_LIB_PATH, tests/test_fm_simt_emulation.py)"""
    env = dict(os.environ, **extra)
    if os.environ.get("SBMF_FM_LIB_PATH"):
        env["LD_PRELOAD"] = os.environ["SBMF_FM_LIB_PATH"]
    return env


def case_libfm_learner():
    """oracle/_ref/libFM_cuda = libFM's own main(), loader, meta groups, initial draws, predict() and -out around fm_learn_cuda::learn
    (oracle/libfm_cuda_learner.h): its "#Iter=" lines and prediction file are those of the binding started from
    the same initial w, v (libFM's rand() draws under SBMF_SHIM_SEED, reproduced by the pinned restatement)"""
    import re
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "libFM_cuda")
    if not os.path.exists(exe):
        print("skipped: oracle/_ref/libFM_cuda not built (needs /root/reference at build time)")
        return
    tr, te, group = load_fixture("fm_general")
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([exe, "-task", "r", "-train", os.path.join(GOLDEN, "fm_general.train"), "-test", os.path.join(GOLDEN, "fm_general.test"),
                            "-meta", os.path.join(GOLDEN, "fm_general.meta"), "-dim", "1,1,3", "-iter", "5", "-method", "mcmc", "-out", "pred.txt"],
                           capture_output=True, text=True, cwd=tmp, env=_preload_env(SBMF_SHIM_SEED="7"))
        assert r.returncode == 0 and "ERROR" not in r.stderr, r.stderr[-2000:]
        rows = re.findall(r"^#Iter=\s*(\d+)\tTrain=(\S+)\tTest=(\S+)$", r.stdout, flags=re.M)
        assert len(rows) == 5, r.stdout[-1000:]
        got = np.array([float(x) for x in open(os.path.join(tmp, "pred.txt")).read().split()])   # libFM's own predict() / DVector::save
    o = fmo.FmOracle(tr, te, 3, attr_group=group)
    o.srand(7)
    o.init()
    s0 = o.state()
    m = sbmf.FmModel(o.p, 3, attr_group=group)
    m.set_train(tr)
    m.set_test(te)
    m.init(s0["w"].astype(np.float32), s0["v"].astype(np.float32))
    m.learn(5)
    a, b = m.rmse_history(0, 5)
    assert [x[1] for x in rows] == [f"{v:g}" for v in a] and [x[2] for x in rows] == [f"{v:g}" for v in b], (rows, a, b)
    assert got.size == te["y"].size and np.max(np.abs(got - m.predict())) <= 2e-5
    m.close()
    o.close()


def case_cli():
    """bin/sbmf -method fm_mcmc: libFM's command line and outputs ("#Iter=" lines, test_rmse_<k0><k1><K>_mcmc, -out) carry the chain of
    the binding with the same seed, value for value"""
    import re
    import subprocess
    import tempfile
    tr, te, group = load_fixture("fm_general")
    p = fmo.num_attributes(tr, te)
    m = sbmf.FmModel(p, 3, attr_group=group, seed=3)
    m.set_train(tr)
    m.set_test(te)
    m.init()
    m.learn(5)
    a, b = m.rmse_history(0, 5)
    pred = m.predict()
    m.close()
    cli = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200", "bin", "sbmf")
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([cli, "-method", "fm_mcmc", "-train", os.path.join(GOLDEN, "fm_general.train"), "-test", os.path.join(GOLDEN, "fm_general.test"),
                            "-meta", os.path.join(GOLDEN, "fm_general.meta"), "-dim", "1,1,3", "-iter", "5", "-seed", "3", "-out", "pred.txt"],
                           capture_output=True, text=True, cwd=tmp, env=_preload_env())
        assert r.returncode == 0, r.stderr
        rows = re.findall(r"^#Iter=\s*(\d+)\tTrain=(\S+)\tTest=(\S+)$", r.stdout, flags=re.M)
        assert [x[1] for x in rows] == [f"{v:g}" for v in a] and [x[2] for x in rows] == [f"{v:g}" for v in b], (rows, a, b)
        assert open(os.path.join(tmp, "test_rmse_113_mcmc")).read().split() == [f"{v:g}" for v in b]
        assert open(os.path.join(tmp, "pred.txt")).read().split() == [f"{float(v):g}" for v in pred]


CASES = {k[5:]: v for k, v in dict(globals()).items() if k.startswith("case_")}

if __name__ == "__main__":
    CASES[sys.argv[1]]()
    print("ok", sys.argv[1])

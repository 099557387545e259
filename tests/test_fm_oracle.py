"""CPU: the restatement of libFM's MCMC learner (oracle/fm_oracle.c) is pinned to the UNMODIFIED reference libFM
(oracle/_ref/libFM_shim; golden files written by tests/golden/make_fm_golden.py): every "#Iter= i Train= Test=" value as printed,
and the sha256 of the complete sampler-argument stream, on a matrix-factorisation fixture and on a general design matrix
(real-valued, multi-hot, grouped attributes, an attribute the train file never mentions), live / zero-noise / -method als."""
import hashlib
import json
import os

import numpy as np
import pytest

import fm_oracle_py as fmo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    tr = fmo.read_libfm(os.path.join(GOLDEN, name + ".train"))
    te = fmo.read_libfm(os.path.join(GOLDEN, name + ".test"))
    meta = os.path.join(GOLDEN, name + ".meta")
    group = np.loadtxt(meta, dtype=np.uint32) if os.path.exists(meta) else None
    return tr, te, group


def make_oracle(name, g, **kw):
    tr, te, group = load_fixture(name)
    if g["mode"] == "als":     # [L]:131-136 + -regular 0.25,1,4
        kw.update(do_sample=0, do_multilevel=0, reg=(0.25, 1.0, 4.0))
    if g["mode"] == "zero":
        kw.update(noise=fmo.NOISE_ZERO)
    o = fmo.FmOracle(tr, te, g["K"], attr_group=group, **kw)
    assert o.p == g["num_attr"]
    return o


@pytest.mark.parametrize("mode", ["live", "zero", "als"])
@pytest.mark.parametrize("name", ["tiny_libfm", "fm_general"])
def test_oracle_reproduces_unmodified_libfm(name, mode, tmp_path):
    g = json.load(open(os.path.join(GOLDEN, f"libfm_{name}_{mode}.json")))
    o = make_oracle(name, g)
    log = str(tmp_path / "args.bin")
    o.set_log(log)
    o.srand(g["seed"])
    o.init()
    rtr, rte = o.learn(g["iters"])
    o.set_log(None)
    assert [f"{v:g}" for v in rtr] == g["train"]
    assert [f"{v:g}" for v in rte] == g["test"]
    raw = open(log, "rb").read()
    assert len(raw) // 24 == g["n_sampler_calls"]
    assert hashlib.sha256(raw).hexdigest() == g["sha256_args"]
    o.close()


def test_columns_are_the_stable_transpose():
    """Data.h:472-528: column i lists its cases in ascending order with the values of the row form"""
    tr, te, group = load_fixture("fm_general")
    o = fmo.FmOracle(tr, te, 2, attr_group=group)
    c = o.columns()
    assert c["col_ptr"][-1] == tr["row_ptr"][-1] and np.all(np.diff(c["col_ptr"]) >= 0)
    dense = np.zeros((o.n, o.p), dtype=np.float32)
    for r in range(o.n):
        for k in range(tr["row_ptr"][r], tr["row_ptr"][r + 1]):
            dense[r, tr["attr"][k]] = tr["x"][k]
    for i in range(o.p):
        cases = c["case"][c["col_ptr"][i]:c["col_ptr"][i + 1]]
        assert np.all(np.diff(cases.astype(np.int64)) > 0)
        assert np.array_equal(c["x"][c["col_ptr"][i]:c["col_ptr"][i + 1]], dense[cases, i])
    o.close()


def test_chain_continues_across_learn_calls():
    tr, te, group = load_fixture("fm_general")
    a = fmo.FmOracle(tr, te, 3, attr_group=group)
    b = fmo.FmOracle(tr, te, 3, attr_group=group)
    a.srand(3); a.init(); ra = a.learn(6)
    b.srand(3); b.init(); rb1 = b.learn(2); rb2 = b.learn(4)
    assert np.array_equal(ra[1], np.concatenate([rb1[1], rb2[1]]))
    assert np.array_equal(a.state()["v"], b.state()["v"])
    a.close(); b.close()

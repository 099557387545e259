"""Checkpoint / resume (sbmf_cuda_set_state, *_pred_sum, sbmf_cuda_checkpoint_*; SURVEY.md 5 -- the reference has no
checkpointing on this path, so the contract is self-consistency: a chain interrupted at a sweep boundary and restored in a
NEW handle continues like the uninterrupted one, because every draw is a function of (seed, site, row, k, sweep)).

CPU part: the checkpoint file format round-trips (pure host code of the C-ABI library).
GPU part (-m gpu): 3 + 3 sweeps through a checkpoint equal 6 sweeps straight."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200")


@pytest.fixture(scope="module")
def sbmf_mod():
    subprocess.run(["make", "-C", PKG], check=True, capture_output=True)
    import sbmf
    return sbmf


def fake_state(sbmf, I, J, K, N, rs, with_E=True):
    arr = {k: rs.standard_normal(shape).astype(dt) for k, (shape, dt) in sbmf._state_shapes(I, J, K, N).items()}
    if not with_E:
        del arr["E"]
    arr.update(b_0=3.5, alpha=1.25, mu_b_0=-0.5, sigma_b_0=2.0, sum_e=-12.5, sum_e2=99.0, sweeps_done=7)
    return arr


@pytest.mark.parametrize("with_E,with_ps", [(True, True), (False, True), (True, False), (False, False)])
def test_checkpoint_file_round_trip(sbmf_mod, tmp_path, with_E, with_ps):
    sbmf = sbmf_mod
    I, J, K, N, Nt = 11, 7, 5, 40, 9
    rs = np.random.RandomState(3)
    arr = fake_state(sbmf, I, J, K, N, rs, with_E)
    ps = rs.standard_normal(Nt) if with_ps else None
    path = str(tmp_path / "c.ckpt")
    sbmf.checkpoint_write(path, arr, I, J, K, N, Nt, hyper_mode=2, pred_sum=ps, seed=2 ** 40 + 5, sample_mode=1, burn_in=3, residual_mode=1,
                          rebuild_every=4)
    assert not os.path.exists(path + ".tmp")
    dims, got, ps2 = sbmf.checkpoint_read(path)
    assert (dims["num_users"], dims["num_items"], dims["K"], dims["n_train"], dims["n_test"], dims["hyper_mode"]) == (I, J, K, N, Nt, 2)
    # the flags of the chain that wrote the state travel with it (a resume under other flags is a different chain)
    assert (dims["seed"], dims["sample_mode"], dims["burn_in"], dims["residual_mode"], dims["rebuild_every"]) == (2 ** 40 + 5, 1, 3, 1, 4)
    assert dims["sweeps_done"] == 7
    assert ((dims["present"] >> 12) & 1) == int(with_E) and ((dims["present"] >> 13) & 1) == int(with_ps)
    for k in sbmf.STATE_ARRAYS:
        if k == "E" and not with_E:
            assert "E" not in got
        else:
            assert np.array_equal(got[k], arr[k]), k
    for k in sbmf.STATE_SCALARS:
        assert got[k] == arr[k], k
    assert (ps2 is None) == (ps is None)
    if ps is not None:
        assert np.array_equal(ps2, ps)
    # header 128 bytes + exactly the present arrays
    want = 128 + 4 * (I * K + K * J + 3 * I + 3 * J) + 8 * 4 * K + (4 * N if with_E else 0) + (8 * Nt if with_ps else 0)
    assert os.path.getsize(path) == want


def test_checkpoint_rejects_garbage_and_truncation(sbmf_mod, tmp_path):
    sbmf = sbmf_mod
    bad = tmp_path / "bad.ckpt"
    bad.write_bytes(b"not a checkpoint at all" * 10)
    with pytest.raises(sbmf.SbmfError):
        sbmf.checkpoint_read(str(bad))
    with pytest.raises(sbmf.SbmfError):
        sbmf.checkpoint_read(str(tmp_path / "missing.ckpt"))
    I, J, K, N, Nt = 4, 3, 2, 6, 2
    arr = fake_state(sbmf, I, J, K, N, np.random.RandomState(0))
    path = str(tmp_path / "t.ckpt")
    sbmf.checkpoint_write(path, arr, I, J, K, N, Nt, pred_sum=np.zeros(Nt))
    data = open(path, "rb").read()
    open(path, "wb").write(data[:-5])
    with pytest.raises(sbmf.SbmfError) as e:
        sbmf.checkpoint_read(path)
    assert "truncated" in str(e.value)
    with pytest.raises(sbmf.SbmfError):   # unwritable target
        sbmf.checkpoint_write(str(tmp_path / "no_such_dir" / "x.ckpt"), arr, I, J, K, N, Nt)


def test_checkpoint_read_refuses_buffers_of_other_dimensions(sbmf_mod, tmp_path):
    """sbmf_cuda_checkpoint_read copies with the sizes in the FILE: a caller whose buffers were sized for another problem must be
    refused before anything is written (round-1 advisor finding: heap overflow when read_dims was skipped)."""
    import ctypes as C
    sbmf = sbmf_mod
    I, J, K, N, Nt = 9, 6, 4, 30, 5
    arr = fake_state(sbmf, I, J, K, N, np.random.RandomState(1))
    path = str(tmp_path / "big.ckpt")
    sbmf.checkpoint_write(path, arr, I, J, K, N, Nt, pred_sum=np.ones(Nt))
    lib = sbmf.load_library()
    for wrong in ({"num_users": 3}, {"num_items": 2}, {"K": 2}, {"n_train": 7}, {"n_test": 1}):
        want = dict(num_users=I, num_items=J, K=K, n_train=N, n_test=Nt)
        want.update(wrong)
        # buffers of the (smaller) size the caller believes in, with a canary after each
        small = {k: np.full(int(np.prod(shape)) + 8, 7.0, dt) for k, (shape, dt) in
                 sbmf._state_shapes(want["num_users"], want["num_items"], want["K"], want["n_train"]).items()}
        st = sbmf.State()
        for k, v in small.items():
            setattr(st, k, v.ctypes.data)
        ps = np.full(want["n_test"] + 8, 7.0)
        dims = sbmf.CheckpointDims(want["num_users"], want["num_items"], want["K"], 0, want["n_train"], want["n_test"], 0, 0, 0, 0, 0, 0, 0)
        have = C.c_int(0)
        rc = lib.sbmf_cuda_checkpoint_read(os.fsencode(path), C.byref(dims), C.byref(st), ps.ctypes.data, C.byref(have))
        assert rc == -1, wrong
        assert b"caller's buffers" in lib.sbmf_cuda_checkpoint_last_error()
        assert all(np.all(v == 7.0) for v in small.values()) and np.all(ps == 7.0)     # nothing was touched
    assert lib.sbmf_cuda_checkpoint_read(os.fsencode(path), None, None, None, None) == -1


# ------------------------------------------------------------------------------------------------------------ GPU
def _model(sbmf, d, K, **cfg):
    m = sbmf.SbmfModel(K=K, **cfg)
    m.set_train(d["train_user"], d["train_item"], d["train_rating"], d["num_users"], d["num_items"])
    m.set_test(d["test_user"], d["test_item"], d["test_rating"])
    return m


def _max_rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("mode,hyper,residual_mode", [(0, 0, 0), (2, 0, 0), (0, 0, 1), (0, 2, 0)])
def test_resume_continues_the_chain(sbmf_mod, ml100k, tmp_path, mode, hyper, residual_mode):
    """6 sweeps straight vs 3 sweeps -> checkpoint file -> new handle -> 3 sweeps.  Same Philox draws on both sides, the residual
    and the prediction sums travel with the checkpoint: states agree to fp32 rounding (tolerance 1e-5 of each array's scale),
    the RMSE of the running mean to 1e-7."""
    sbmf = sbmf_mod
    d, K = ml100k, 20
    I, J, N, Nt = d["num_users"], d["num_items"], d["train_user"].size, d["test_user"].size
    cfg = dict(sample_mode=mode, hyper_mode=hyper, residual_mode=residual_mode, seed=11)
    rs = np.random.RandomState(5)
    U0 = (0.1 * rs.standard_normal((I, K))).astype(np.float32)
    V0 = (0.1 * rs.standard_normal((K, J))).astype(np.float32)

    a = _model(sbmf, d, K, **cfg)
    a.init_factors(U0, V0)
    a.sweep(6)
    want, want_ps = a.get_state(), a.get_pred_sum()
    want_r, want_rs = a.rmse_history(3, 3)
    a.close()

    b = _model(sbmf, d, K, **cfg)
    b.init_factors(U0, V0)
    b.sweep(3)
    path = str(tmp_path / "mid.ckpt")
    sbmf.checkpoint_write(path, b.get_state(), I, J, K, N, Nt, hyper_mode=hyper, pred_sum=b.get_pred_sum())
    b.close()

    dims, st, ps = sbmf.checkpoint_read(path)
    assert dims["sweeps_done"] == 3 and ps is not None
    c = _model(sbmf, d, K, **cfg)
    c.set_state(st)
    c.set_pred_sum(ps)
    c.sweep(3)
    got, got_ps = c.get_state(), c.get_pred_sum()
    got_r, got_rs = c.rmse_history(3, 3)
    c.close()

    assert got["sweeps_done"] == 6
    worst = {k: _max_rel(got[k], want[k]) for k in sbmf.STATE_ARRAYS}
    for k in ("b_0", "alpha", "mu_b_0", "sigma_b_0", "sum_e2"):
        worst[k] = abs(got[k] - want[k]) / max(abs(want[k]), 1e-30)
    bad = {k: v for k, v in worst.items() if not v <= 1e-5}
    assert not bad, f"resumed chain differs: {bad} (all: {worst})"
    assert np.max(np.abs(got_r - want_r)) <= 1e-7 and np.max(np.abs(got_rs - want_rs)) <= 1e-6, (got_r, want_r)
    assert _max_rel(got_ps, want_ps) <= 1e-7


@pytest.mark.gpu
def test_resume_without_residual_rebuilds_it(sbmf_mod, ml100k):
    """A checkpoint without E: the next sweep rebuilds the residual stand-alone ([T]:342-359) -- equal to the carried one up to
    fp32 rounding, so the chain agrees within the parity tolerance (1e-4) instead of 1e-5."""
    sbmf = sbmf_mod
    d, K = ml100k, 20
    cfg = dict(sample_mode=2, seed=4)
    a = _model(sbmf, d, K, **cfg)
    a.init_factors()
    a.sweep(5)
    want = a.get_state()
    a.close()
    b = _model(sbmf, d, K, **cfg)
    b.init_factors()
    b.sweep(2)
    st, ps = b.get_state(with_E=False), b.get_pred_sum()
    b.close()
    c = _model(sbmf, d, K, **cfg)
    c.set_state(st)
    c.set_pred_sum(ps)
    c.sweep(3)
    got = c.get_state()
    c.close()
    for k in ("U", "V", "b_i", "b_j", "E"):
        assert _max_rel(got[k], want[k]) <= 1e-4, k


@pytest.mark.gpu
def test_set_state_argument_checks(sbmf_mod, ml100k):
    sbmf = sbmf_mod
    d, K = ml100k, 8
    m = sbmf.SbmfModel(K=K)
    m.I, m.J, m.N = 3, 2, 0
    with pytest.raises(sbmf.SbmfError) as e:   # before set_train
        m.set_state({"U": np.zeros((3, K), np.float32), "V": np.zeros((K, 2), np.float32)})
    assert e.value.code == -4
    m.close()
    m = _model(sbmf, d, K)
    with pytest.raises(sbmf.SbmfError) as e:   # U / V missing
        m.set_state({"U": np.zeros((m.I, K), np.float32)})
    assert e.value.code == -1
    m.close()


@pytest.mark.gpu
def test_cli_save_and_load_state(sbmf_mod, ml100k, tmp_path):
    """bin/sbmf -iter 6 prints the same `rmse is` lines as -iter 2 -save_state followed by -load_state -iter 4 (the CLI turns
    per-phase timing off, so the last sweeps of both runs are CUDA-graph replays: a resumed handle must not capture before its
    lazily built maps exist)."""
    d = ml100k
    tr, te = tmp_path / "tr", tmp_path / "te"
    np.savetxt(tr, np.c_[d["train_user"], d["train_item"], d["train_rating"]], fmt="%d\t%d\t%g")
    np.savetxt(te, np.c_[d["test_user"], d["test_item"], d["test_rating"]], fmt="%d\t%d\t%g")
    cli = os.path.join(PKG, "bin", "sbmf")
    base = [cli, "-train", str(tr), "-test", str(te), "-seed", "9"]
    dim = ["-dim", "1,1,12"]

    def rmses(extra):
        r = subprocess.run(base + dim + extra, capture_output=True, text=True, cwd=tmp_path)
        assert r.returncode == 0, r.stderr
        return [float(l.split()[-1]) for l in r.stdout.splitlines() if l.startswith("rmse is")]

    straight = rmses(["-iter", "6", "-out", "p4.txt"])
    first = rmses(["-iter", "2", "-save_state", "s.ckpt"])
    second = rmses(["-iter", "4", "-load_state", "s.ckpt", "-out", "p22.txt"])
    assert len(straight) == 6 and len(first) == 2 and len(second) == 4
    assert np.allclose(first + second, straight, rtol=0, atol=2e-6), (first, second, straight)
    p4, p22 = np.loadtxt(tmp_path / "p4.txt"), np.loadtxt(tmp_path / "p22.txt")
    assert np.max(np.abs(p4 - p22)) <= 1e-4
    r = subprocess.run(base + ["-iter", "1", "-dim", "1,1,13", "-load_state", "s.ckpt"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "another problem" in r.stderr
    # same problem, other seed: the chain would not be the one that was interrupted
    r = subprocess.run([cli, "-train", str(tr), "-test", str(te), "-seed", "10"] + dim + ["-iter", "1", "-load_state", "s.ckpt"],
                       capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "other flags" in r.stderr


@pytest.mark.gpu
def test_rmse_history_grows_past_its_first_allocation(sbmf_mod, tiny):
    """The history starts with room for 4096 sweeps and grows with the chain (it used to return an error beyond that); the graph
    replay survives the reallocation."""
    sbmf = sbmf_mod
    m = _model(sbmf, tiny, 3, sample_mode=0, seed=2)
    m.init_factors()
    m.set_timing_enabled(0)
    m.sweep(4090)
    m.sweep(20)
    r, rs = m.rmse_history(0, 4110)
    assert np.all(np.isfinite(r)) and np.all(r > 0) and np.all(rs > 0)
    m2 = _model(sbmf, tiny, 3, sample_mode=0, seed=2)
    m2.init_factors()
    m2.sweep(5)
    assert np.array_equal(m2.rmse_history(0, 5)[0], r[:5])
    m.close(); m2.close()

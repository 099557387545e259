"""Host program (bin/sbmf, csrc/main.cpp): libFM flag grammar and error behaviour on CPU; on the GPU the same trajectory and
prediction file as the ctypes binding, from triple files and from libFM text files."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200")
CLI = os.path.join(PKG, "bin", "sbmf")


@pytest.fixture(scope="module")
def cli():
    subprocess.run(["make", "-C", PKG], check=True, capture_output=True)
    return CLI


def run(cli, *args, cwd=None):
    return subprocess.run([cli, *args], capture_output=True, text=True, cwd=cwd)


def test_help_lists_libfm_flags(cli):
    r = run(cli, "-help")
    assert r.returncode == 0
    for flag in ("-train", "-test", "-dim", "-iter", "-out", "-rlog", "-seed", "-init_stdev", "-method", "-task", "-verbosity"):
        assert flag in r.stdout


def test_unknown_and_duplicate_flags_error_like_libfm(cli):
    r = run(cli, "-bogus", "1")
    assert r.returncode == 1 and "ERROR: the parameter bogus does not exist" in r.stderr
    r = run(cli, "-iter", "1", "--iter", "2")
    assert r.returncode == 1 and "ERROR: the parameter iter is already specified" in r.stderr
    r = run(cli, "stray")
    assert r.returncode == 1 and "ERROR: cannot parse stray" in r.stderr


def test_missing_and_malformed_input(cli, tmp_path):
    r = run(cli, "-train", str(tmp_path / "nope"), "-test", str(tmp_path / "nope"))
    assert r.returncode == 1 and "unable to open" in r.stderr
    bad = tmp_path / "bad"
    bad.write_text("0\t1\t3\nthis is not a rating\n")
    r = run(cli, "-train", str(bad), "-test", str(bad))
    assert r.returncode == 1 and "malformed rating line 2" in r.stderr
    r = run(cli, "-dim", "1,0,8", "-train", str(bad), "-test", str(bad))
    assert r.returncode == 1 and "k0 and k1 must be 1" in r.stderr


def write_triples(path, u, i, r):
    with open(path, "w") as f:
        for a, b, c in zip(u, i, r):
            f.write(f"{a}\t{b}\t{int(c)}\n")


def write_libfm(path, u, i, r, offset):
    with open(path, "w") as f:
        for a, b, c in zip(u, i, r):
            f.write(f"{int(c)} {a}:1 {b + offset}:1\n")


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["triples", "libfm"])
def test_cli_matches_binding(cli, tmp_path, ml100k, fmt):
    import sbmf
    d = ml100k
    tr, te = str(tmp_path / "train"), str(tmp_path / "test")
    if fmt == "triples":
        write_triples(tr, d["train_user"], d["train_item"], d["train_rating"])
        write_triples(te, d["test_user"], d["test_item"], d["test_rating"])
    else:
        write_libfm(tr, d["train_user"], d["train_item"], d["train_rating"], d["num_users"])
        write_libfm(te, d["test_user"], d["test_item"], d["test_rating"], d["num_users"])
    out, rlog = str(tmp_path / "pred"), str(tmp_path / "rlog")
    r = run(cli, "-task", "r", "-train", tr, "-test", te, "-dim", "1,1,20", "-iter", "6", "-seed", "42", "-out", out, "-rlog", rlog)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[:3] == ["number rows =90570", "number of user =943", "number of items =1682"]
    got = [ln.split()[-1] for ln in lines if ln.startswith("rmse is ")]
    m = sbmf.SbmfModel(K=20, seed=42)
    m.set_train(d["train_user"], d["train_item"], d["train_rating"], d["num_users"], d["num_items"])
    m.set_test(d["test_user"], d["test_item"], d["test_rating"])
    m.init_factors()
    m.sweep(6)
    want = ["%g" % x for x in m.rmse_history(0, 6)[0]]
    assert got == want
    pred = np.loadtxt(out)
    assert pred.shape == (9430,) and np.max(np.abs(pred - m.get_pred())) < 1e-5
    assert pred.min() >= 0.5 and pred.max() <= 5.0
    log = open(rlog).read().splitlines()
    assert log[0].split("\t") == ["rmse", "rmse_mcmc_this", "rmse_mcmc_all", "time_learn", "alpha", "b_0"] and len(log) == 7
    # -method mcmc: libFM's front-end outputs (fm_learn_mcmc_simultaneous.h:57-62, 244-245) instead of [T]'s "rmse is" lines
    r2 = run(cli, "-task", "r", "-train", tr, "-test", te, "-dim", "1,1,20", "-iter", "6", "-seed", "42", "-method", "mcmc", cwd=str(tmp_path))
    assert r2.returncode == 0, r2.stderr
    it = [ln for ln in r2.stdout.splitlines() if ln.startswith("#Iter=")]
    assert len(it) == 6 and it[0].startswith("#Iter=  0\tTrain=") and [ln.split("Test=")[1] for ln in it] == want
    assert "rmse is" not in r2.stdout
    assert open(tmp_path / "test_rmse_1120_mcmc").read().split() == want
    m.close()


@pytest.mark.gpu
def test_cli_no_arguments_uses_reference_paths(cli, tmp_path, tiny):
    """[T] opens ../../data/ra.{train,test}_sbpmf relative to its CWD ([T]:32, 98), D=20, T=100."""
    (tmp_path / "data").mkdir()
    (tmp_path / "a" / "b").mkdir(parents=True)
    d = tiny
    write_triples(tmp_path / "data" / "ra.train_sbpmf", d["train_user"], d["train_item"], d["train_rating"])
    write_triples(tmp_path / "data" / "ra.test_sbpmf", d["test_user"], d["test_item"], d["test_rating"])
    r = run(cli, cwd=str(tmp_path / "a" / "b"))
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    assert lines[1] == "number of user =50" and lines[2] == "number of items =40"
    assert sum(ln.startswith("rmse is ") for ln in lines) == 100


def test_three_input_formats_parse_to_the_same_triples(cli, tmp_path):
    """triples, libFM text and libFM binary (bytes written by the reference's own convert tool) -> identical (user, item, rating)."""
    G = os.path.join(ROOT, "tests", "golden")
    want = open(os.path.join(G, "tiny_unsorted.train")).read().split()
    want = [(int(want[i]), int(want[i + 1]), float(want[i + 2])) for i in range(0, len(want), 3)]
    outs = {}
    for name, tr, te in (("triples", "tiny_unsorted.train", "tiny_unsorted.test"), ("text", "tiny_libfm.train", "tiny_libfm.test"),
                         ("binary", "tiny_libfm.train_bin", "tiny_libfm.test_bin")):
        dump = str(tmp_path / f"dump_{name}")
        r = run(cli, "-train", os.path.join(G, tr), "-test", os.path.join(G, te), "-dry_run", "1", "-dump_triples", dump)
        assert r.returncode == 0, r.stderr
        assert r.stdout.splitlines()[:3] == ["number rows =560", "number of user =50", "number of items =40"], r.stdout
        got = open(dump).read().split()
        outs[name] = [(int(got[i]), int(got[i + 1]), float(got[i + 2])) for i in range(0, len(got), 3)]
    assert outs["triples"] == want and outs["text"] == want and outs["binary"] == want


def test_binary_header_checks(cli, tmp_path):
    import struct
    p = str(tmp_path / "bad")
    open(p + ".x", "wb").write(struct.pack("<IIQII", 3, 4, 2, 1, 2))
    open(p + ".y", "wb").write(struct.pack("<III", 1, 4, 1))
    r = run(cli, "-train", p, "-test", p, "-dry_run", "1")
    assert r.returncode == 1 and "file id != 2" in r.stderr


# ------------------------------------------------------------------------------------------ text reader: sscanf semantics, fast path
ODD_TRIPLES = [
    "0\t0\t3.5", "1 2 4", "  2   3    5  ", "3,4,1.5 trailing words", "4;5;2e0", "5\t6\t+3.0", "6 7 0x1p1", "7 8 .5", "9 10 4.\r",
    "", "   \t ", "10 11 1e-1", "0000012 13 3.25", "+13 14 2", "14 +15 2.5", "15 16 -1.5", "16 17 3.5abc", "17 18 1e5f",
    "4000000000 19 1", "20 4000000001 2", "123456789 21 3", "22:23:4", "23 24 inf",
]


def test_text_reader_fast_path_equals_sscanf(cli, tmp_path):
    """The hand-written line parser (csrc/rating_reader.h) must accept exactly what sscanf("%u%c%u%c%lf") / ("%lf %u:%lf %u:%lf")
    accepts and produce the same values: the same files through SBMF_SLOW_PARSER=1 (every line through sscanf) give the same dump,
    on odd spellings and on a multi-chunk file (> 1 MB, parsed by several threads)."""
    rs = np.random.RandomState(1)

    def both(train_lines, name):
        tr = tmp_path / f"{name}.train"
        tr.write_text("\n".join(train_lines) + "\n")
        te = tmp_path / f"{name}.test"
        te.write_text(train_lines[0] + "\n")
        outs = []
        for slow in (False, True):
            dump = str(tmp_path / f"{name}.{int(slow)}.dump")
            env = dict(os.environ, **({"SBMF_SLOW_PARSER": "1"} if slow else {}))
            env.pop("SBMF_SLOW_PARSER", None) if not slow else None
            r = subprocess.run([cli, "-train", str(tr), "-test", str(te), "-dry_run", "1", "-dump_triples", dump], capture_output=True, text=True, env=env)
            outs.append((r.returncode, r.stdout, r.stderr, open(dump).read() if os.path.exists(dump) else None))
        assert outs[0] == outs[1], name
        return outs[0]

    rc, out, err, dump = both(ODD_TRIPLES, "odd")
    assert rc == 0, err
    rows = [ln.split("\t") for ln in dump.splitlines()]
    assert len(rows) == len([ln for ln in ODD_TRIPLES if ln.strip()])        # every non-blank line is a rating for sscanf
    assert rows[0] == ["0", "0", "3.5"] and rows[3] == ["3", "4", "1.5"] and rows[6] == ["6", "7", "2"] and rows[19][:2] == ["22", "23"]
    libfm = ["3.5 0:1 60:1", " 4  1:1   61:1  ", "1e0 2:1 62:1 tail", "+2 3:1.0 63:1", "0x1p1 4:1 64:1", "2.5\t5:1\t65:1\r", "", "3 6:1 66:+1"]
    rc, out, err, dump = both(libfm, "libfm")
    assert rc == 0, err
    assert len(dump.splitlines()) == 7
    big = [f"{u}\t{i}\t{r}" for u, i, r in zip(np.sort(rs.randint(0, 5000, 150000)), rs.randint(0, 3000, 150000), rs.randint(1, 11, 150000) / 2.0)]
    big[70000] = "  " + big[70000].replace("\t", "  ") + " \r"
    rc, out, err, dump = both(big, "big")
    assert rc == 0 and out.splitlines()[0] == "number rows =150000", err
    # malformed lines are reported with the same (1-based, physical) line number by both parsers, also deep inside a large file
    for bad_at, bad in ((3, "1::2::3"), (100001, "12 x 3"), (149999, "7 8")):
        lines = list(big)
        lines[bad_at - 1] = bad
        rc, out, err, dump = both(lines, f"bad{bad_at}")
        assert rc == 1 and f"malformed rating line {bad_at} " in err, err


# ---- general FM Gibbs front-end (-method fm_mcmc | fm_als, csrc/fm_main.h): the readers, on CPU ----------------------------------
def _dump(cli, tmp_path, *args):
    out = str(tmp_path / "design.txt")
    r = run(cli, *args, "-dry_run", "1", "-dump_design", out, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    return r.stdout, np.loadtxt(out, ndmin=2)


def test_fm_front_end_reads_general_libfm_text(cli, tmp_path):
    """the design matrix parsed by the host program == the reader the parity tests use (Data.h:184-278 semantics), attribute
    count = libFM's max id + 2 ([L]:326), groups from -meta"""
    import fm_oracle_py as fmo
    G = os.path.join(ROOT, "tests", "golden")
    tr = fmo.read_libfm(os.path.join(G, "fm_general.train"))
    te = fmo.read_libfm(os.path.join(G, "fm_general.test"))
    stdout, d = _dump(cli, tmp_path, "-method", "fm_mcmc", "-train", os.path.join(G, "fm_general.train"), "-test", os.path.join(G, "fm_general.test"),
                      "-meta", os.path.join(G, "fm_general.meta"), "-dim", "1,1,3")
    assert f"#cases train=600\ttest=150\t#attr={fmo.num_attributes(tr, te)}\t#groups=4" in stdout
    rows = np.repeat(np.arange(600), np.diff(tr["row_ptr"]))
    assert np.array_equal(d[:, 0], rows) and np.array_equal(d[:, 1], tr["attr"])
    assert np.array_equal(d[:, 2].astype(np.float32), tr["x"]) and np.array_equal(d[:, 3].astype(np.float32), tr["y"][rows])
    # comments, blank lines and trailing blanks are skipped like Data.h:196-197, 209-212; garbage is an error like libFM's throw
    t = tmp_path / "c.train"
    t.write_text("# header\n\n4.5 3:1 7:0.5  \n  2 0:1\t#tail\n")
    _, d = _dump(cli, tmp_path, "-method", "fm_als", "-train", str(t), "-test", str(t))
    assert d.tolist() == [[0, 3, 1, 4.5], [0, 7, 0.5, 4.5], [1, 0, 1, 2]]
    t.write_text("4.5 3:1 oops\n")
    r = run(cli, "-method", "fm_mcmc", "-train", str(t), "-test", str(t), "-dry_run", "1")
    assert r.returncode == 1 and "ERROR: cannot parse line" in r.stderr
    r = run(cli, "-method", "fm_mcmc", "-train", os.path.join(G, "fm_general.train"), "-test", os.path.join(G, "fm_general.test"), "-regular", "1,2",
            "-dry_run", "1")
    assert r.returncode == 1 and "-regular" in r.stderr


def test_fm_front_end_reads_general_libfm_binary(cli, tmp_path):
    """F.x / F.y in the format of fmatrix.h:34-52 (rows of any length) parse to the same design matrix as the text file"""
    import struct
    import fm_oracle_py as fmo
    G = os.path.join(ROOT, "tests", "golden")
    tr = fmo.read_libfm(os.path.join(G, "fm_general.train"))
    base = str(tmp_path / "bin.train")
    with open(base + ".x", "wb") as f:
        f.write(struct.pack("<IIQII", 2, 4, int(tr["row_ptr"][-1]), tr["y"].size, int(tr["attr"].max()) + 1))
        for r in range(tr["y"].size):
            a, b = tr["row_ptr"][r], tr["row_ptr"][r + 1]
            f.write(struct.pack("<I", b - a))
            for k in range(a, b):
                f.write(struct.pack("<If", int(tr["attr"][k]), float(tr["x"][k])))
    with open(base + ".y", "wb") as f:
        f.write(struct.pack("<III", 1, 4, tr["y"].size))
        f.write(tr["y"].astype("<f4").tobytes())
    _, d_bin = _dump(cli, tmp_path, "-method", "fm_mcmc", "-train", base, "-test", base)
    _, d_txt = _dump(cli, tmp_path, "-method", "fm_mcmc", "-train", os.path.join(G, "fm_general.train"), "-test", os.path.join(G, "fm_general.test"))
    assert np.array_equal(d_bin, d_txt)
    with open(base + ".x", "r+b") as f:
        f.truncate(200)
    r = run(cli, "-method", "fm_mcmc", "-train", base, "-test", base, "-dry_run", "1")
    assert r.returncode == 1 and "truncated" in r.stderr

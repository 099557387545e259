"""libFM's transposed binary design matrix (".xt"): the device-built CSR/CSC layout written in the format of
fmatrix.h:34-52 must be byte-identical to what the reference's own src/libfm/tools/transpose.cpp writes for the same ".x"
(golden file produced by tests/golden/make_libfm_fixtures.py with the unmodified tool).  CPU: the writer on the oracle's
layout.  GPU: the CLI's -dump_xt on the layout built on the device."""
import os
import subprocess

import numpy as np  # noqa: F401
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


def test_xt_writer_matches_reference_transpose(cli, tmp_path, tiny):
    """csrc/xt_writer.cpp on the (oracle's) CSR/CSC layout of the tiny fixture == the bytes the reference's own
    src/libfm/tools/transpose.cpp wrote for tiny_libfm.train_bin.x (tests/golden/make_libfm_fixtures.py).  Host code only."""
    import oracle_py as orc
    import sbmf
    d = tiny
    o = orc.Oracle(d["train_user"], d["train_item"], d["train_rating"], d["test_user"], d["test_item"], d["test_rating"],
                   d["num_users"], d["num_items"], 4)
    L = o.layout()
    golden = open(os.path.join(ROOT, "tests", "golden", "tiny_libfm.train_bin.xt"), "rb").read()
    out = str(tmp_path / "t.xt")
    nfeat = 50 + int(d["train_item"].max()) + 1          # item features start at 50 in the fixture; feature count of the TRAIN file
    sbmf.write_libfm_xt(out, L, nfeat, d["num_users"], d["num_items"], 50)
    assert open(out, "rb").read() == golden
    # a gap between the users and the item features becomes empty rows; an offset inside the user range is refused
    sbmf.write_libfm_xt(out, L, nfeat + 7, d["num_users"], d["num_items"], 57)
    import struct
    hdr = struct.unpack("<IIQII", open(out, "rb").read()[:24])
    assert hdr == (2, 4, 2 * d["train_user"].size, nfeat + 7, d["train_user"].size)
    assert os.path.getsize(out) == len(golden) + 7 * 4
    with pytest.raises(sbmf.SbmfError):
        sbmf.write_libfm_xt(out, L, nfeat, d["num_users"], d["num_items"], 10)
    with pytest.raises(sbmf.SbmfError):                  # too few feature rows for the rated items
        sbmf.write_libfm_xt(out, L, 60, d["num_users"], d["num_items"], 50)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["binary", "text", "triples"])
def test_cli_dump_xt_is_the_reference_transpose(cli, tmp_path, fmt):
    """-dump_xt writes the device-built CSR/CSC layout as libFM's .xt: byte-identical to the reference's transpose tool."""
    G = os.path.join(ROOT, "tests", "golden")
    tr, te = {"binary": ("tiny_libfm.train_bin", "tiny_libfm.test_bin"), "text": ("tiny_libfm.train", "tiny_libfm.test"),
              "triples": ("tiny_unsorted.train", "tiny_unsorted.test")}[fmt]
    out = str(tmp_path / "o.xt")
    r = run(cli, "-train", os.path.join(G, tr), "-test", os.path.join(G, te), "-dim", "1,1,4", "-iter", "1", "-dump_xt", out, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(G, "tiny_libfm.train_bin.xt"), "rb").read()

"""The reference-side binding of INTEGRATION.md section 1, made real: oracle/_ref/gibbs_ref_cuda is the UNMODIFIED head of the
reference's gibbs_sbpmf2.cpp (its includes and the file passes of its main(), [T]:1-153) spliced on a pipe with
oracle/ref_cuda_body.inc (the sbmf_cuda_* calls that replace [T]:156-666) and linked against libsbmf_cuda.so
(recipe: oracle/Makefile, target $(OUT)/gibbs_ref_cuda).

CPU: the binary exists, parses the fixture with the reference's own code and stops at sbmf_cuda_create (no CPU fallback).
GPU: its `rmse is` lines are those of the host CLI (same library, same defaults), line for line."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "oracle", "_ref", "gibbs_ref_cuda")
sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.fixture(scope="module")
def patched():
    subprocess.run(["make", "-C", PKG], check=True, capture_output=True)
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)   # keeps the prebuilt files without /root/reference
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/gibbs_ref_cuda not built (needs /root/reference at build time)")
    return BIN


def test_patched_reference_reaches_the_library(patched):
    if os.path.exists("/dev/nvidia0"):
        pytest.skip("GPU present: covered by the GPU test")
    from run_ref import run_ref
    r = run_ref(patched, os.path.join(GOLDEN, "tiny_unsorted.train"), os.path.join(GOLDEN, "tiny_unsorted.test"))
    assert (r["num_rows"], r["num_users"], r["num_items"]) == (560, 50, 40)     # parsed by the reference's own passes
    assert r["returncode"] == 1 and r["rmse"] == []                             # and refused by create: no device, no fallback


@pytest.mark.gpu
def test_patched_reference_prints_the_cli_trajectory(patched, tmp_path):
    from run_ref import run_ref
    tr, te = os.path.join(ROOT, "oracle", "_ref", "data", "m100k", "train_sbpmf"), os.path.join(ROOT, "oracle", "_ref", "data", "m100k", "test_sbpmf")
    r = run_ref(patched, tr, te)
    assert r["returncode"] == 0 and len(r["rmse_text"]) == 100, r["stdout"][-500:]
    assert (r["num_rows"], r["num_users"], r["num_items"]) == (90570, 943, 1682)
    cli = subprocess.run([os.path.join(PKG, "bin", "sbmf"), "-train", tr, "-test", te], capture_output=True, text=True, cwd=tmp_path)
    assert cli.returncode == 0, cli.stderr
    want = [ln.split()[-1] for ln in cli.stdout.splitlines() if ln.startswith("rmse is ")]
    assert r["rmse_text"] == want
    assert 0.95 < r["rmse"][-1] < 0.975     # the reference's own run ends at 0.9615 (tests/golden/ref_ml100k_K20_T100_rmse.txt)


# ---- libFM's own main() with its MCMC learner on the GPU: oracle/_ref/libFM_cuda (oracle/libfm_cuda_learner.h) ----
LIBFM_CUDA = os.path.join(ROOT, "oracle", "_ref", "libFM_cuda")
FM_ARGS = ["-task", "r", "-train", os.path.join(GOLDEN, "fm_general.train"), "-test", os.path.join(GOLDEN, "fm_general.test"),
           "-meta", os.path.join(GOLDEN, "fm_general.meta"), "-dim", "1,1,3", "-iter", "5", "-method", "mcmc"]


def test_patched_libfm_reaches_the_library(patched, tmp_path):
    if not os.path.exists(LIBFM_CUDA):
        pytest.skip("oracle/_ref/libFM_cuda not built (needs /root/reference at build time)")
    if os.path.exists("/dev/nvidia0"):
        pytest.skip("GPU present: covered by the GPU test")
    r = subprocess.run([LIBFM_CUDA] + FM_ARGS, capture_output=True, text=True, cwd=tmp_path, env=dict(os.environ, SBMF_SHIM_SEED="7"))
    assert "Loading train" in r.stdout and "#relations: 0" in r.stdout          # libFM's own loader and set-up ran
    assert "ERROR: sbmf_fm_create" in r.stderr and "no CPU fallback" in r.stderr and "#Iter" not in r.stdout


@pytest.mark.gpu
def test_patched_libfm_prints_the_binding_trajectory(patched):
    """tests/fm_gpu_cases.py::case_libfm_learner in a process of its own (the CPU suite runs the same case on the CPU execution of the
    kernels, tests/test_fm_simt_emulation.py)"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fm_gpu_cases.py"), "libfm_learner"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "ok libfm_learner" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

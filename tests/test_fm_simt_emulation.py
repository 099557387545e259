"""CPU: the KERNELS of the general FM Gibbs path (csrc/fm.cu), executed without a GPU.

csrc/fm.cu is compiled a second time with g++ against tools/emu_include (a host stand-in for the CUDA runtime and the SIMT execution
model: device memory = host memory, the threads of a CTA are host threads, __syncthreads / __shfl_xor_sync are barriers and an
exchange buffer, __shared__ is a static, CUB's radix sort is a stable sort) into a shared library that exports the same sbmf_fm_*
C ABI.  The GPU parity cases of tests/fm_gpu_cases.py then run against it unchanged (SBMF_FM_LIB_PATH): the transposed design
matrix bit-exact, the conflict-free runs, zero-noise parity with the pinned libFM restatement at 1e-4 after 10 iterations (MF and
general design matrices, -method als, K = 0 / 3 / 4 / 20, no w0 / no w), error behaviour, live chains as distributions.
Two builds: tier limits shrunk (warp <= 8 entries, CTA <= 40, 16-entry slices, 64-thread CTAs) so that the small fixtures reach
the warp, CTA and sliced column kernels; and the production geometry (256-thread CTAs, 512 / 16384 / 8192) on the long-column
matrix.  This checks indexing, work lists, launch geometry, reduction trees, barrier placement and the host orchestration --
everything but the hardware.  Test infrastructure only: the product never loads this library."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200")
OUT = os.path.join(ROOT, "tools", "build")
SHRUNK = ["-DFM_BLOCK_T=64", "-DFM_WARP_COL_MAX=8", "-DFM_BLOCK_COL_MAX=40", "-DFM_SLICE_LEN=16", "-DFM_HYPER_CHUNK=16"]


def build(name, defs):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, name)
    subprocess.run([gxx, "-O1", "-std=c++17", "-pthread", "-fPIC", "-shared", "-x", "c++", "-DSBMF_SIMT_EMU", *defs, "-I", os.path.join(ROOT, "tools", "emu_include"),
                    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"), "-o", so, os.path.join(PKG, "csrc", "fm.cu")], check=True, capture_output=True)
    return so


@pytest.fixture(scope="module")
def emu_shrunk():
    return build("libsbmf_fm_emu.so", SHRUNK)


@pytest.fixture(scope="module")
def emu_prod():
    return build("libsbmf_fm_emu_prod.so", [])


def run_case(lib, case, **env):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fm_gpu_cases.py"), case], capture_output=True, text=True, timeout=1500,
                       env=dict(os.environ, SBMF_FM_LIB_PATH=lib, **env))
    assert r.returncode == 0 and f"ok {case}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("case", ["columns", "zero_mf", "zero_general", "zero_als", "zero_variants", "zero_general_k20", "errors"])
def test_kernels_on_cpu_threads_all_tiers(emu_shrunk, case):
    run_case(emu_shrunk, case)


def test_front_ends_on_cpu_threads(emu_shrunk):
    """the programs that LINK the library follow onto the CPU execution through LD_PRELOAD: the host CLI (-method fm_mcmc: "#Iter="
    lines, test_rmse_* file, -out) and libFM's own main() with the CUDA learner spliced in (oracle/_ref/libFM_cuda) print the
    binding's chain"""
    subprocess.run(["make", "-C", PKG], check=True, capture_output=True)
    run_case(emu_shrunk, "cli")
    run_case(emu_shrunk, "libfm_learner")


def test_kernels_on_cpu_threads_live_sampling(emu_shrunk):
    run_case(emu_shrunk, "live_small", FM_LIVE_SEEDS="8")


def test_kernels_on_cpu_threads_production_geometry(emu_prod):
    run_case(emu_prod, "long_columns")

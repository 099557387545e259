"""CPU: the kernels of the SBMF sweep (csrc/kernels.cu, storage.cu behind the C ABI of api.cu) executed without a GPU.

tools/build_emu.sh compiles the library's own sources a second time with g++ against tools/emu_include (host stand-in for the CUDA
runtime and the SIMT execution model: the threads of a CTA are host threads, barriers and shuffles are real synchronisation, see
tests/test_fm_simt_emulation.py) and a selection of the GPU parity tests runs against that build unchanged (SBMF_LIB_PATH +
SBMF_EMULATED): device-built CSR/CSC/permutation bit-exact, zero-noise parity with the pinned restatement on the small fixtures,
edge cases, heavy (sliced) rows.  The kernels are GPU-validated already; what this adds is a second, independent
execution of the same code -- and the build that tools/sbmf_sanitize.sh instruments with AddressSanitizer / ThreadSanitizer, the
memory- and race-check that compute-sanitizer would do on a pool that allows it.  Test infrastructure only."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# fast members of tests/test_parity_gpu.py (~1 minute on host threads; all 8 heavy-row variants take 5, the ML-100K sweeps longer)
SELECT = "layout_bit_exact or ragged_random or edge_ or error_behaviour or (zero_noise_heavy_rows and (8-False-1 or 8-True-0))"


@pytest.fixture(scope="module")
def emu_lib():
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "build_emu.sh")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return os.path.join(ROOT, "tools", "build", "libsbmf_cuda_emu.so")


def test_sweep_kernels_on_cpu_threads(emu_lib):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_parity_gpu.py"), "-x", "-q", "-k", SELECT],
                       capture_output=True, text=True, timeout=3000, env=dict(os.environ, SBMF_EMULATED="1", SBMF_LIB_PATH=emu_lib))
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_smoke_job_on_cpu_threads(emu_lib):
    """__graft_entry__.smoke() -- ML-100K, K = 20: layout bit-exact, 3 zero-noise sweeps within 1e-4 of the oracle, 5 live sweeps on the
    reference's RMSE trajectory -- with every kernel of the job running on host threads (all resident-row bins of a real data set)."""
    code = "import sys; sys.path.insert(0, %r); import __graft_entry__ as g; g.smoke()" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1500,
                       env=dict(os.environ, SBMF_EMULATED="1", SBMF_LIB_PATH=emu_lib))
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_bindings_refuse_the_emulation_build_unasked(emu_lib):
    """The host build is reachable only when the harness names it AND says so (SBMF_EMULATED=1): the binding and bench.py refuse it."""
    env = {k: v for k, v in os.environ.items() if k != "SBMF_EMULATED"}
    code = ("import sys; sys.path.insert(0, %r); import sbmf\n"
            "try:\n    sbmf.load_library()\nexcept OSError as e:\n    print('refused:', e)\n" % os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(env, SBMF_LIB_PATH=emu_lib))
    assert "refused:" in r.stdout and "no CPU path" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       env=dict(os.environ, SBMF_EMULATED="1", SBMF_LIB_PATH=emu_lib))
    assert r.returncode != 0 and "test infrastructure" in r.stderr and "{" not in r.stdout, r.stdout + r.stderr


def test_host_cli_on_cpu_threads(emu_lib):
    """The product's host program (bin/sbmf, linked against libsbmf_cuda.so) with the emulation build preloaded over it: -dump_xt writes
    the device-built layout byte-identical to the reference's transpose tool, for libFM text, libFM binary and triple input."""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_xt_format.py"), "-x", "-q", "-m", "gpu", "-k", "cli_dump_xt"],
                       capture_output=True, text=True, timeout=1500, env=dict(os.environ, SBMF_EMULATED="1", SBMF_LIB_PATH=emu_lib, LD_PRELOAD=emu_lib))
    assert r.returncode == 0 and "3 passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_planner_and_second_fixture_on_cpu_threads(emu_lib):
    """Device-side exchange planner == host planner for every rank at world 2 / 3 / 8 (csrc/storage.cu), and the layout of the second
    real fixture (975 empty item rows) bit-exact -- the GPU cases, on host threads."""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_sharding.py"),
                        os.path.join(ROOT, "tests", "test_y_parity_second_fixture.py"), "-x", "-q", "-m", "gpu", "-k", "device_planner or layout_bit_exact_second"],
                       capture_output=True, text=True, timeout=1500, env=dict(os.environ, SBMF_EMULATED="1", SBMF_LIB_PATH=emu_lib))
    assert r.returncode == 0 and " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]

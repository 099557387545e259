"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/sbmf_cuda.h declares, and refuses to
run without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200")
HEADER = os.path.join(ROOT, "include", "sbmf_cuda.h")


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-C", PKG], check=True, capture_output=True)
    return os.path.join(PKG, "lib", "libsbmf_cuda.so")


FM_HEADER = os.path.join(ROOT, "include", "sbmf_fm_cuda.h")


def declared_symbols():
    src = open(HEADER).read() + open(FM_HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbmf_(?:cuda|fm)_\w+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("sbmf_cuda_create", "sbmf_cuda_set_train", "sbmf_cuda_set_test", "sbmf_cuda_get_layout", "sbmf_cuda_init_factors",
              "sbmf_cuda_sweep", "sbmf_cuda_eval", "sbmf_cuda_get_state", "sbmf_cuda_get_pred", "sbmf_cuda_get_timing",
              "sbmf_cuda_last_error", "sbmf_cuda_destroy",
              # general FM Gibbs: libFM's fm_learn interface (init / learn / predict)
              "sbmf_fm_create", "sbmf_fm_set_groups", "sbmf_fm_set_train", "sbmf_fm_set_test", "sbmf_fm_init", "sbmf_fm_learn", "sbmf_fm_predict",
              "sbmf_fm_rmse_history", "sbmf_fm_get_state", "sbmf_fm_get_columns", "sbmf_fm_get_runs", "sbmf_fm_last_error", "sbmf_fm_destroy"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.sbmf_cuda_abi_version() == 2


def test_header_compiles_as_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "sbmf_cuda.h"\n#include "sbmf_fm_cuda.h"\nint main(void){ sbmf_config c; sbmf_fm_config f; return (int)(sizeof(c) + sizeof(f)) == 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o", str(tmp_path / "t.o")], check=True)


def test_ctypes_struct_layout_matches_header(built, tmp_path):
    import sys
    sys.path.insert(0, PKG)
    import sbmf
    c = tmp_path / "sz.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sbmf_cuda.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(sbmf_config),'
                 ' sizeof(sbmf_priors), sizeof(sbmf_state), sizeof(sbmf_timing), sizeof(sbmf_synth_spec), offsetof(sbmf_config, priors), offsetof(sbmf_config, nccl_id),'
                 ' sizeof(sbmf_checkpoint_dims));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(sbmf.Config), ctypes.sizeof(sbmf.Priors), ctypes.sizeof(sbmf.State), ctypes.sizeof(sbmf.Timing),
            ctypes.sizeof(sbmf.SynthSpec), sbmf.Config.priors.offset, sbmf.Config.nccl_id.offset, ctypes.sizeof(sbmf.CheckpointDims)]
    assert got == want


def test_fm_ctypes_struct_layout_matches_header(built, tmp_path):
    import sys
    sys.path.insert(0, PKG)
    import sbmf
    c = tmp_path / "szfm.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sbmf_fm_cuda.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(sbmf_fm_config),'
                 ' sizeof(sbmf_fm_state), offsetof(sbmf_fm_config, seed), offsetof(sbmf_fm_config, regv), offsetof(sbmf_fm_state, w0));return 0;}\n')
    exe = tmp_path / "szfm"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(sbmf.FmConfig), ctypes.sizeof(sbmf.FmState), sbmf.FmConfig.seed.offset, sbmf.FmConfig.regv.offset, sbmf.FmState.w0.offset]


def test_no_cpu_fallback(built):
    """Without a visible GPU, create must FAIL with a message (never silently compute on the CPU)."""
    if os.path.exists("/dev/nvidia0"):
        pytest.skip("GPU present")
    import sys
    sys.path.insert(0, PKG)
    import sbmf
    with pytest.raises(sbmf.SbmfError) as e:
        sbmf.SbmfModel(K=8)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(sbmf.SbmfError) as e:
        sbmf.FmModel(10, 4)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may mention it."""
    bad = []
    for base in (PKG, os.path.join(ROOT, "include")):
        for dp, dn, fn in os.walk(base):
            if os.path.basename(dp) in ("build", "lib", "bin", "__pycache__"):
                continue
            for f in fn:
                if f.endswith((".cu", ".cuh", ".h", ".cpp", ".py", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"liboracle|sbmf_oracle|oracle_py|oracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_packed_gram_accumulation_equals_scalar(tmp_path):
    """csrc/gram.cuh: the FFMA2 (fma.rn.f32x2) form of the per-rating Gram accumulation visits every sum with the same fused
    multiply-adds in the same order as the scalar form -- checked bit for bit on the CPU (host emulation = two fmaf)."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "ffma2_check"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(PKG, "csrc"), "-o", str(exe),
                    os.path.join(ROOT, "tools", "ffma2_check.cu")], check=True, capture_output=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "agree bit for bit" in r.stdout, r.stdout + r.stderr

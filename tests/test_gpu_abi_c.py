"""The C ABI without ctypes: tests/abi_smoke.c is compiled with gcc (C99, -pedantic) against include/dmt.h and runs init_paths!, one
blocking sweep and one accept step on the device through dlopen/dlsym — what a Julia `ccall`, a cgo stub or any other FFI does."""
import os
import shutil
import subprocess

import pytest

import dmt_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe, "-ldl", "-lm"], check=True, capture_output=True)
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_header_compiles_as_c99(tmp_path):
    """(CPU) include/dmt.h is valid C, the program links without CUDA, and without a GPU the library refuses loudly"""
    import torch
    exe = _compile(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, dmt_b200._lib.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 3 and "dmt_create" in r.stderr, (r.returncode, r.stderr)     # no CPU fallback


@pytest.mark.gpu
@pytest.mark.own_lanes
@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_c_program_runs_a_sweep_on_the_device(tmp_path):
    exe = _compile(tmp_path)
    r = subprocess.run([exe, dmt_b200._lib.LIB_PATH], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "abi_smoke:" in r.stdout and "bad = 0" in r.stdout

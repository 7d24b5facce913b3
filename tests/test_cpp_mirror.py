"""The header-only C++ mirror of the reference classes (include/gms.hpp): compiled with g++ and run
against the oracle library on CPU boxes and against libgms.so (CUDA) on the GPU box."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(lib_path, tmp_path):
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    assert gxx, "g++ not found"
    exe = str(tmp_path / "mirror_test")
    libdir, libname = os.path.dirname(lib_path), os.path.basename(lib_path)
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"), "-o", exe, f"-L{libdir}",
                           f"-l:{libname}", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mirror_test OK" in out.stdout


def test_cpp_mirror_over_oracle(oracle, tmp_path):
    _build_and_run(oracle.path, tmp_path)


@pytest.mark.gpu
def test_cpp_mirror_over_cuda(cuda, tmp_path):
    _build_and_run(cuda.path, tmp_path)

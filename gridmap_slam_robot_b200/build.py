"""Builds the native code in-tree: csrc/libgms.so (CUDA, sm_100a) and oracle/libgms_ref.so (CPU oracle).

nvcc cross-compiles without a GPU.  --fmad=false: Java never contracts a*b+c and the bit-exact
parity claims depend on it.  -lineinfo keeps the ncu source page usable.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-shared", "-cudart", "static",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_libgms(force=False, verbose=False):
    out = os.path.join(CSRC, "libgms.so")
    srcs = [os.path.join(CSRC, f) for f in ("gms.cu", "kernels.cuh", "device_math.cuh")]
    srcs.append(os.path.join(ROOT, "include", "gms.h"))
    if not force and not _newer(out, srcs):
        return out
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, os.path.join(CSRC, "gms.cu")]
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd, env=env)
    return out


def build_e2e_host(force=False):
    """The compiled host loop over the C-ABI that bench.py times as `e2e_native` (plain g++: it only dlopens the library)."""
    out = os.path.join(CSRC, "e2e_host")
    src = os.path.join(CSRC, "e2e_host.cpp")
    if force or _newer(out, [src, os.path.join(ROOT, "include", "gms.h")]):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
        subprocess.check_call([gxx, "-std=c++17", "-O2", "-o", out, src, "-ldl"])
    return out


def build_oracle(force=False):
    odir = os.path.join(ROOT, "oracle")
    out = os.path.join(odir, "libgms_ref.so")
    srcs = [os.path.join(odir, "gms_ref.c"), os.path.join(ROOT, "include", "gms.h")]
    if force or _newer(out, srcs):
        subprocess.check_call(["make", "-C", odir] + (["-B"] if force else []))
    return out


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_libgms(force=force, verbose="-v" in sys.argv))
    print(build_e2e_host(force=force))
    print(build_oracle(force=force))

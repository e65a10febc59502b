"""Build libp3d_b200.so in-tree with nvcc for sm_100a (B200).

Usage: python build_native.py [--force]   (also called by __graft_entry__.build()).
The library travels to the GPU box with the repo snapshot; nothing is JIT-compiled at run time.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libp3d_b200.so")
SOURCES = ["p3d_core.cu", "p3d_camera.cu", "p3d_carve.cu", "p3d_deform.cu", "p3d_mesh.cu"]
HEADERS = ["p3d_common.cuh", "p3d_project.cuh", "p3d_mc_table.inc", os.path.join("..", "..", "include", "p3d_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",                       # bit-exact paths spell every fma() explicitly
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-cudart", "static",
    "-shared",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile the library.  `defines` / `out` build an experimental variant next to the default one
    (selected at run time with P3D_LIB=<path>; used only for tuning runs)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    target = out or LIB
    if not force and out is None and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + srcs
    subprocess.check_call(cmd, cwd=CSRC)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

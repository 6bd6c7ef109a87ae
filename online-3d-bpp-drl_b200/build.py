"""Builds libbpp_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python online-3d-bpp-drl_b200/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "b200bpp", "libbpp_b200.so")
SOURCES = ["bpp_api.cu", "bpp_step.cu", "bpp_mask.cu"]
HEADERS = ["bpp_common.cuh", "bpp_launch.h", os.path.join("..", "..", "include", "bpp_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SOURCES
    env = dict(os.environ)
    env.pop("CC", None)      # the image exports a gcc wrapper in $CC that nvcc must not pick up
    env.pop("CXX", None)
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

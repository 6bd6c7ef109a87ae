"""Builds libbpp_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python online-3d-bpp-drl_b200/build.py [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "b200bpp", "libbpp_b200.so")
SOURCES = ["bpp_api.cu", "bpp_step.cu", "bpp_mask.cu", "bpp_gen.cu"]
HEADERS = ["bpp_common.cuh", "bpp_launch.h", "bpp_params.h", os.path.join("..", "..", "include", "bpp_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]


# Named kernel build variants (switches in csrc/bpp_common.cuh): tools/ab_bench.py times them against the default build on a B200;
# tests/test_kernel_emul.py checks them bit for bit on the CPU.  Built next to the default as b200bpp/variants/*.so.
VARIANTS = {
    # round-1 kernel switches all off (the timing baseline of round 1) and the ordinary LDG/STG path of this round
    "baseline": ["BPP_PACKED_SWEEP=0", "BPP_LANE_CHECK=0", "BPP_SQUARE_COPY=0", "BPP_PLAIN_STORES=0", "BPP_TMA=0", "BPP_UNIFORM=0"],
    "notma": ["BPP_TMA=0"],
    "nouni": ["BPP_UNIFORM=0"],                   # ordinary instantiations without the uniform-datapath trick
    # bulk-copy path: bins per warp batch x resident CTAs per SM
    "n16": ["BPP_TMA_BPW_LOG2=4"],
}


def variant_path(name):
    return os.path.join(HERE, "b200bpp", "variants", "libbpp_b200_%s.so" % name)


def build_variants(names=None, verbose=False, jobs=4):
    """nvcc is single-threaded per translation unit: a few variants at a time (~30 s each, ~3 min for all)."""
    from concurrent.futures import ThreadPoolExecutor
    names = list(names or sorted(VARIANTS))
    with ThreadPoolExecutor(max_workers=jobs) as pool:
        return list(pool.map(lambda n: build(out=variant_path(n), defines=VARIANTS[n], verbose=verbose), names))


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest(defines=()):
    """Content hash of sources, headers, flags and this script (mtimes lie after a `mv` or a `git checkout`)."""
    import hashlib
    h = hashlib.sha1((" ".join(NVCC_FLAGS) + "|" + " ".join(defines)).encode())
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        h.update(open(os.path.join(CSRC, f), "rb").read())
    return h.hexdigest()


def up_to_date(out=None, defines=()):
    out = out or OUT
    try:
        return os.path.exists(out) and open(out + ".digest").read().strip() == _digest(defines)
    except OSError:
        return False


def build(force=False, verbose=False, out=None, defines=()):
    """out/defines build an experimental variant (e.g. defines=["BPP_MIN_BLOCKS=6"]) next to the default."""
    out = out or OUT
    if not force and up_to_date(out, defines):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    env = dict(os.environ)
    env.pop("CC", None)      # the image exports a gcc wrapper in $CC that nvcc must not pick up
    env.pop("CXX", None)
    subprocess.check_call(cmd, cwd=CSRC, env=env)
    with open(out + ".digest", "w") as f:
        f.write(_digest(defines) + "\n")
    return out


if __name__ == "__main__":
    if "--variants" in sys.argv:
        print("\n".join(build_variants(verbose="-v" in sys.argv)))
        sys.exit(0)
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=os.path.abspath(outs[0]) if outs else None,
                defines=defs))

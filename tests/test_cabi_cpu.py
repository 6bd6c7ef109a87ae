"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/bpp_b200.h declares, argument validation works, and nothing silently falls back to a CPU path."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _build():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bpp_build", os.path.join(ROOT, "online-3d-bpp-drl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


@pytest.fixture(scope="module")
def cabi():
    _build()
    from b200bpp import _cabi
    _cabi.lib()
    return _cabi


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "bpp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(bpp_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_exported_and_bound(cabi):
    names = declared_symbols()
    assert len(names) >= 15
    raw = C.CDLL(cabi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libbpp_b200.so does not export %s" % n
        assert n in cabi.SYMBOLS, "ctypes binding lacks %s" % n
    assert set(cabi.SYMBOLS) == set(names)
    assert raw.bpp_abi_version() == cabi.ABI_VERSION


def test_config_struct_layout_matches_header(cabi):
    # sizeof(bpp_config) as the C compiler sees it
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "bpp_b200.h"\nint main(){printf("%zu", sizeof(bpp_config));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        size = int(subprocess.check_output([os.path.join(d, "t")]))
    assert size == C.sizeof(cabi.Config)


def test_no_cpu_fallback(cabi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU behaviour")
    cfg = cabi.Config()
    cfg.abi_version = cabi.ABI_VERSION
    cfg.n_bins, cfg.L, cfg.W, cfg.H, cfg.mask_variant = 4, 10, 10, 10, 2
    h = C.c_void_p()
    rc = cabi.lib().bpp_create(C.byref(cfg), C.byref(h))
    assert rc == -3 and b"no CUDA device" in cabi.lib().bpp_last_error()
    import b200bpp
    with pytest.raises(RuntimeError):
        b200bpp.BinPackingVecEnv(4)


def test_argument_validation(cabi):
    l = cabi.lib()
    cfg = cabi.Config()
    h = C.c_void_p()
    assert l.bpp_create(C.byref(cfg), C.byref(h)) == -1          # ABI version 0
    cfg.abi_version = cabi.ABI_VERSION
    cfg.n_bins, cfg.L, cfg.W, cfg.H, cfg.mask_variant = 4, 40, 10, 10, 2
    assert l.bpp_create(C.byref(cfg), C.byref(h)) == -1 and b"[1, 32]" in l.bpp_last_error()
    cfg.L, cfg.mask_variant = 10, 3
    assert l.bpp_create(C.byref(cfg), C.byref(h)) == -1
    assert l.bpp_mask_from_obs(None, 1, 10, 10, 300, 0, 2, None, None) == -1
    assert l.bpp_destroy(None) == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "online-3d-bpp-drl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "bpp_oracle" not in text and "from oracle" not in text, f


def test_library_has_no_cpu_interpretation_inside(cabi):
    """tests/simt/ compiles the kernel sources for the CPU *for tests*; nothing of it may end up in the product library."""
    import subprocess
    syms = subprocess.check_output(["nm", "-D", "--defined-only", cabi.LIB_PATH]).decode()
    assert "simt" not in syms and "emul" not in syms
    exported = sorted(ln.split()[-1] for ln in syms.splitlines() if " T " in ln and "bpp_" in ln.split()[-1] and "_Z" not in ln)
    assert exported == declared_symbols()          # the C ABI and nothing else with a bpp_ name


def test_lazy_infos_and_spaces():
    import numpy as np
    from b200bpp.vec_env import LazyInfos, default_item_set, _pack_trajectories
    from b200bpp.spaces import Box, Discrete
    # built from the step's 16-byte records: reward bits | placed volume | counter | next action
    rec = np.array([[0, 200, 3, 17], [0, 500, 7, 4]], np.int32)
    infos = LazyInfos(np.array([0, 1], bool), rec, 1000.0, 0.0, 100)
    assert len(infos) == 2 and "episode" not in infos[0] and infos[0]["counter"] == 3 and infos[0]["ratio"] == 0.2
    assert infos[1]["episode"]["r"] == 5.0 and infos[1]["episode"]["l"] == 8 and infos[1]["mask"].shape == (100,)
    # Monitor rounds the float64 return (baselines/bench/monitor.py:64): exact from the integer volume, for any bin size
    odd = LazyInfos(np.array([1], bool), np.array([[0, 613, 9, 0]], np.int32), 8 * 12 * 9.0, 0.0, 192)
    assert odd[0]["episode"]["r"] == round(10.0 * 613 / 864, 6) and odd[0]["ratio"] == 613 / 864
    assert infos.done_indices().tolist() == [1]
    assert len(default_item_set()) == 64 and default_item_set()[1] == (2, 2, 3)
    traj, n = _pack_trajectories([[[2, 3, 4]], [[1, 1, 1], [5, 5, 5]]])
    assert traj.shape == (2, 2, 3) and n.tolist() == [1, 2]
    assert Discrete(200).__class__.__name__ == "Discrete" and Box(0, 10, (400,)).shape == (400,)

"""Puts the unmodified Python reference on sys.path (TEST INFRASTRUCTURE ONLY).

The reference tree exists only in the build container (/root/reference).  Elsewhere the byte-for-byte copy of its
`envs/`, `acktr/`, `baselines/` packages that oracle/make_ref.py vendors into oracle/_ref/ (git-ignored, shipped with
the gpurun snapshot) is used; with neither this returns None and callers fall back to the fixtures under tests/golden/.
"""
import contextlib
import io
import json
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_path():
    base = os.path.join(_HERE, "..", "BASELINE.json")
    try:
        with open(base) as f:
            p = json.load(f).get("reference_path", "/root/reference")
    except OSError:
        p = "/root/reference"
    if os.path.isdir(os.path.join(p, "envs", "bpp0")):
        return p
    vendored = os.path.join(_HERE, "_ref")      # oracle/make_ref.py: byte-for-byte copy that travels to the GPU box
    return vendored if os.path.isdir(os.path.join(vendored, "envs", "bpp0")) else None


def load_reference():
    """Returns a namespace with the reference modules, or None when the tree is absent."""
    p = reference_path()
    if p is None:
        return None
    shim = os.path.join(_HERE, "ref_shim")
    for q in (p, shim):
        if q not in sys.path:
            sys.path.insert(0, q)
    with contextlib.redirect_stdout(io.StringIO()):
        import envs.bpp0 as bpp0                      # noqa: E402
        import envs.bpp0.space as space               # noqa: E402
        import envs.bpp0.binCreator as binCreator     # noqa: E402
        import envs.bpp0.cutCreator as cutCreator     # noqa: E402
        import envs.bpp0.mdCreator as mdCreator       # noqa: E402
        import acktr.utils as utils                   # noqa: E402

    class _NS:
        pass
    ns = _NS()
    ns.PackingGame = bpp0.PackingGame
    ns.Space = space.Space
    ns.BoxCreator = binCreator.BoxCreator
    ns.RandomBoxCreator = binCreator.RandomBoxCreator
    ns.LoadBoxCreator = binCreator.LoadBoxCreator
    ns.CuttingBoxCreator = cutCreator.CuttingBoxCreator
    ns.MDlayerBoxCreator = mdCreator.MDlayerBoxCreator
    ns.utils = utils
    ns.path = p
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints on construction (bin3D.py:22,28; binCreator.py:36,46)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield

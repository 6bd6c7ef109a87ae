#!/usr/bin/env python
"""Recipe: vendors the UNMODIFIED reference files of the hot path into oracle/_ref/ (TEST / BASELINE INFRASTRUCTURE).

    python oracle/make_ref.py            # run by __graft_entry__.build() where /root/reference exists

The reference is pure Python -- nothing compiles -- so "building oracle/_ref" is copying the package directories the
reference's own CPU path imports (`envs/`, `acktr/`, `baselines/`) byte for byte.  oracle/_ref/ is git-ignored (reference
sources never enter this repository's history) but not gpurun-ignored: it travels to the GPU box with the snapshot, where
`bench.py --impl reference` times the reference itself (`make_vec_envs` -> `ShmemVecEnv(context='fork')` + the host mask
loop of main.py:163-169) on that box's host cores.  A manifest with the SHA-256 of every copied file is written next to
the copy; `verify()` re-checks it against /root/reference when that tree is present, so "unmodified" is testable.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
PACKAGES = ("envs", "acktr", "baselines")          # + nothing else: no datasets, no checkpoints, no GUI


def source_root():
    try:
        with open(os.path.join(HERE, "..", "BASELINE.json")) as f:
            p = json.load(f).get("reference_path", "/root/reference")
    except OSError:
        p = "/root/reference"
    return p if os.path.isdir(os.path.join(p, "envs", "bpp0")) else None


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _walk(root):
    for pkg in PACKAGES:
        for d, _, files in os.walk(os.path.join(root, pkg)):
            if "__pycache__" in d:
                continue
            for fn in sorted(files):
                if fn.endswith(".py"):
                    full = os.path.join(d, fn)
                    yield os.path.relpath(full, root), full


def build(force=False):
    """Copies the packages; returns DEST, or None when neither the reference tree nor an earlier copy exists."""
    src = source_root()
    manifest_path = os.path.join(DEST, "MANIFEST.json")
    if src is None:
        return DEST if os.path.exists(manifest_path) else None
    manifest = {rel: _sha(full) for rel, full in _walk(src)}
    if not force and os.path.exists(manifest_path):
        try:
            if json.load(open(manifest_path))["files"] == manifest and all(
                    os.path.exists(os.path.join(DEST, rel)) for rel in manifest):
                return DEST
        except (ValueError, KeyError):
            pass
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    for rel, full in _walk(src):
        out = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(full, out)
    with open(manifest_path, "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return DEST


def verify():
    """True when every file under oracle/_ref equals the manifest (and the reference tree, where present)."""
    manifest = json.load(open(os.path.join(DEST, "MANIFEST.json")))["files"]
    src = source_root()
    for rel, digest in manifest.items():
        if _sha(os.path.join(DEST, rel)) != digest:
            return False
        if src is not None and _sha(os.path.join(src, rel)) != digest:
            return False
    return True


if __name__ == "__main__":
    d = build(force="--force" in sys.argv)
    print(d if d else "no reference tree and no earlier copy: nothing to do")

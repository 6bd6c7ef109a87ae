#!/usr/bin/env python
"""Is the in-tree libbpp_b200.so, instruction for instruction, the build that last passed the GPU suite?

    python tools/check_validated_sass.py            # compare with profiles/validated_sass.md5
    python tools/check_validated_sass.py --record   # after a green `pytest -m gpu` on this exact build

The round's GPU budget is small, so kernel sources change more often than they can be re-run on a B200; switches that
are off must not perturb the generated code.  The hash covers the SASS of every kernel (cuobjdump), not the host code.
"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "online-3d-bpp-drl_b200", "b200bpp", "libbpp_b200.so")
REC = os.path.join(ROOT, "profiles", "validated_sass.md5")


def sass_hash(lib=LIB):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    keep = [ln for ln in out.splitlines() if "Function :" in ln or re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln)]
    return hashlib.md5("\n".join(keep).encode()).hexdigest(), len(keep)


if __name__ == "__main__":
    h, n = sass_hash()
    if "--record" in sys.argv:
        open(REC, "w").write("%s  %d SASS lines  libbpp_b200.so\n" % (h, n))
        print("recorded", h)
        sys.exit(0)
    want = open(REC).read().split()[0] if os.path.exists(REC) else None
    print("in-tree build %s (%d SASS lines); validated %s -> %s" % (h, n, want, "same" if h == want else "DIFFERENT"))
    sys.exit(0 if h == want else 1)

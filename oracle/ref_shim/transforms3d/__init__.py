"""Stand-in for `transforms3d` (absent from the image); envs/bpp0/mdCreator.py:3 imports it but the
values it returns never reach item sizes (mdCreator.py:19,33)."""
from . import euler   # noqa: F401

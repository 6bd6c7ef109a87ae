"""b200bpp -- Blackwell-native batched 3D bin-packing environment (host side).

Python mirror of the reference's interfaces for ONE hot path (envs/bpp0 + acktr/utils.py:8-94 +
acktr/envs.py), over the C ABI of libbpp_b200.so.  See DESIGN.md / INTEGRATION.md at the repo root.
"""
from ._cabi import BppError, MASK_V1, MASK_V2   # noqa: F401
from .vec_env import BinPackingVecEnv, default_item_set   # noqa: F401
from .envs import make_vec_envs   # noqa: F401
from . import utils   # noqa: F401
from .bpp0 import PackingGame   # noqa: F401,E402
from . import rollout, dist   # noqa: F401,E402
from .policy import PackingPolicy, greedy_policy   # noqa: F401,E402
from . import lookahead   # noqa: F401,E402

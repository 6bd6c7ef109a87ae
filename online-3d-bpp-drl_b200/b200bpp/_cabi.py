"""ctypes binding of libbpp_b200.so (include/bpp_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is present the
import / first call raises.  The oracle under oracle/ is test infrastructure and is never used here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BPP_B200_LIB lets kernel experiments A/B a differently built library; the default is the in-tree build
LIB_PATH = os.environ.get("BPP_B200_LIB") or os.path.join(_HERE, "libbpp_b200.so")

ABI_VERSION = 2
GEN_RS, GEN_REPLAY, GEN_CUT1, GEN_CUT2 = 0, 1, 2, 3
GEN_KINDS = {"rs": GEN_RS, "replay": GEN_REPLAY, "cut1": GEN_CUT1, "cut2": GEN_CUT2}
MASK_V1, MASK_V2 = 1, 2


class BppError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbpp_b200 error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("n_bins", C.c_int32),
        ("L", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("rot", C.c_int32),
        ("mask_variant", C.c_int32), ("gen_kind", C.c_int32),
        ("n_items", C.c_int32), ("item_set_host", C.c_void_p),
        ("n_traj", C.c_int32), ("traj_cap", C.c_int32), ("traj_host", C.c_void_p),
        ("traj_len_host", C.c_void_p), ("first_traj", C.c_int32),
        ("lo", C.c_int32 * 3), ("hi", C.c_int32 * 3),
        ("seed", C.c_uint64), ("global_offset", C.c_int64), ("global_n", C.c_int64),
    ]


# name -> (restype, argtypes); every symbol include/bpp_b200.h declares
_vp, _i32 = C.c_void_p, C.c_int32
SYMBOLS = {
    "bpp_abi_version": (_i32, []),
    "bpp_last_error": (C.c_char_p, []),
    "bpp_launch_count": (C.c_uint64, []),
    "bpp_build_info": (C.c_char_p, []),
    "bpp_create": (_i32, [C.POINTER(Config), C.POINTER(_vp)]),
    "bpp_destroy": (_i32, [_vp]),
    "bpp_reset": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "bpp_step": (_i32, [_vp] * 10),
    "bpp_step_host": (_i32, [_vp] * 10),
    "bpp_step_packed": (_i32, [_vp] * 7 + [_i32, _vp]),
    "bpp_step_packed_host": (_i32, [_vp] * 8),
    "bpp_step_packed_host_i32": (_i32, [_vp, _vp, _i32] + [_vp] * 6),
    "bpp_host_alloc": (_i32, [_i32, C.c_uint64, C.POINTER(_vp)]),
    "bpp_host_free": (_i32, [_vp]),
    "bpp_bind_thread_to_device": (_i32, [_i32]),
    "bpp_mask_from_obs": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "bpp_mask_from_plain": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "bpp_check_box": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "bpp_export_state": (_i32, [_vp] * 6),
    "bpp_import_state": (_i32, [_vp] * 6),
    "bpp_observe": (_i32, [_vp, _vp, _vp, _vp]),
    "bpp_metrics_host": (_i32, [_vp, _vp, _vp]),
    "bpp_episode_items_host": (_i32, [_vp, C.c_int64, _vp, _i32, _vp, _vp]),
    "bpp_state_bytes": (C.c_int64, [_vp]),
    "bpp_save_state": (_i32, [_vp, _vp, _vp]),
    "bpp_load_state": (_i32, [_vp, _vp, _vp]),
    "bpp_debug_draws": (_i32, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, _i32, _vp, _vp]),
    "bpp_debug_curand_draws": (_i32, [C.c_uint64, C.c_uint64, _i32, _vp, _vp]),
}

_lib = None


def lib():
    """Loads the shared library (built by online-3d-bpp-drl_b200/build.py); raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libbpp_b200.so not found at %s -- run `python online-3d-bpp-drl_b200/build.py` "
                              "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        if l.bpp_abi_version() != ABI_VERSION:
            raise ImportError("libbpp_b200.so ABI %d != binding ABI %d" % (l.bpp_abi_version(), ABI_VERSION))
        _lib = l
    return _lib


def check(code):
    if code != 0:
        raise BppError(code, lib().bpp_last_error().decode("utf-8", "replace"))

"""Minimal stand-in for the `gym` package (absent from the image, no network).

TEST INFRASTRUCTURE: lets tests/golden/make_golden.py and the live differential tests import the
unmodified reference from /root/reference.  Only the names the reference touches exist.
"""
from . import spaces, core, envs          # noqa: F401
from .core import Env, Wrapper, ObservationWrapper, RewardWrapper, ActionWrapper  # noqa: F401
from .envs.registration import register, make   # noqa: F401

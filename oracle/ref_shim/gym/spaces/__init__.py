from .box import Box            # noqa: F401
from .discrete import Discrete  # noqa: F401


class Dict(dict):
    def __init__(self, spaces=None):
        super().__init__(spaces or {})
        self.spaces = dict(spaces or {})


class Tuple(tuple):
    def __new__(cls, spaces=()):
        obj = super().__new__(cls, spaces)
        obj.spaces = tuple(spaces)
        return obj


class MultiDiscrete(object):
    def __init__(self, nvec):
        self.nvec = nvec


class MultiBinary(object):
    def __init__(self, n):
        self.n = n
        self.shape = (n,)

"""The two gym spaces the reference's consumers look at (gym itself is not a dependency).

acktr/model.py:26-29 reads `action_space.__class__.__name__ == "Discrete"` and `.n`;
acktr/storage.py:19-25 reads `observation_space.shape`; envs/bpp0/bin3D.py:38-41 defines both.
"""
import numpy as np


class Box(object):
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high = float(low), float(high)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return "Box(%g, %g, %s, %s)" % (self.low, self.high, self.shape, self.dtype)


class Discrete(object):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def __repr__(self):
        return "Discrete(%d)" % self.n

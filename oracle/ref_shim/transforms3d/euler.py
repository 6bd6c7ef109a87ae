import numpy as np


def quat2euler(q, axes="sxyz"):
    return (0.0, 0.0, 0.0)


def quat2mat(q):
    return np.eye(3)

import numpy as np


def dot(x, y):
    return float(np.dot(np.ravel(x), np.ravel(y)))

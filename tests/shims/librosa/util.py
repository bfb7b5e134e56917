import numpy as np


def pad_center(data, *, size, axis=-1, **kw):
    n = data.shape[axis]
    lp = (size - n) // 2
    widths = [(0, 0)] * data.ndim
    widths[axis] = (lp, size - n - lp)
    return np.pad(data, widths)


def tiny(x):
    dt = np.asarray(x).dtype
    return np.finfo(dt if np.issubdtype(dt, np.floating) else np.float32).tiny


def normalize(x, **kw):
    return x / max(1e-12, np.abs(x).max())

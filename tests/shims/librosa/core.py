import numpy as np


def resample(y, *, orig_sr, target_sr, **kw):
    if orig_sr == target_sr:
        return y
    from scipy.signal import resample_poly
    from math import gcd
    g = gcd(int(orig_sr), int(target_sr))
    return resample_poly(y, int(target_sr) // g, int(orig_sr) // g).astype(np.asarray(y).dtype)

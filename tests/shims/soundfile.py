"""Stand-in for soundfile (see README.md): 16-bit PCM wav read / write through scipy.io.wavfile."""
import numpy as np
from scipy.io import wavfile


def read(path, **kw):
    sr, data = wavfile.read(path)
    if data.dtype == np.int16:
        data = data.astype(np.float64) / 32768.0
    elif data.dtype == np.int32:
        data = data.astype(np.float64) / 2147483648.0
    else:
        data = data.astype(np.float64)
    return data, sr


def write(path, data, samplerate, subtype=None, **kw):
    x = np.asarray(data, dtype=np.float64)
    if subtype in (None, "PCM_16"):
        # libsndfile's float -> PCM_16 conversion: scale by 0x8000, round to nearest, clip
        q = np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16)
        wavfile.write(path, int(samplerate), q)
    else:
        wavfile.write(path, int(samplerate), x.astype(np.float32))

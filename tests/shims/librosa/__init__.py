"""Stand-in for librosa (see ../README.md): only what dataset.py / Models/istftnet.py / the scripts touch."""
import numpy as np

from . import core, filters, util  # noqa: F401
from .core import resample  # noqa: F401


def load(path, sr=None, mono=True, **kw):
    """librosa.load for PCM wav files (dataset.py:15): float32 in [-1, 1), resampled only if asked to."""
    import soundfile as sf
    data, file_sr = sf.read(path)
    data = np.asarray(data, dtype=np.float32)
    if data.ndim > 1 and mono:
        data = data.mean(axis=1)
    if sr is not None and sr != file_sr:
        data = resample(data, orig_sr=file_sr, target_sr=sr)
        file_sr = sr
    return data, file_sr

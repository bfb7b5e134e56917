"""Slaney mel filterbank, the host-side constant the front-end needs.

The reference builds it with ``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)``
(dataset.py:73; librosa==0.10.2.post1, defaults htk=False, norm="slaney").  When librosa
is importable it is used, so a reference user gets bit-identical filters; otherwise the
same published construction is evaluated here in float64."""
from __future__ import annotations

import math

import numpy as np

_F_SP = 200.0 / 3.0          # Hz per mel below the knee
_KNEE_HZ = 1000.0
_KNEE_MEL = _KNEE_HZ / _F_SP
_LOGSTEP = math.log(6.4) / 27.0


def _mel_of_hz(hz):
    hz = float(hz)
    return _KNEE_MEL + math.log(hz / _KNEE_HZ) / _LOGSTEP if hz >= _KNEE_HZ else hz / _F_SP


def _hz_of_mel(mel):
    mel = np.asarray(mel, dtype=np.float64)
    return np.where(mel >= _KNEE_MEL, _KNEE_HZ * np.exp(_LOGSTEP * (mel - _KNEE_MEL)), _F_SP * mel)


def slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax):
    """float32 ``[n_mels, n_fft // 2 + 1]`` triangular, area-normalised mel filters."""
    try:  # exactly what the reference calls, when the dependency exists
        from librosa.filters import mel as _librosa_mel  # type: ignore
        return np.ascontiguousarray(_librosa_mel(sr=sr, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax), dtype=np.float32)
    except ImportError:
        pass
    if fmax is None:
        fmax = sr / 2.0
    bins = np.arange(n_fft // 2 + 1, dtype=np.float64) * (1.0 / (n_fft * (1.0 / sr)))  # == np.fft.rfftfreq
    edges = _hz_of_mel(np.linspace(_mel_of_hz(fmin), _mel_of_hz(fmax), n_mels + 2))
    width = edges[1:] - edges[:-1]
    basis = np.empty((n_mels, bins.shape[0]), dtype=np.float32)
    for m in range(n_mels):
        rising = (bins - edges[m]) / width[m]
        falling = (edges[m + 2] - bins) / width[m + 1]
        tri = np.clip(np.minimum(rising, falling), 0.0, None).astype(np.float32)
        # librosa stores the triangle in float32 and then applies the float64 Slaney factor
        basis[m] = (tri.astype(np.float64) * (2.0 / (edges[m + 2] - edges[m]))).astype(np.float32)
    return basis

"""``dataset`` as the reference scripts import it (``from dataset import Dataset, mel_spectrogram,
get_dataset_filelist`` -- train_time_wi_inv.py:17; ``from dataset import mel_spectrogram, load_wav`` --
infers/inference_hifigan.py:10; ``from dataset import inverse_mel`` -- Models/hddemucas.py:19), selected by
putting this directory on PYTHONPATH: the scripts stay byte-identical.

Everything the reference's own ``dataset.py`` defines (``Dataset``, ``load_wav``, ``get_dataset_filelist``, ...) is
re-exported unchanged from the reference checkout found on ``sys.path``; the three signal functions on the
accelerated path are replaced by the sm_100a kernels:

  mel_spectrogram   dataset.py:53-91     inverse_mel   dataset.py:94-121     amp_pha_specturm   dataset.py:124-139

``Dataset.__getitem__`` (dataset.py:218-241) keeps calling the REFERENCE's functions: it is the reference's class
and resolves those names in the reference module's globals.  That is the data-loading path, running in forked
DataLoader workers (``num_workers=4``) where a CUDA context cannot be created, with ``in_dataset=True`` asking for
a CPU result -- so an explicit ``in_dataset=True`` call made through this module is routed to the reference's own
function as well (its code on its data path; nothing of this package runs there).  Every other call runs on the
GPU or raises: there is no CPU implementation here."""
from __future__ import annotations

from _locate import load_package as _load_package, load_reference_module as _load_reference_module

_pkg = _load_package()
_fast = _pkg.dataset
_ref = _load_reference_module("_nvse_reference_dataset", "dataset.py")

if _ref is not None:  # the reference's data pipeline, unchanged
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})

# caches and helpers of the accelerated functions (same names as dataset.py:27-50)
mel_window = _fast.mel_window
inv_mel_window = _fast.inv_mel_window
param_string = _fast.param_string
dynamic_range_compression_torch = _fast.dynamic_range_compression_torch
spectral_normalize_torch = _fast.spectral_normalize_torch
spectral_de_normalize_torch = _fast.spectral_de_normalize_torch


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=True, in_dataset=False):
    """dataset.py:53-91.  ``in_dataset=True`` (the DataLoader-worker call) -> the reference's own CPU function."""
    if in_dataset and _ref is not None:
        return _ref.mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=center,
                                    in_dataset=True)
    return _fast.mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=center,
                                 in_dataset=in_dataset)


def inverse_mel(mel, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, in_dataset=False):
    """dataset.py:94-121."""
    if in_dataset and _ref is not None:
        return _ref.inverse_mel(mel, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, in_dataset=True)
    return _fast.inverse_mel(mel, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, in_dataset=in_dataset)


def amp_pha_specturm(y, n_fft, hop_size, win_size):
    """dataset.py:124-139."""
    return _fast.amp_pha_specturm(y, n_fft, hop_size, win_size)

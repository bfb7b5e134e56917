"""``Models`` as the reference scripts import it (``from Models import HiFiGAN`` -- infers/inference_hifigan.py:20;
``from Models import (HiFiGAN, iSTFTNet, HDDemucas, ConvTasNet)`` and ``from Models.models import ...`` --
train_time_wi_inv.py:18-32), selected by putting the parent directory on PYTHONPATH.

``HiFiGAN`` and ``iSTFTNet`` are the B200-backed modules (same constructor, state-dict keys,
``remove_weight_norm()`` and call signature as Models/hifigan.py:83-133 and Models/istftnet.py:271-328), so
``eval(h.model_name)(h)`` builds them.  Everything else -- the eight other model families and ``Models.models``
(discriminators, losses) -- is the reference's own code: its ``Models`` directory is appended to this package's
``__path__`` and the class names of Models/__init__.py:1-10 are resolved lazily, so importing this package does not
import all ten models (and matplotlib / librosa with them) the way the reference's ``__init__`` does."""
from __future__ import annotations

import importlib
import os

from _locate import find_reference as _find_reference, load_package as _load_package

_pkg = _load_package()
HiFiGAN = _pkg.HiFiGAN
iSTFTNet = _pkg.iSTFTNet

_ref_init = _find_reference(os.path.join("Models", "__init__.py"))
if _ref_init is not None:
    __path__.append(os.path.dirname(_ref_init))

_REFERENCE_CLASSES = {  # Models/__init__.py:1-8
    "APNet": "apnet", "APNet2": "apnet2", "BSRNN": "bsrnn", "BSRNN_24k": "bsrnn_24k", "ConvTasNet": "convtasnet",
    "FreeV": "freeV", "GCRN": "gcrn", "HDDemucas": "hddemucas",
}
__all__ = ["HiFiGAN", "iSTFTNet"] + sorted(_REFERENCE_CLASSES)


def __getattr__(name):
    sub = _REFERENCE_CLASSES.get(name)
    if sub is None:
        raise AttributeError(f"module 'Models' has no attribute {name!r}")
    if _ref_init is None:
        raise ImportError(f"Models.{name} is the reference's own class, but no reference checkout is on sys.path "
                          "(set NVSE_REFERENCE_ROOT or run the script from the reference tree)")
    value = getattr(importlib.import_module(f"{__name__}.{sub}"), name)
    globals()[name] = value
    return value

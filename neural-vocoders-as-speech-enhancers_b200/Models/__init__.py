"""Drop-in ``Models`` package for the two generators on the accelerated path.

``from Models import HiFiGAN, iSTFTNet`` (train_time_wi_inv.py:28-29,
infers/inference_hifigan.py:20) resolves here when this package's ``dropin`` directory
precedes the reference checkout on ``sys.path``; ``eval(h.model_name)(h)`` then builds the
B200-backed module.  ``Models.models`` holds the waveform discriminators (MultiPeriodDiscriminator,
MultiScaleDiscriminator) and the trainer's loss functions.  The other eight model families of the reference are out of scope."""
from .hifigan import HiFiGAN
from .istftnet import iSTFTNet
from . import models  # noqa: F401
from .models import MultiPeriodDiscriminator, MultiScaleDiscriminator  # noqa: F401

__all__ = ["HiFiGAN", "iSTFTNet", "MultiPeriodDiscriminator", "MultiScaleDiscriminator"]

"""B200-native (sm_100a) time-domain vocoding hot path of
Andong-Li-speech/Neural-Vocoders-as-Speech-Enhancers: ``dataset.mel_spectrogram`` and the
``Models.HiFiGAN`` / ``Models.iSTFTNet`` generator forwards, behind the reference's own
Python API, backed by hand-written CUDA kernels through a C ABI (include/nvse_b200.h).

The directory name contains hyphens, so import it with
``importlib.import_module("neural-vocoders-as-speech-enhancers_b200")`` (what
``__graft_entry__.load_package()`` does; it also registers the alias ``nvse_b200``)."""
from . import _lib  # noqa: F401
from .dataset import mel_spectrogram, inverse_mel, amp_pha_specturm, istft  # noqa: F401
from .Models import HiFiGAN, iSTFTNet  # noqa: F401
from .pipeline import Vocoder  # noqa: F401
from .shard import shard_range, shard_by_cost, bucket_by_length, allreduce_gradients  # noqa: F401
from . import melbasis, _engine  # noqa: F401

__all__ = ["mel_spectrogram", "inverse_mel", "amp_pha_specturm", "istft", "HiFiGAN", "iSTFTNet", "Vocoder", "shard_range", "shard_by_cost", "bucket_by_length",
           "allreduce_gradients"]

"""``Models.models`` as train_time_wi_inv.py:18-26 imports it.

``MultiPeriodDiscriminator`` / ``MultiScaleDiscriminator`` (with ``DiscriminatorP`` / ``DiscriminatorS``) are the B200-backed
modules -- same constructors, checkpoint keys and random initialisation as Models/models.py:15-113,187-246, every
convolution forward and backward in the sm_100a kernels of csrc/disc.cu.  Every other name of the reference's module (the
losses, ``MultiResolutionDiscriminator``, the CQT discriminator, ``MultiResolutionMelLoss`` ...) is the reference's own object,
taken from its ``Models/models.py`` when that file can be imported (it needs torchaudio; the CQT discriminator needs nnAudio
at call time only); the loss functions the time-domain trainer uses are also provided by the package itself, so the trainer
runs without the reference's module being importable at all."""
from _locate import load_package as _load_package, load_reference_module as _load_reference_module

_m = _load_package().Models.models

try:
    _ref = _load_reference_module("_nvse_reference_models", "Models/models.py")
except Exception:  # a missing optional dependency of the reference's module must not take the trainer down
    _ref = None
if _ref is not None:
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})

DiscriminatorP, MultiPeriodDiscriminator = _m.DiscriminatorP, _m.MultiPeriodDiscriminator
DiscriminatorS, MultiScaleDiscriminator = _m.DiscriminatorS, _m.MultiScaleDiscriminator
for _name in ("feature_loss", "hinge_generator_loss", "hinge_discriminator_loss", "ls_generator_loss", "ls_discriminator_loss",
              "LRELU_SLOPE", "get_padding"):
    globals().setdefault(_name, getattr(_m, _name))

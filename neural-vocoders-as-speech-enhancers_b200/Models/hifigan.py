"""``Models.HiFiGAN`` drop-in (reference Models/hifigan.py:83-133)."""
from ._blocks import GEN_HIFIGAN, ResBlock1, ResBlock2, _GeneratorBase  # noqa: F401

LRELU_SLOPE = 0.1  # hifigan.py:7 (the final activation uses F.leaky_relu's default 0.01, hifigan.py:120)


class HiFiGAN(_GeneratorBase):
    """HiFi-GAN generator.  ``HiFiGAN(h)(mel[B,80,F]) -> wav[B, prod(upsample_rates)*F]``;
    state-dict keys, ``remove_weight_norm()`` and seeded initialisation match the reference."""

    _kind = GEN_HIFIGAN

    def __init__(self, h):
        super().__init__(h, post_channels=1)

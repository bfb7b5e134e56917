"""``Models.iSTFTNet`` drop-in (reference Models/istftnet.py:271-328)."""
from ._blocks import GEN_ISTFTNET, ResBlock1, ResBlock2, _GeneratorBase  # noqa: F401

LRELU_SLOPE = 0.1


class iSTFTNet(_GeneratorBase):
    """iSTFTNet generator: two upsampling stages, then conv_post -> (exp, sin) -> inverse
    STFT (n_fft = h.gen_istft_n_fft, hop = h.gen_istft_hop_size) fused in one head kernel."""

    _kind = GEN_ISTFTNET

    def __init__(self, h):
        super().__init__(h, post_channels=h.gen_istft_n_fft + 2)
        self.post_n_fft = h.gen_istft_n_fft

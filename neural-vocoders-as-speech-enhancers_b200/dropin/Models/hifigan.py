"""``Models.hifigan`` (reference Models/hifigan.py:1-133): the generator and its blocks, B200-backed."""
from _locate import load_package as _load_package

_m = _load_package().Models.hifigan
HiFiGAN, ResBlock1, ResBlock2, LRELU_SLOPE = _m.HiFiGAN, _m.ResBlock1, _m.ResBlock2, _m.LRELU_SLOPE

"""``Models.istftnet`` (reference Models/istftnet.py:196-328): the generator and its blocks, B200-backed."""
from _locate import load_package as _load_package

_m = _load_package().Models.istftnet
iSTFTNet, ResBlock1, ResBlock2, LRELU_SLOPE = _m.iSTFTNet, _m.ResBlock1, _m.ResBlock2, _m.LRELU_SLOPE

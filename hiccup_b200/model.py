"""Enums and the three-channel holder that cross the drop-in boundary.

Mirrors reference hiccup/model.py:10-74 (same names, members and attribute names), so objects are
interchangeable with the reference's by duck typing.  `CompressedImage` in DCT mode holds three 2-D
coefficient planes (int32 from compression, float64 from decode); in wavelet mode three lists of
ten 2-D int32 sub-bands.
"""
import enum

import numpy as np


class Compression(enum.Enum):
    JPEG = "JPEG"
    HIC = "HIC"


class Coefficient(enum.Enum):
    DC = "DC"
    AC = "AC"


class QTables(enum.Enum):
    JPEG_LUMINANCE = "jpeg standard luminance"
    JPEG_CHROMINANCE = "jpeg standard chrominance"


class Wavelet(enum.Enum):
    DAUBECHIE = "db1"
    HAAR = "haar"
    COIF = "coif1"
    SYM = "sym2"


class CompressedImage:
    CHANNELS = ("lum", "cr", "cb")

    def __init__(self, lum, cr, cb):
        self.luminance_component = lum
        self.red_chrominance_component = cr
        self.blue_chrominance_component = cb

    @classmethod
    def from_dict(cls, d):
        assert len(d) == 3
        return cls(d["lum"], d["cr"], d["cb"])

    @property
    def as_dict(self):
        return dict(zip(self.CHANNELS, (self.luminance_component, self.red_chrominance_component,
                                        self.blue_chrominance_component)))

    @property
    def shape(self):
        return self.luminance_component.shape, self.red_chrominance_component.shape

    def __eq__(self, other):
        if not all(hasattr(other, a) for a in ("luminance_component", "red_chrominance_component",
                                               "blue_chrominance_component")):
            return False
        mine, theirs = self.as_dict, CompressedImage.as_dict.fget(other)
        return all(np.array_equiv(mine[c], theirs[c]) for c in self.CHANNELS)

    __hash__ = None

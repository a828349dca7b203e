"""The `.hic` container: a pickled list of byte strings, one per payload.

Wire layout (reference hiccup/hicimage.py:30-183), preserved byte for byte:

    [ b"JPEG" | b"HIC",
      <tables>   each = pickle.dumps({"type": TupP, "data": [pickle.dumps((symbol, code)), ...]}),
      <bit data> each = iohelper.padded_bs_2_bytes(bits),
      pickle.dumps((h, w)), pickle.dumps((h, w)) ]

DCT mode: 9 tables, 9 bit strings, 2 shapes (21 entries); wavelet mode: 6, 6, 2 (15 entries).

Unlike the reference, a bit payload may be held as its framed bytes (what the GPU bit-packer
produces and what `from_bytes` receives) and is only expanded to a '0'/'1' string when `.payload`
is read, so a multi-megabit stream never passes through Python string code on the fast path.
"""
import pickle

from hiccup_b200 import _compat, iohelper, model


class Payload:
    @classmethod
    def from_bytes(cls, b):
        raise NotImplementedError

    @property
    def byte_stream(self):
        raise NotImplementedError


class TupP(Payload):
    """A pair: an image/sub-band shape, or one Huffman table row (symbol, code string)."""

    def __init__(self, n1, n2):
        self.n1, self.n2 = n1, n2

    @classmethod
    def from_bytes(cls, b):
        a, c = pickle.loads(b)
        return cls(a, c)

    @property
    def numbers(self):
        return self.n1, self.n2

    @property
    def byte_stream(self):
        return pickle.dumps((self.n1, self.n2))

    def __eq__(self, other):
        return hasattr(other, "numbers") and tuple(other.numbers) == self.numbers

    __hash__ = None


class BitStringP(Payload):
    """Huffman-coded data.  Construct from a '0'/'1' string (reference signature) or, via
    `from_bytes` / `from_framed`, from the framed bytes."""

    def __init__(self, string=None, framed=None):
        assert (string is None) != (framed is None)
        self._bits, self._framed = string, (None if framed is None else bytes(framed))

    @classmethod
    def from_bytes(cls, b):
        return cls(framed=b)

    from_framed = from_bytes

    @property
    def payload(self) -> str:
        if self._bits is None:
            self._bits = iohelper.padded_bytes_2_bs(self._framed)
        return self._bits

    @property
    def bit_count(self) -> int:
        return len(self._bits) if self._framed is None else iohelper.payload_bit_count(self._framed)

    @property
    def byte_stream(self):
        if self._framed is None:
            self._framed = iohelper.padded_bs_2_bytes(self._bits)
        return self._framed

    def __eq__(self, other):
        if not hasattr(other, "byte_stream") or not hasattr(other, "payload"):
            return False
        return bytes(other.byte_stream) == self.byte_stream

    __hash__ = None


class PlainStringP(Payload):
    ENCODING = "ascii"

    def __init__(self, string):
        self.payload = string

    @classmethod
    def from_bytes(cls, b):
        return cls(bytes(b).decode(cls.ENCODING))

    @property
    def byte_stream(self):
        return self.payload.encode(self.ENCODING)

    def __eq__(self, other):
        return getattr(other, "payload", None) == self.payload

    __hash__ = None


class PayloadStringP(Payload):
    """A run of payloads of one type -- in practice the rows of one Huffman table."""

    def __init__(self, t, payloads):
        self.t, self.payloads = t, payloads

    @classmethod
    def from_bytes(cls, b):
        d = pickle.loads(b)
        # rows are plain pickled pairs; parse them directly rather than through d["type"] so that
        # files written by the reference and by this package read the same way
        return cls(d["type"], [TupP.from_bytes(x) for x in d["data"]])

    @classmethod
    def from_rows(cls, rows):
        """rows: iterable of (symbol, code string)."""
        return cls(_compat.wire_tuple_class(), [TupP(s, c) for s, c in rows])

    @property
    def rows(self):
        return [p.numbers for p in self.payloads]

    @property
    def byte_stream(self):
        return pickle.dumps({"type": _compat.wire_tuple_class(),
                             "data": [p.byte_stream for p in self.payloads]})

    def __eq__(self, other):
        return hasattr(other, "payloads") and list(other.payloads) == list(self.payloads)

    __hash__ = None


class HicImage:
    LAYOUT = {model.Compression.JPEG: (9, 9, 2), model.Compression.HIC: (6, 6, 2)}

    def __init__(self, hic_type, settings, payloads):
        self.hic_type, self.settings, self._payloads = hic_type, settings, payloads

    @classmethod
    def jpeg_image(cls, payloads):
        return cls(model.Compression.JPEG, [PlainStringP(model.Compression.JPEG.value)], payloads)

    @classmethod
    def wavelet_image(cls, payloads):
        return cls(model.Compression.HIC, [PlainStringP(model.Compression.HIC.value)], payloads)

    @classmethod
    def from_bytes(cls, raw_data):
        kind = model.Compression(PlainStringP.from_bytes(raw_data[0]).payload)
        n_tab, n_bits, n_shape = cls.LAYOUT[kind]
        a, b, c = 1, 1 + n_tab, 1 + n_tab + n_bits
        payloads = ([PayloadStringP.from_bytes(x) for x in raw_data[a:b]]
                    + [BitStringP.from_bytes(x) for x in raw_data[b:c]]
                    + [TupP.from_bytes(x) for x in raw_data[c:c + n_shape]])
        return cls.jpeg_image(payloads) if kind == model.Compression.JPEG else cls.wavelet_image(payloads)

    @classmethod
    def from_file(cls, path):
        with open(path, "rb") as f:
            raw = pickle.load(f)
        assert raw is not None
        return cls.from_bytes(raw)

    @property
    def payloads(self):
        return self._payloads

    def byte_stream(self):
        return [p.byte_stream for p in self.settings + self._payloads]

    def write_file(self, path):
        with open(path, "wb") as f:
            pickle.dump(self.byte_stream(), f)

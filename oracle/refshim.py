"""Make the UNMODIFIED reference importable in this container.  TEST INFRASTRUCTURE ONLY.

The reference (`/root/reference/hiccup`) imports three packages that are not installed here
(`rawpy`, `bitstring`, `pywt` -- reference hiccup/iohelper.py:1-2, hiccup/transform.py:5).  This
module installs small stand-ins for them into `sys.modules` and puts `/root/reference` on
`sys.path`, so that `import hiccup.compression` etc. run the reference's own code unchanged.

Used only by `tests/golden/gen_golden.py` (to produce the committed golden vectors) and by the
`-m "not gpu"` pinning tests, which skip when `/root/reference` is absent (it does not exist on the
GPU box).  Nothing under `hiccup_b200/` imports this file.

Stand-ins:
  * rawpy      -- empty module (only `iohelper.open_raw_img` touches it; never called).
  * bitstring  -- `BitArray` / `Bits` with exactly the operations iohelper.py:32-56 uses, following
                  bitstring's published semantics (an int initialiser means "that many zero bits";
                  `.int` is the two's-complement value; `<<=` keeps the length).
                  Pinned by the reference's own iohelpertest.py KAT ("101" -> b"\\x05\\xa0").
  * pywt       -- `oracle.pywt_standin` (db1/haar only; PARITY UNPINNED, see that file).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HICCUP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hiccup", "codec.py"))


class _Bits:
    """Bit container over a Python str of '0'/'1' (enough for iohelper.py)."""

    def __init__(self, auto=None, bytes=None, uint=None, length=None):
        if bytes is not None:
            data = builtins_bytes(bytes)
            self._s = bin(int.from_bytes(data, "big"))[2:].zfill(8 * len(data)) if len(data) else ""
        elif uint is not None:
            self._s = bin(uint)[2:].zfill(length)
            assert len(self._s) == length
        elif auto is None:
            self._s = ""
        elif isinstance(auto, int):
            self._s = "0" * auto            # bitstring: an integer creates that many zero bits
        elif isinstance(auto, str):
            assert auto.startswith("0b")
            self._s = auto[2:]
        elif isinstance(auto, _Bits):
            self._s = auto._s
        else:
            raise TypeError(type(auto))

    def __len__(self):
        return len(self._s)

    def __getitem__(self, key):
        out = _Bits()
        out._s = self._s[key]
        return out

    @property
    def int(self):
        if not self._s:
            raise ValueError("empty")
        v = int(self._s, 2)
        if self._s[0] == "1":
            v -= 1 << len(self._s)
        return v

    @property
    def bin(self):
        return self._s

    def tobytes(self):
        s = self._s + "0" * ((-len(self._s)) % 8)
        return int(s, 2).to_bytes(len(s) // 8, "big") if s else b""

    @property
    def bytes(self):
        assert len(self._s) % 8 == 0
        return self.tobytes()


builtins_bytes = bytes


class _BitArray(_Bits):
    def append(self, other):
        self._s = self._s + _Bits(other)._s

    def prepend(self, other):
        self._s = _Bits(other)._s + self._s

    def __ilshift__(self, n):
        n = min(n, len(self._s))
        self._s = self._s[n:] + "0" * n
        return self


def install():
    """Idempotently install the stand-ins and the reference path."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "rawpy" not in sys.modules:
        sys.modules["rawpy"] = types.ModuleType("rawpy")
    if "bitstring" not in sys.modules:
        m = types.ModuleType("bitstring")
        m.Bits = _Bits
        m.BitArray = _BitArray
        sys.modules["bitstring"] = m
    if "pywt" not in sys.modules:
        from oracle import pywt_standin
        sys.modules["pywt"] = pywt_standin
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # a hiccup alias installed by hiccup_b200._compat must not shadow the real package
    for name in [k for k in sys.modules if k == "hiccup" or k.startswith("hiccup.")]:
        mod = sys.modules[name]
        # (the alias package itself carries the flag; its sub-modules are hiccup_b200's own modules under a second name)
        if getattr(mod, "__hiccup_b200_alias__", False) or getattr(mod, "__name__", "").startswith("hiccup_b200"):
            del sys.modules[name]
    import hiccup.settings as settings
    settings.DEBUG = False
    return settings

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

Two things the reference's reader does not have:

  * A tolerant, closed reader (`loads`).  The pickles inside a `.hic` file name a class
    (`hiccup.hicimage.TupP`, hicimage.py:117-121) and -- for DC symbols -- numpy's scalar
    reconstructor, whose module path differs between numpy 1 (`numpy.core.multiarray`) and numpy 2
    (`numpy._core.multiarray`), so a file written in one environment does not load in another with
    plain `pickle.loads`.  `loads` resolves exactly the handful of globals a `.hic` file can contain,
    whatever environment wrote it (any pickle protocol, either numpy spelling, with or without the
    reference package importable), and refuses every other global -- a `.hic` file from an untrusted
    source cannot run code here.  `HicImage.write_file(path, portable=True)` writes symbols as plain
    Python ints so that ANY environment's reference reader loads the file (not byte-identical to the
    reference's own output, hence opt-in).
  * Extension entries.  The reference's reader takes list entries 0..20 (DCT) or 0..14 (wavelet) and
    never looks further (hicimage.py:124-142), so entries appended after them travel with the file and
    the reference still decodes it.  `RestartP` (restart records for the parallel Huffman decode,
    include/hiccup_b200.h) is such an entry.

The format's one pickle per table row (thousands per image) is written and parsed by host code of the library
(`_NativeRows` over csrc/hic_hicfile.cu: rows, whole table payloads, and whole files of a batch on host threads), each
level calibrated against `pickle.dumps` in this environment before it is used and replaced by plain pickle otherwise.
"""
import ctypes
import io
import os
import pickle
import struct
import threading

import numpy as np

from hiccup_b200 import _compat, iohelper, model


def _np_scalar(dtype, raw):
    """numpy's scalar reconstructor, for either spelling of its module path."""
    return np.frombuffer(raw, dtype=dtype, count=1)[0]


class _HicUnpickler(pickle.Unpickler):
    """Resolves the globals a `.hic` pickle may name and nothing else."""

    def find_class(self, module, name):
        if name == "TupP" and module.endswith("hicimage"):
            return _compat.wire_tuple_class()
        if module in ("numpy.core.multiarray", "numpy._core.multiarray") and name == "scalar":
            return _np_scalar
        if module == "numpy" and name == "dtype":
            return np.dtype
        if module == "_codecs" and name == "encode":           # how protocols < 3 spell a bytes object
            import codecs
            return codecs.encode
        raise pickle.UnpicklingError("a .hic file may not reference %s.%s" % (module, name))


def loads(b):
    """pickle.loads for the pickles inside a `.hic` file: tolerant of the writer's environment, closed to
    everything a `.hic` file has no business containing."""
    return _HicUnpickler(io.BytesIO(bytes(b))).load()


def load(f):
    return _HicUnpickler(f).load()


class _RowCodec:
    """The pickle of one Huffman table row, `pickle.dumps((symbol, code))` (hicimage.py:57-60 through TupP), written
    and parsed directly.  A table has hundreds to thousands of rows and the format pickles each one on its own: for
    one 512x512 image that is 3 600 `pickle.dumps` + 3 600 unpicklings, three quarters of the time of a whole
    encode + decode call on the GPU.  A row is (int | np.int32, str of '0'/'1'), whose protocol-4 pickle has a fixed
    shape; the byte strings built here are calibrated against `pickle.dumps` when the module loads (prefix and
    infix of the numpy scalar form are READ from a sample, not assumed) and verified on a set of samples -- if
    anything differs (another default protocol, another numpy), `ok` stays False and every row goes through
    `pickle` as before.  The parser accepts exactly the canonical forms and hands anything else to the closed
    unpickler above."""
    HEAD = b"\x80\x04\x95"
    TAIL = b"\x94\x86\x94."

    def __init__(self):
        self.ok = False
        self.np_pre = self.np_mid = None
        try:
            self._calibrate()
        except Exception:
            self.ok = False

    def _calibrate(self):
        if pickle.DEFAULT_PROTOCOL != 4:
            return
        sample = pickle.dumps((np.int32(0x01020304), "01"))
        marker = b"C\x04" + (0x01020304).to_bytes(4, "little")
        pos = sample.index(marker)
        self.np_pre = sample[11:pos + 2]
        rest = sample[pos + 6:]
        cut = rest.index(b"\x8c\x0201")
        self.np_mid = rest[:cut]
        if rest[cut + 4:] != self.TAIL:
            return
        self.ok = True
        codes = ["0", "1", "101", "0" * 58, "1100110011"]
        syms = [0, 1, 5, 255, 256, 300, 65535, 65536, 70000, -1, -7, -300, -70000, 2 ** 31 - 1, -2 ** 31]
        for code in codes:
            for v in syms:
                for sym in ((v, np.int32(v))):
                    want = pickle.dumps((sym, code))
                    got = self.dumps(sym, code)
                    back = self.loads(want)
                    if got != want or back is None or type(back[0]) is not type(sym) or back[0] != sym or back[1] != code:
                        self.ok = False
                        return

    def dumps(self, sym, code):
        """The bytes `pickle.dumps((sym, code))` would give, or None when the row is not of the fast shape."""
        if not self.ok or type(code) is not str or len(code) > 255:
            return None
        t = type(sym)
        if t is int:
            if 0 <= sym < 256:
                e = b"K" + bytes((sym,))
            elif 256 <= sym < 65536:
                e = b"M" + sym.to_bytes(2, "little")
            elif -2147483648 <= sym < 2147483648:
                e = b"J" + sym.to_bytes(4, "little", signed=True)
            else:
                return None
        elif t is np.int32:
            e = self.np_pre + int(sym).to_bytes(4, "little", signed=True) + self.np_mid
        else:
            return None
        try:
            c = code.encode("ascii")
        except UnicodeEncodeError:
            return None
        body = e + b"\x8c" + bytes((len(c),)) + c + self.TAIL
        return self.HEAD + len(body).to_bytes(8, "little") + body

    def loads(self, row):
        """(sym, code) of a canonical row pickle, or None (the caller then uses the closed unpickler)."""
        if not self.ok:
            return None
        row = bytes(row)
        n = len(row)
        if n < 20 or row[:3] != self.HEAD or int.from_bytes(row[3:11], "little") != n - 11 or row[-4:] != self.TAIL:
            return None
        op = row[11]
        if op == 0x4B:                                   # BININT1
            sym, p = row[12], 13
        elif op == 0x4D:                                 # BININT2
            sym, p = int.from_bytes(row[12:14], "little"), 14
        elif op == 0x4A:                                 # BININT
            sym, p = int.from_bytes(row[12:16], "little", signed=True), 16
        elif row.startswith(self.np_pre, 11):
            a = 11 + len(self.np_pre)
            p = a + 4 + len(self.np_mid)
            if row[a + 4:p] != self.np_mid:
                return None
            sym = np.int32(int.from_bytes(row[a:a + 4], "little", signed=True))
        else:
            return None
        if p + 2 > n - 4 or row[p] != 0x8C or p + 2 + row[p + 1] != n - 4:
            return None
        try:
            return sym, row[p + 2:n - 4].decode("ascii")
        except UnicodeDecodeError:
            return None


ROWS = _RowCodec()


class _NativeRows:
    """The same row pickles written and parsed for a whole table at once by the C library (csrc/hic_hicfile.cu; host
    code, no device): a table is then three arrays (symbols, code lengths, codes) + a per-row flag "the symbol is a
    numpy.int32 scalar", and the thousands of Python objects per image never exist unless somebody asks for `.rows`.
    Calibrated like _RowCodec: `ok` only if the library's bytes equal pickle.dumps on the sample set."""

    def __init__(self):
        self.ok = False
        self.table_ok = False
        self.files_ok = False
        self._heads = {}
        self._tls = threading.local()
        try:
            self._calibrate()
        except Exception:
            self.ok = False
        try:
            self._calibrate_tables()
        except Exception:
            self.table_ok = False
        try:
            self._calibrate_files()
        except Exception:
            self.files_ok = False

    def _calibrate(self):
        if not ROWS.ok:
            return
        from hiccup_b200 import _lib
        self._lib = _lib
        self._fn = _lib.load()
        self._pre = np.frombuffer(ROWS.np_pre, np.uint8).copy()
        self._mid = np.frombuffer(ROWS.np_mid, np.uint8).copy()
        self.ok = True
        syms = np.array([0, 1, 255, 256, 65535, 65536, -1, -300, 2 ** 31 - 1, -2 ** 31, 7, 7], np.int32)
        lens = np.array([1, 3, 58, 10, 2, 5, 7, 1, 20, 33, 4, 4], np.uint8)
        codes = np.array([1, 5, (1 << 57) | 1, 0x2AA, 2, 17, 100, 0, 12345, 1 << 32, 9, 9], np.uint64)
        flags = np.array([0, 1] * 6, np.uint8)
        want = [pickle.dumps(((np.int32(v) if f else int(v)), format(int(c), "0%db" % int(n))))
                for v, n, c, f in zip(syms, lens, codes, flags)]
        got = self.pack(syms, lens, codes, flags)
        back = self.parse(want)
        if got != want or back is None or not all(np.array_equal(a, b) for a, b in zip(back, (syms, lens, codes, flags))):
            self.ok = False

    def pack(self, symbols, lens, codes, flags):
        """[bytes] of the rows, byte-identical to pickle.dumps((symbol, code)) of each."""
        n = int(symbols.size)
        if n == 0:
            return []
        cap = n * (11 + len(self._pre) + 4 + len(self._mid) + 2 + 58 + 4)
        out = np.empty(cap, np.uint8)
        off = np.empty(n + 1, np.uint64)
        symbols = np.ascontiguousarray(symbols, np.int32)
        lens = np.ascontiguousarray(lens, np.uint8)
        codes = np.ascontiguousarray(codes, np.uint64)
        flags = np.ascontiguousarray(flags, np.uint8)
        self._lib.check(self._fn.hic_hicfile_pack_rows(symbols.ctypes.data, lens.ctypes.data, codes.ctypes.data, n, flags.ctypes.data,
                                                       self._pre.ctypes.data, self._pre.size, self._mid.ctypes.data, self._mid.size,
                                                       out.ctypes.data, cap, off.ctypes.data))
        raw = out[:int(off[n])].tobytes()
        o = off.tolist()
        return [raw[a:b] for a, b in zip(o, o[1:])]

    # ---- whole table payloads --------------------------------------------------------------------
    def _head(self, cls):
        """The bytes pickle puts between PROTO/FRAME and the first row of {"type": cls, "data": [...]}, or None if this
        pickle module does not lay the empty table out the way csrc/hic_hicfile.cu continues it."""
        if cls not in self._heads:
            raw = pickle.dumps({"type": cls, "data": []})
            ok = (raw[:3] == b"\x80\x04\x95" and int.from_bytes(raw[3:11], "little") == len(raw) - 11 and raw[-4:] == b"]\x94u."
                  and raw[11:13] == b"}\x94")
            self._heads[cls] = np.frombuffer(raw[11:-2], np.uint8).copy() if ok else None
        return self._heads[cls]

    def _calibrate_tables(self):
        """pack_table against pickle.dumps: no row, one row (APPEND, no MARK), a batch boundary (1000 rows), and tables
        past one and two 64 KiB frames, with both kinds of symbol."""
        if not self.ok or pickle.DEFAULT_PROTOCOL != 4:
            return
        cls = _compat.wire_tuple_class()
        self.table_ok = True
        rng = np.random.default_rng(7)
        for n, flag in ((0, 0), (1, 0), (1, 1), (2, 1), (999, 0), (1000, 1), (1001, 0), (2300, 0), (2300, 1)):
            lens = rng.integers(1, 59, n).astype(np.uint8)
            codes = rng.integers(0, 1 << 62, n, dtype=np.uint64) & ((np.uint64(1) << lens.astype(np.uint64)) - np.uint64(1))
            syms = rng.integers(-70000, 70000, n).astype(np.int32)
            flags = np.full(n, flag, np.uint8)
            want = pickle.dumps({"type": cls, "data": self.pack(syms, lens, codes, flags)})
            back = self.parse_table(want)
            if self.pack_table(cls, syms, lens, codes, flags) != want or back is None or \
                    not all(np.array_equal(a, b) for a, b in zip(back, (syms, lens, codes, flags))):
                self.table_ok = False
                return

    def pack_table(self, cls, symbols, lens, codes, flags):
        """bytes of a whole table payload: pickle.dumps({"type": cls, "data": [row pickles]})."""
        head = self._head(cls)
        if head is None:
            return None
        n = int(symbols.size)
        row_max = 11 + len(self._pre) + 4 + len(self._mid) + 2 + 58 + 4
        cap = 64 + head.size + n * (row_max + 6) + 2 * (n // 1000 + 1)
        cap += 9 * (cap // 65536 + 2)
        tls = self._tls
        out = getattr(tls, "out", None)                    # per-thread scratch, grown on demand
        if out is None or out.size < cap:
            out = tls.out = np.empty(max(cap, 1 << 20), np.uint8)
            cap = out.size
        size = ctypes.c_uint64(0)
        symbols = np.ascontiguousarray(symbols, np.int32)
        lens = np.ascontiguousarray(lens, np.uint8)
        codes = np.ascontiguousarray(codes, np.uint64)
        flags = np.ascontiguousarray(flags, np.uint8)
        self._lib.check(self._fn.hic_hicfile_pack_table(symbols.ctypes.data, lens.ctypes.data, codes.ctypes.data, n, flags.ctypes.data,
                                                        self._pre.ctypes.data, self._pre.size, self._mid.ctypes.data, self._mid.size,
                                                        head.ctypes.data, head.size, out.ctypes.data, cap, ctypes.byref(size)))
        return out[:size.value].tobytes()

    # ---- whole files of a batch ---------------------------------------------------------------------
    def pack_files(self, cls, index, symbols, packed, data, byte_off, byte_len, stream_of, flag_mode, lead, trail, threads=None, reuse=None):
        """The `.hic` files of a batch in one call, written by host threads of the library: file i is
        pickle.dumps([lead, tables..., bit strings..., *trail]) with its k-th table and bit string taken from symbol stream
        stream_of[i][k] of an encode result (index / symbols / packed: the packed table layout; data / byte_off / byte_len:
        the framed payloads); flag_mode[k] as in include/hiccup_b200.h.  Returns (buffer, offsets, sizes): file i is
        buffer[offsets[i]:offsets[i] + sizes[i]]; sizes[i] == 0 where the library left the file to the caller.
        reuse: a dict that keeps the output buffer between calls (the files of the previous call are then overwritten):
        a fresh buffer of a batch's size costs as much in page faults as the writing itself."""
        head = self._head(cls)
        if head is None or not self.files_ok:
            return None
        lib = self._lib
        stream_of = np.ascontiguousarray(stream_of, np.uint32)
        n, tables = stream_of.shape
        flag_mode = np.ascontiguousarray(flag_mode, np.uint8)
        assert flag_mode.size == tables
        index = np.ascontiguousarray(index, np.uint32)
        symbols = np.ascontiguousarray(symbols, np.int32)
        packed = np.ascontiguousarray(packed, np.uint64)
        data = np.ascontiguousarray(data, np.uint8)
        byte_off = np.ascontiguousarray(byte_off, np.uint64)
        byte_len = np.ascontiguousarray(byte_len, np.uint64)
        lead_a = np.frombuffer(bytes(lead), np.uint8)
        trail_a = np.frombuffer(b"".join(trail) or b"\0", np.uint8)
        trail_len = np.array([len(t) for t in trail] or [0], np.uint64)
        env = lib.HicfileEnv(self._pre.ctypes.data, self._mid.ctypes.data, head.ctypes.data, self._pre.size, self._mid.size, head.size, 0)
        batch = lib.HicfileBatch(n, tables, len(trail), stream_of.ctypes.data, flag_mode.ctypes.data, index.ctypes.data, symbols.ctypes.data,
                                 packed.ctypes.data, data.ctypes.data, byte_off.ctypes.data, byte_len.ctypes.data, lead_a.ctypes.data,
                                 lead_a.size, trail_a.ctypes.data, trail_len.ctypes.data)
        if threads is None:
            threads = min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 32)
        bound = np.empty(n, np.uint64)
        lib.check(self._fn.hic_hicfile_files_bound(ctypes.byref(env), ctypes.byref(batch), bound.ctypes.data, int(threads)))
        off = np.zeros(n + 1, np.uint64)
        np.cumsum((bound + np.uint64(63)) & ~np.uint64(63), out=off[1:])
        out = None if reuse is None else reuse.get("out")
        if out is None or out.size < int(off[n]):
            out = np.empty(int(off[n]) + (int(off[n]) // 8 if reuse is not None else 0), np.uint8)
            if reuse is not None:
                reuse["out"] = out
        sizes = np.zeros(n, np.uint64)
        lib.check(self._fn.hic_hicfile_pack_files(ctypes.byref(env), ctypes.byref(batch), out.ctypes.data, off.ctypes.data,
                                                  sizes.ctypes.data, int(threads)))
        return out, off[:-1], sizes

    def parse_files(self, files, stream_of, n_streams, n_trail, threads=None):
        """The reverse of pack_files for a list of bytes-like `.hic` files with stream_of.shape[1] tables each: returns
        (index, symbols, packed, data, byte_off, byte_len, nbits, lead, trail) -- the encode-result layout over n_streams
        symbol streams (streams no file names stay empty) plus the mode entry and the n_trail entries after the bit strings
        of every file as lists of bytes -- or None if any file is not in the canonical form (the caller then reads the
        batch with the unpickler)."""
        if not self.files_ok:
            return None
        lib = self._lib
        stream_of = np.ascontiguousarray(stream_of, np.uint32)
        n, tables = stream_of.shape
        assert n == len(files)
        n_items = 1 + 2 * tables + n_trail
        views = [np.frombuffer(f, np.uint8) for f in files]         # (kept: the pointers below are theirs)
        sizes = np.fromiter((v.size for v in views), np.uint64, n)
        ptrs = np.fromiter((v.ctypes.data if v.size else 0 for v in views), np.uint64, n)
        if threads is None:
            threads = min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 32)
        item_off, item_len = np.zeros((n, n_items), np.uint64), np.zeros((n, n_items), np.uint64)
        rows, canonical = np.zeros((n, tables), np.uint32), np.zeros(n, np.uint8)
        lib.check(self._fn.hic_hicfile_scan_files(ptrs.ctypes.data, sizes.ctypes.data, n, tables, n_items, item_off.ctypes.data,
                                                  item_len.ctypes.data, rows.ctypes.data, canonical.ctypes.data, int(threads)))
        if not canonical.all():
            return None
        index = np.zeros((n_streams, 2), np.uint32)
        index[stream_of.reshape(-1), 1] = rows.reshape(-1)
        total = int(index[:, 1].sum(dtype=np.uint64))
        if total >= 1 << 32:
            return None
        index[1:, 0] = np.cumsum(index[:-1, 1], dtype=np.uint64).astype(np.uint32)
        byte_len = np.zeros(n_streams, np.uint64)
        byte_len[stream_of.reshape(-1)] = item_len[:, 1 + tables:1 + 2 * tables].reshape(-1)
        padded = (byte_len + np.uint64(3)) & ~np.uint64(3)
        byte_off = np.zeros(n_streams, np.uint64)
        byte_off[1:] = np.cumsum(padded[:-1])
        data = np.zeros(int(padded.sum()) + 16, np.uint8)
        symbols, packed = np.empty(max(1, total), np.int32), np.empty(max(1, total), np.uint64)
        nbits, ok = np.zeros(n_streams, np.uint64), np.zeros(n, np.uint8)
        head = self._head(_compat.wire_tuple_class())
        env = lib.HicfileEnv(self._pre.ctypes.data, self._mid.ctypes.data, None if head is None else head.ctypes.data,
                             self._pre.size, self._mid.size, 0 if head is None else head.size, 0)
        lib.check(self._fn.hic_hicfile_parse_files(ctypes.byref(env), ptrs.ctypes.data, n, tables, n_items, item_off.ctypes.data,
                                                   item_len.ctypes.data, stream_of.ctypes.data, index.ctypes.data, symbols.ctypes.data,
                                                   packed.ctypes.data, data.ctypes.data, byte_off.ctypes.data, nbits.ctypes.data,
                                                   ok.ctypes.data, int(threads)))
        if not ok.all():
            return None
        entry = lambda i, e: views[i][int(item_off[i, e]):int(item_off[i, e] + item_len[i, e])].tobytes()
        lead = [entry(i, 0) for i in range(n)]
        trail = [[entry(i, 1 + 2 * tables + t) for t in range(n_trail)] for i in range(n)]
        return index, symbols[:total], packed[:total], data, byte_off, byte_len, nbits, lead, trail

    def _calibrate_files(self):
        """pack_files against pickle.dumps of the list of entries: entries below 256 bytes, below and above the 64 KiB
        frame size, in both flag modes."""
        self.files_ok = False
        if not self.table_ok:
            return
        self.files_ok = True
        cls = _compat.wire_tuple_class()
        rng = np.random.default_rng(3)
        rows = np.array([3, 1200, 2, 2300, 40, 1], np.uint32)
        first = np.concatenate(([0], np.cumsum(rows)[:-1])).astype(np.uint32)
        total = int(rows.sum())
        symbols = rng.integers(-300, 70000, total).astype(np.int32)
        symbols[::7] = 0
        lens = rng.integers(1, 59, total).astype(np.uint64)
        packed = (lens << np.uint64(58)) | (rng.integers(0, 1 << 62, total, dtype=np.uint64) & ((np.uint64(1) << lens) - np.uint64(1)))
        sizes = np.array([5, 300, 70000, 2, 65536, 65535], np.uint64)
        byte_off = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.uint64)
        data = rng.integers(0, 256, int(sizes.sum()), dtype=np.uint8)
        stream_of = np.array([[0, 1, 2], [3, 4, 5], [5, 3, 1]], np.uint32)
        modes = [1, 0, 2]
        trail = [pickle.dumps((426, 640)), pickle.dumps((213, 320))]
        res = self.pack_files(cls, np.stack([first, rows], 1), symbols, packed, data, byte_off, sizes, stream_of, modes, b"JPEG", trail, threads=2)
        ok = res is not None
        if ok:
            out, off, got = res
            for i in range(stream_of.shape[0]):
                entries = [b"JPEG"]
                for k, s_ in enumerate(stream_of[i]):
                    a, b = int(first[s_]), int(first[s_] + rows[s_])
                    sym = symbols[a:b]
                    flags = {0: np.zeros(b - a, np.uint8), 1: np.ones(b - a, np.uint8), 2: (sym != 0).astype(np.uint8)}[modes[k]]
                    entries.append(self.pack_table(cls, sym, (packed[a:b] >> np.uint64(58)).astype(np.uint8),
                                                   packed[a:b] & np.uint64((1 << 58) - 1), flags))
                entries += [data[int(byte_off[s_]):int(byte_off[s_] + sizes[s_])].tobytes() for s_ in stream_of[i]]
                entries += trail
                if out[int(off[i]):int(off[i] + got[i])].tobytes() != pickle.dumps(entries):
                    ok = False
        if ok:                                             # and back: files 0 and 1 name every stream once
            back = self.parse_files([out[int(off[i]):int(off[i] + got[i])] for i in range(2)], stream_of[:2], 6, 2, threads=2)
            ok = back is not None and np.array_equal(back[0][:, 1], rows) and np.array_equal(back[1], symbols) and \
                np.array_equal(back[2], packed) and np.array_equal(back[5], sizes) and back[7] == [b"JPEG"] * 2 and back[8] == [trail] * 2 and \
                all(np.array_equal(back[3][int(a):int(a + n_)], data[int(b):int(b + n_)]) for a, b, n_ in zip(back[4], byte_off, sizes))
        self.files_ok = ok

    def parse_table(self, payload):
        """(symbols, lens, codes, flags) of a whole table payload, or None if it is not in the canonical form (another
        pickle protocol, a foreign class, a row of another shape, ...) -- the caller then takes the unpickler."""
        size = len(payload)
        cap = size // 23 + 1
        symbols, lens = np.empty(cap, np.int32), np.empty(cap, np.uint8)
        codes, flags = np.empty(cap, np.uint64), np.empty(cap, np.uint8)
        n, canonical = ctypes.c_uint64(0), ctypes.c_int32(0)
        data = np.frombuffer(payload, np.uint8) if size else np.zeros(1, np.uint8)
        self._lib.check(self._fn.hic_hicfile_parse_table(data.ctypes.data, size, self._pre.ctypes.data, self._pre.size,
                                                         self._mid.ctypes.data, self._mid.size, symbols.ctypes.data, lens.ctypes.data,
                                                         codes.ctypes.data, flags.ctypes.data, cap, ctypes.byref(n), ctypes.byref(canonical)))
        if not canonical.value:
            return None
        k = int(n.value)
        return symbols[:k].copy(), lens[:k].copy(), codes[:k].copy(), flags[:k].copy()

    def parse(self, rows):
        """(symbols, lens, codes, flags) of a list of row pickles, or None if any row is not canonical."""
        import ctypes
        n = len(rows)
        sizes = np.fromiter((len(r) for r in rows), np.uint64, n)
        off = np.zeros(n + 1, np.uint64)
        np.cumsum(sizes, out=off[1:])
        data = np.frombuffer(b"".join(rows), np.uint8) if n else np.zeros(1, np.uint8)
        symbols, lens = np.empty(n, np.int32), np.empty(n, np.uint8)
        codes, flags = np.empty(n, np.uint64), np.empty(n, np.uint8)
        bad = ctypes.c_int64(-1)
        self._lib.check(self._fn.hic_hicfile_parse_rows(data.ctypes.data, off.ctypes.data, n, self._pre.ctypes.data, self._pre.size,
                                                        self._mid.ctypes.data, self._mid.size, symbols.ctypes.data, lens.ctypes.data,
                                                        codes.ctypes.data, flags.ctypes.data, ctypes.byref(bad)))
        if bad.value >= 0:
            return None
        return symbols, lens, codes, flags


NATIVE = None


def _native():
    global NATIVE
    if NATIVE is None:
        NATIVE = _NativeRows()
    return NATIVE


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
        fast = ROWS.loads(b)
        a, c = fast if fast is not None else loads(b)
        return cls(a, c)

    @property
    def numbers(self):
        return self.n1, self.n2

    @property
    def byte_stream(self):
        fast = ROWS.dumps(self.n1, self.n2)
        return fast if fast is not None else pickle.dumps((self.n1, self.n2))

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
    """A run of payloads of one type -- in practice the rows of one Huffman table.  Held either as the reference
    holds it (a list of TupP) or as arrays (`from_arrays`, and `from_bytes` when every row is canonical); the other
    form is made when somebody asks for it."""

    def __init__(self, t, payloads):
        self.t, self._payloads, self._arrays = t, payloads, None

    @classmethod
    def from_arrays(cls, symbols, lens, codes, numpy_scalar):
        """symbols int32, lens uint8 (1..58), codes uint64; numpy_scalar: bool or per-row uint8 -- the symbols are
        numpy.int32 scalars in the pickles (DC tables, non-zero wavelet values) rather than Python ints."""
        obj = cls(_compat.wire_tuple_class(), None)
        symbols = np.ascontiguousarray(symbols, np.int32)
        flags = (np.full(symbols.size, 1 if numpy_scalar else 0, np.uint8) if np.isscalar(numpy_scalar) or isinstance(numpy_scalar, bool)
                 else np.ascontiguousarray(numpy_scalar, np.uint8))
        obj._arrays = (symbols, np.ascontiguousarray(lens, np.uint8), np.ascontiguousarray(codes, np.uint64), flags)
        return obj

    @classmethod
    def from_bytes(cls, b):
        nat = _native()
        if nat.table_ok:
            arrays = nat.parse_table(b)
            if arrays is not None:
                obj = cls(_compat.wire_tuple_class(), None)
                obj._arrays = arrays
                return obj
        d = loads(b)
        # rows are plain pickled pairs; parse them directly rather than through d["type"] so that
        # files written by the reference and by this package read the same way
        data = d["data"]
        if nat.ok and data:
            arrays = nat.parse([bytes(x) for x in data])
            if arrays is not None:
                obj = cls(d["type"], None)
                obj._arrays = arrays
                return obj
        return cls(d["type"], [TupP.from_bytes(x) for x in data])

    @classmethod
    def from_rows(cls, rows):
        """rows: iterable of (symbol, code string)."""
        return cls(_compat.wire_tuple_class(), [TupP(s, c) for s, c in rows])

    @property
    def payloads(self):
        if self._payloads is None:
            sym, lens, codes, flags = self._arrays
            self._payloads = [TupP(np.int32(v) if f else v, format(c, "0%db" % n))
                              for v, n, c, f in zip(sym.tolist(), lens.tolist(), codes.tolist(), flags.tolist())]
        return self._payloads

    def arrays(self):
        """(symbols int32, lens uint8, codes uint64, numpy-scalar flags uint8), or None when a row does not fit them
        (a float symbol, a code of other characters, ...)."""
        if self._arrays is None and self._payloads is not None:
            try:
                sym, lens, codes, flags = [], [], [], []
                for p in self._payloads:
                    a, c = p.n1, p.n2
                    t = type(a)
                    if t is int:
                        flags.append(0)
                    elif t is np.int32:
                        flags.append(1)
                    else:
                        return None
                    if type(c) is not str or not 1 <= len(c) <= 58 or not -2147483648 <= int(a) < 2147483648:
                        return None
                    sym.append(int(a))
                    lens.append(len(c))
                    codes.append(int(c, 2))
                self._arrays = (np.array(sym, np.int32), np.array(lens, np.uint8), np.array(codes, np.uint64), np.array(flags, np.uint8))
            except ValueError:
                return None
        return self._arrays

    @property
    def rows(self):
        return [p.numbers for p in self.payloads]

    @property
    def byte_stream(self):
        nat = _native()
        arrays = self._arrays if self._arrays is not None else (self.arrays() if nat.ok else None)
        if nat.table_ok and arrays is not None:
            whole = nat.pack_table(_compat.wire_tuple_class(), *arrays)
            if whole is not None:
                return whole
        if nat.ok and arrays is not None:
            rows = nat.pack(*arrays)
        else:
            rows = [p.byte_stream for p in self.payloads]
        return pickle.dumps({"type": _compat.wire_tuple_class(), "data": rows})

    def portable(self):
        """The same table with every symbol as a plain Python number (no numpy scalar pickles)."""
        if self._arrays is not None:
            sym, lens, codes, _ = self._arrays
            return PayloadStringP.from_arrays(sym, lens, codes, False)
        plain = lambda v: v if type(v) in (int, float) else (float(v) if isinstance(v, (float, np.floating)) else int(v))
        return PayloadStringP(self.t, [TupP(plain(p.n1), p.n2) for p in self.payloads])

    def __eq__(self, other):
        return hasattr(other, "payloads") and list(other.payloads) == list(self.payloads)

    __hash__ = None


class RestartP(Payload):
    """Extension entry: restart records of the file's bit payloads, in the file's payload order.
    records[i] = (off, cnt), two uint8 arrays with one element per 128-bit subsequence of the framed
    payload i (include/hiccup_b200.h: hic_decode_run_restarts).  12.5 % of the coded size."""
    MAGIC = b"hiccup_b200.restart.v1\0"

    def __init__(self, records):
        self.records = [(np.ascontiguousarray(o, np.uint8), np.ascontiguousarray(c, np.uint8)) for o, c in records]

    @classmethod
    def matches(cls, b):
        return bytes(b[:len(cls.MAGIC)]) == cls.MAGIC

    @classmethod
    def from_bytes(cls, b):
        b = bytes(b)
        if not cls.matches(b):
            raise ValueError("not a restart entry")
        pos = len(cls.MAGIC)
        (n,) = struct.unpack_from("<I", b, pos)
        pos += 4
        records = []
        for _ in range(n):
            (k,) = struct.unpack_from("<I", b, pos)
            pos += 4
            if pos + 2 * k > len(b):
                raise ValueError("truncated restart entry")
            records.append((np.frombuffer(b, np.uint8, k, pos), np.frombuffer(b, np.uint8, k, pos + k)))
            pos += 2 * k
        return cls(records)

    @property
    def byte_stream(self):
        parts = [self.MAGIC, struct.pack("<I", len(self.records))]
        for off, cnt in self.records:
            parts += [struct.pack("<I", off.size), off.tobytes(), cnt.tobytes()]
        return b"".join(parts)

    def __eq__(self, other):
        return isinstance(other, RestartP) and self.byte_stream == other.byte_stream

    __hash__ = None


class HicImage:
    LAYOUT = {model.Compression.JPEG: (9, 9, 2), model.Compression.HIC: (6, 6, 2)}

    def __init__(self, hic_type, settings, payloads, extensions=None):
        self.hic_type, self.settings, self._payloads = hic_type, settings, payloads
        self.extensions = list(extensions or [])       # entries after the reference's own (it never reads them)

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
        out = cls.jpeg_image(payloads) if kind == model.Compression.JPEG else cls.wavelet_image(payloads)
        for x in raw_data[c + n_shape:]:
            out.extensions.append(RestartP.from_bytes(x) if RestartP.matches(x) else bytes(x))
        return out

    @classmethod
    def from_file(cls, path):
        with open(path, "rb") as f:
            raw = load(f)
        assert raw is not None
        return cls.from_bytes(raw)

    @property
    def payloads(self):
        return self._payloads

    @property
    def restarts(self):
        """The file's restart records (a RestartP), or None."""
        for e in self.extensions:
            if isinstance(e, RestartP):
                return e
        return None

    def portable(self):
        """A copy whose tables hold plain Python numbers: loads in any environment's reference reader."""
        n_tab = self.LAYOUT[self.hic_type if isinstance(self.hic_type, model.Compression) else model.Compression(self.hic_type.value)][0]
        payloads = [p.portable() if i < n_tab else p for i, p in enumerate(self._payloads)]
        return HicImage(self.hic_type, self.settings, payloads, self.extensions)

    def byte_stream(self):
        return ([p.byte_stream for p in self.settings + self._payloads]
                + [e if isinstance(e, (bytes, bytearray)) else e.byte_stream for e in self.extensions])

    def write_file(self, path, portable=False):
        with open(path, "wb") as f:
            pickle.dump((self.portable() if portable else self).byte_stream(), f)

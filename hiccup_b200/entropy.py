"""Entropy stage plumbing shared by codec.py (single image, reference signatures) and batch.py.

Wraps the plan objects of include/hiccup_b200.h.  A symbol stream is indexed
s = (image * 3 + channel) * 3 + kind with kind 0 = DC differences, 1 = run-length values,
2 = run-length zero counts (channel order lum, cr, cb as in reference codec.py:304-317).
"""
import ctypes

import numpy as np

from hiccup_b200 import _lib

KIND_DC, KIND_VALUE, KIND_LENGTH = 0, 1, 2
DEFAULT_VALUE_BINS = 8192


class EncodedStreams:
    """Host-side result of the entropy encoder for a batch: code tables and framed bit payloads."""

    def __init__(self, layout, rows, nsym, nbits, byte_off, byte_len, symbols, lens, codes, data):
        self.layout = layout
        self.rows, self.nsym, self.nbits = rows, nsym, nbits
        self.byte_off, self.byte_len = byte_off, byte_len
        self.symbols, self.lens, self.codes = symbols, lens, codes
        self.data = data                                   # uint8, all framed payloads, 4-byte aligned each
        self.row_off = np.concatenate(([0], np.cumsum(rows, dtype=np.int64)))

    def table(self, s):
        """[(symbol, code string)] of symbol stream s in first-occurrence order."""
        a, b = int(self.row_off[s]), int(self.row_off[s + 1])
        return [(int(sym), format(int(code), "0%db" % int(ln)))
                for sym, ln, code in zip(self.symbols[a:b], self.lens[a:b], self.codes[a:b])]

    def framed(self, s):
        """The framed bytes of symbol stream s (what iohelper.padded_bs_2_bytes returns)."""
        a = int(self.byte_off[s])
        return self.data[a:a + int(self.byte_len[s])].tobytes()


class EntropyEncoder:
    """Owns a device plan for one batch shape; reusable across batches."""

    def __init__(self, layout, value_bins=DEFAULT_VALUE_BINS):
        _lib.require_device()
        self.lib = _lib.load()
        self.layout = layout
        self.value_bins = value_bins
        self.n_streams = layout.n_images * 9
        p = ctypes.c_void_p()
        _lib.check(self.lib.hic_entropy_plan_create(ctypes.byref(layout), value_bins, ctypes.byref(p)))
        self.plan = p
        self._out = None

    def close(self):
        if self.plan:
            self.lib.hic_entropy_plan_destroy(self.plan)
            self.plan = None
        if self._out is not None:
            self._out.free()
            self._out = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def symbolize(self, d_coef, stream=None):
        _lib.check(self.lib.hic_entropy_symbolize(self.plan, d_coef, stream))

    def build_codes(self, stream=None, on_device=False):
        """E2.  on_device=False: host heapq replay (reference-faithful default, any alphabet size);
        on_device=True: the same replay by one CTA per stream on the GPU (alphabets <= 8192)."""
        if on_device:
            _lib.check(self.lib.hic_entropy_build_codes_device(self.plan, stream))
        else:
            _lib.check(self.lib.hic_entropy_build_codes(self.plan, stream))
        n = self.n_streams
        self.rows = np.zeros(n, np.uint32)
        self.nsym = np.zeros(n, np.uint32)
        self.nbits = np.zeros(n, np.uint64)
        self.byte_off = np.zeros(n, np.uint64)
        self.byte_len = np.zeros(n, np.uint64)
        tr, tb = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _lib.check(self.lib.hic_entropy_stream_info(self.plan, self.rows.ctypes.data, self.nsym.ctypes.data,
                                                    self.nbits.ctypes.data, self.byte_off.ctypes.data,
                                                    self.byte_len.ctypes.data, ctypes.byref(tr), ctypes.byref(tb)))
        self.total_rows, self.total_bytes = tr.value, tb.value

    def tables(self):
        sym = np.zeros(self.total_rows, np.int32)
        lens = np.zeros(self.total_rows, np.uint8)
        codes = np.zeros(self.total_rows, np.uint64)
        _lib.check(self.lib.hic_entropy_tables(self.plan, sym.ctypes.data, lens.ctypes.data, codes.ctypes.data))
        return sym, lens, codes

    def pack(self, stream=None):
        """Returns the device buffer holding all framed payloads (valid until the next pack)."""
        need = int(self.total_bytes) + 16
        if self._out is None or self._out.nbytes < need:
            if self._out is not None:
                self._out.free()
            self._out = _lib.DeviceBuffer(need + need // 4)
        _lib.check(self.lib.hic_entropy_pack(self.plan, self._out.ptr, stream))
        return self._out

    def device_tables(self):
        """Device pointers (index, row symbols, packed rows) for EntropyDecoder.set_tables_device."""
        ptrs = [ctypes.c_void_p() for _ in range(5)]
        _lib.check(self.lib.hic_entropy_device_tables(self.plan, *[ctypes.byref(q) for q in ptrs]))
        return tuple(q.value for q in ptrs)

    def encode(self, d_coef, stream=None, download=True, on_device=False):
        """Full entropy encode of device-resident zigzag blocks."""
        self.symbolize(d_coef, stream)
        self.build_codes(stream, on_device=on_device)
        out = self.pack(stream)
        if not download:
            _lib.sync(stream)
            return None
        data = out.download(np.uint8, int(self.total_bytes), stream)
        sym, lens, codes = self.tables()
        return EncodedStreams(self.layout, self.rows, self.nsym, self.nbits, self.byte_off, self.byte_len,
                              sym, lens, codes, data)

    def symbol_arrays(self, stream=None):
        """Download DC differences and run-length symbols (tests / band stitching)."""
        d_dc, d_val, d_len = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(self.lib.hic_entropy_symbol_buffers(self.plan, ctypes.byref(d_dc), ctypes.byref(d_val),
                                                       ctypes.byref(d_len)))
        lay = self.layout
        total_blocks = lay.n_images * lay.blocks_per_image
        dc = np.empty(total_blocks, np.int16)
        val = np.empty(total_blocks * 64, np.int16)
        ln = np.empty(total_blocks * 64, np.uint8)
        _lib.check(self.lib.hic_memcpy_d2h(dc.ctypes.data, d_dc, dc.nbytes, stream))
        _lib.check(self.lib.hic_memcpy_d2h(val.ctypes.data, d_val, val.nbytes, stream))
        _lib.check(self.lib.hic_memcpy_d2h(ln.ctypes.data, d_len, ln.nbytes, stream))
        _lib.sync(stream)
        return dc, val, ln


class EntropyDecoder:
    def __init__(self, layout):
        _lib.require_device()
        self.lib = _lib.load()
        self.layout = layout
        self.n_streams = layout.n_images * 9
        p = ctypes.c_void_p()
        _lib.check(self.lib.hic_decode_plan_create(ctypes.byref(layout), ctypes.byref(p)))
        self.plan = p
        self._in = None

    def close(self):
        if self.plan:
            self.lib.hic_decode_plan_destroy(self.plan)
            self.plan = None
        if self._in is not None:
            self._in.free()
            self._in = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tables_device(self, d_index, d_row_sym, d_row_packed, stream=None):
        _lib.check(self.lib.hic_decode_set_tables_device(self.plan, d_index, d_row_sym, d_row_packed, stream))

    def run(self, d_data, byte_off, nbits, d_coef, stream=None):
        byte_off = np.ascontiguousarray(byte_off, np.uint64)
        nbits = np.ascontiguousarray(nbits, np.uint64)
        _lib.check(self.lib.hic_decode_run(self.plan, d_data, byte_off.ctypes.data, nbits.ctypes.data, d_coef, stream))

    def decode(self, rows, symbols, lens, codes, data, byte_off, nbits, d_coef, stream=None, d_data=None):
        """rows/symbols/lens/codes: concatenated code tables; data: uint8 host array holding every
        framed payload at byte_off[s] (4-byte aligned); nbits[s]: payload bits.  Fills d_coef."""
        rows = np.ascontiguousarray(rows, np.uint32)
        symbols = np.ascontiguousarray(symbols, np.int32)
        lens = np.ascontiguousarray(lens, np.uint8)
        codes = np.ascontiguousarray(codes, np.uint64)
        byte_off = np.ascontiguousarray(byte_off, np.uint64)
        nbits = np.ascontiguousarray(nbits, np.uint64)
        _lib.check(self.lib.hic_decode_set_tables(self.plan, rows.ctypes.data, symbols.ctypes.data, lens.ctypes.data,
                                                  codes.ctypes.data, stream))
        if d_data is None:
            data = np.ascontiguousarray(data, np.uint8)
            need = data.nbytes + 16
            if self._in is None or self._in.nbytes < need:
                if self._in is not None:
                    self._in.free()
                self._in = _lib.DeviceBuffer(need + need // 4)
            self._in.upload(data, stream)
            d_data = self._in.ptr
        _lib.check(self.lib.hic_decode_run(self.plan, d_data, byte_off.ctypes.data, nbits.ctypes.data, d_coef, stream))

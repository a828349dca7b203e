"""Entropy-stage entry points, drop-in for reference hiccup/codec.py.

    jpeg_encode(compressed)     codec.py:275-334     jpeg_decode(hic)     codec.py:337-426
    wavelet_encode(compressed)  codec.py:116-163     wavelet_decode(hic)  codec.py:199-239

Same argument meaning and return types (`HicImage` with the reference's payload order).  DC
differencing, run-length coding, histogramming, bit packing and their inverses run in the CUDA
kernels behind include/hiccup_b200.h; the Huffman trees are built on the host by an exact replay of
the reference's heapq construction (csrc/hic_huffman.cuh).
"""
import numpy as np

from hiccup_b200 import _lib, compression, entropy, hicimage, iohelper, model, settings

CHANNELS = ("lum", "cr", "cb")


def _pick_value_bins(max_abs):
    """Smallest power-of-two bin count whose half covers DC differences (up to 2 * max_abs)."""
    need = 2 * int(max_abs) + 2
    bins = entropy.DEFAULT_VALUE_BINS
    while bins // 2 < need:
        bins *= 2
    if bins > 65536:
        raise ValueError("coefficients up to %d do not fit the int16 symbol path" % max_abs)
    return bins


def _integral_planes(planes):
    out = []
    for a in planes:
        a = np.asarray(a)
        if a.dtype.kind == "f":
            r = np.rint(a)
            if not np.array_equal(r, a):
                raise NotImplementedError("non-integral coefficients are not on the CUDA entropy path")
            a = r
        out.append(a)
    return out


def _symbol_types(dtype):
    """Python types the reference's symbols have in the pickled tables (SURVEY hard part 3): DC
    symbols keep the plane's numpy scalar type (codec.py:47-52 on ndarray elements); run-length
    values pass through ndarray.tolist() (transform.py:264) -> Python int/float; zero counts are
    Python ints."""
    dc_type = np.dtype(dtype).type
    ac_type = float if np.dtype(dtype).kind == "f" else int
    return dc_type, ac_type


def jpeg_encode(compressed: model.CompressedImage, restarts: bool = False) -> hicimage.HicImage:
    """restarts=True appends the restart records of the nine bit strings as an extension entry (hicimage.RestartP):
    the reference's reader ignores it, this package's decoder then skips the synchronisation passes."""
    hic = _jpeg_encode(compressed)
    return add_restart_records(hic) if restarts else hic


def _jpeg_encode(compressed: model.CompressedImage) -> hicimage.HicImage:
    settings.check_supported()
    _lib.require_device()
    d = compressed.as_dict if hasattr(compressed, "as_dict") else model.CompressedImage.as_dict.fget(compressed)
    src_dtype = np.asarray(d["lum"]).dtype
    planes = _integral_planes([d[c] for c in CHANNELS])
    max_abs = max(int(np.abs(p).max()) if p.size else 0 for p in planes)
    bins = _pick_value_bins(max_abs)
    comp = model.CompressedImage(*planes)
    coef, g = compression.planes_to_device_blocks(comp)
    layout = _lib.layout_dct(1, g.h, g.w)
    enc = entropy.cached_encoder(layout, bins)
    try:
        res = enc.encode(coef.ptr)
    finally:
        coef.free()
    dc_type, ac_type = _symbol_types(src_dtype)
    tables, bits = [], []
    for kind in (entropy.KIND_DC, entropy.KIND_VALUE, entropy.KIND_LENGTH):
        for c in range(3):
            s = c * 3 + kind
            conv = dc_type if kind == entropy.KIND_DC else (ac_type if kind == entropy.KIND_VALUE else int)
            if conv is int or conv is np.int32:       # the usual case: the table stays three arrays (hicimage.from_arrays)
                sym, lens, codes = res.stream_rows(s)
                tables.append(hicimage.PayloadStringP.from_arrays(sym, lens, codes, conv is np.int32))
            else:
                tables.append(hicimage.PayloadStringP.from_rows([(conv(sym), code) for sym, code in res.table(s)]))
            bits.append(hicimage.BitStringP.from_framed(res.framed(s)))
    lum_shape = tuple(int(v) for v in planes[0].shape)
    cr_shape = tuple(int(v) for v in planes[1].shape)
    payloads = tables + bits + [hicimage.TupP(*lum_shape), hicimage.TupP(*cr_shape)]
    return hicimage.HicImage.jpeg_image(payloads)


def _tables_to_arrays(table_payloads):
    rows, syms, lens, codes = [], [], [], []
    for t in table_payloads:
        arrays = t.arrays() if hasattr(t, "arrays") else None
        if arrays is not None:
            rows.append(int(arrays[0].size))
            syms.append(arrays[0])
            lens.append(arrays[1])
            codes.append(arrays[2])
            continue
        r = t.rows if hasattr(t, "rows") else [p.numbers for p in t.payloads]
        rows.append(len(r))
        syms.append(np.array([int(sym) for sym, _ in r], np.int32))
        lens.append(np.array([len(code) for _, code in r], np.uint8))
        codes.append(np.array([int(code, 2) for _, code in r], np.uint64))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt, copy=False) if xs else np.zeros(0, dt)
    return rows, cat(syms, np.int32), cat(lens, np.uint8), cat(codes, np.uint64)


def _gather_payload_bytes(bit_payloads):
    """Lay the framed payloads out 4-byte aligned in one host array."""
    offs, nbits, chunks, pos = [], [], [], 0
    for b in bit_payloads:
        framed = bytes(b.byte_stream)
        offs.append(pos)
        nbits.append(iohelper.payload_bit_count(framed))
        pad = (-len(framed)) % 4
        chunks.append(framed + b"\0" * pad)
        pos += len(framed) + pad
    data = np.frombuffer(b"".join(chunks) + b"\0" * 16, dtype=np.uint8)
    return data, offs, nbits


def _stream_order(hic):
    """File payload index of the table / bit string of every stream s of the entropy stage, (tables, bits);
    None where the stream does not exist (the DC streams of a wavelet file)."""
    kind_is_jpeg = hic.hic_type == model.Compression.JPEG or getattr(hic.hic_type, "value", None) == "JPEG"
    if kind_is_jpeg:       # reference order is (kind, channel); stream order is s = channel * 3 + kind
        order = [kind * 3 + c for c in range(3) for kind in range(3)]
        return order, [9 + i for i in order], 9
    tabs, bits = [], []
    for c in range(3):     # kind 0 (DC) is absent in flat mode
        tabs += [None, c, 3 + c]
        bits += [None, 6 + c, 9 + c]
    return tabs, bits, 6


def _records_in_stream_order(hic, nbits):
    """The file's restart records re-ordered to stream order and concatenated, or None when the file has none
    (or they do not fit its bit strings: the extension is advisory, the decoder then synchronises by itself)."""
    ext = getattr(hic, "restarts", None)
    if ext is None:
        return None
    _, bits, first = _stream_order(hic)
    n_bits = len(bits) - bits.count(None)
    if len(ext.records) != n_bits:
        return None
    want = entropy.EntropyDecoder.subsequences(nbits)
    offs, cnts = [], []
    for s, i in enumerate(bits):
        if i is None:
            continue
        off, cnt = ext.records[i - first]
        if off.size != want[s] or cnt.size != want[s]:
            return None
        offs.append(off)
        cnts.append(cnt)
    return np.concatenate(offs) if offs else np.zeros(0, np.uint8), np.concatenate(cnts) if cnts else np.zeros(0, np.uint8)


def _entropy_decode(dec, hic, rows, syms, lens, codes, data, offs, nbits, d_coef, stream=None):
    records = _records_in_stream_order(hic, nbits)
    if records is not None:
        try:
            dec.decode(rows, syms, lens, codes, data, offs, nbits, d_coef, stream, restarts=records)
            return True
        except _lib.HicError:
            pass           # records that do not belong to the streams: decode the long way (and fail there if the streams are bad)
    dec.decode(rows, syms, lens, codes, data, offs, nbits, d_coef, stream)
    return False


def add_restart_records(hic: hicimage.HicImage) -> hicimage.HicImage:
    """A copy of `hic` carrying the restart records of its bit strings (an extension entry the reference's reader
    never reads): D1's synchronisation passes run once, here, instead of at every decode."""
    _lib.require_device()
    p = hic.payloads
    tabs, bits, first = _stream_order(hic)
    empty = hicimage.PayloadStringP.from_rows([])
    rows, syms, lens, codes = _tables_to_arrays([empty if i is None else p[i] for i in tabs])
    jpeg = first == 9
    offs, nbits, chunks, pos = [], [], [], 0
    for i in bits:
        if i is None:
            offs.append(0)
            nbits.append(0)
            continue
        framed = bytes(p[i].byte_stream)
        offs.append(pos)
        nbits.append(iohelper.payload_bit_count(framed))
        pad = (-len(framed)) % 4
        chunks.append(framed + b"\0" * pad)
        pos += len(framed) + pad
    data = np.frombuffer(b"".join(chunks) + b"\0" * 16, dtype=np.uint8)
    if jpeg:
        h, w = (int(v) for v in p[18].numbers)
        layout = _lib.layout_dct(1, h, w)
    else:
        from hiccup_b200 import wavelet
        layout = _lib.layout_flat(1, int(wavelet.pyramid_of_file(hic).len))
    dec = entropy.cached_decoder(layout)
    off, cnt = dec.decode(rows, syms, lens, codes, data, offs, nbits, None, sync_only=True)
    n_sub = entropy.EntropyDecoder.subsequences(nbits)
    records, at = {}, 0
    for s, i in enumerate(bits):
        if i is None:
            continue
        records[i - first] = (off[at:at + n_sub[s]], cnt[at:at + n_sub[s]])
        at += n_sub[s]
    ext = hicimage.RestartP([records[k] for k in sorted(records)])
    others = [e for e in hic.extensions if not isinstance(e, hicimage.RestartP)]
    return hicimage.HicImage(hic.hic_type, hic.settings, list(p), others + [ext])


def jpeg_decode(hic: hicimage.HicImage) -> model.CompressedImage:
    settings.check_supported()
    _lib.require_device()
    assert hic.hic_type == model.Compression.JPEG or getattr(hic.hic_type, "value", None) == "JPEG"
    p = hic.payloads
    lum_shape, cr_shape = tuple(p[18].numbers), tuple(p[19].numbers)
    h, w = int(lum_shape[0]), int(lum_shape[1])
    g = _lib.geometry(h, w)
    if (g.hc, g.wc) != (int(cr_shape[0]), int(cr_shape[1])):
        raise ValueError("chroma shape %r does not belong to luminance shape %r" % (cr_shape, lum_shape))
    # reorder the reference's (kind, channel) payload order into stream order s = channel * 3 + kind
    order = [kind * 3 + c for c in range(3) for kind in range(3)]
    rows, syms, lens, codes = _tables_to_arrays([p[i] for i in order])
    data, offs, nbits = _gather_payload_bytes([p[9 + i] for i in order])
    layout = _lib.layout_dct(1, h, w)
    coef = _lib.scratch(g.blocks_per_image * 128)
    dec = entropy.cached_decoder(layout)
    lib = _lib.load()
    try:
        _entropy_decode(dec, hic, rows, syms, lens, codes, data, offs, nbits, coef.ptr)
        lum = _lib.scratch(4 * h * w)
        cr = _lib.scratch(4 * g.hc * g.wc)
        cb = _lib.scratch(4 * g.hc * g.wc)
        _lib.check(lib.hic_blocks_to_planes(coef.ptr, 1, h, w, lum.ptr, cr.ptr, cb.ptr, None))
        # the reference's decoded planes are float64 (transform.izigzag builds them with np.zeros)
        out = model.CompressedImage(lum.download(np.int32, h * w).reshape(h, w).astype(np.float64),
                                    cr.download(np.int32, g.hc * g.wc).reshape(g.hc, g.wc).astype(np.float64),
                                    cb.download(np.int32, g.hc * g.wc).reshape(g.hc, g.wc).astype(np.float64))
        for b in (lum, cr, cb):
            b.free()
    finally:
        coef.free()
    return out


def wavelet_encode(compressed: model.CompressedImage, restarts: bool = False) -> hicimage.HicImage:
    from hiccup_b200 import wavelet
    hic = wavelet.wavelet_encode(compressed)
    return add_restart_records(hic) if restarts else hic


def wavelet_decode(hic: hicimage.HicImage) -> model.CompressedImage:
    from hiccup_b200 import wavelet
    return wavelet.wavelet_decode(hic)

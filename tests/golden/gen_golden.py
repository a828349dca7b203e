"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py [--only NAME ...]

For each case the reference's own functions are called end to end
(compression.jpeg_compression -> codec.jpeg_encode -> HicImage.byte_stream -> HicImage.from_bytes
-> codec.jpeg_decode -> compression.jpeg_decompression), with codec.differential_coding and
codec.run_length_coding wrapped only to RECORD what they return.  Outputs go to
tests/golden/<name>.npz.  Missing third-party imports are stood in by oracle/refshim.py
(bitstring: pinned by the reference's iohelpertest; pywt: UNPINNED stand-in, wavelet cases only).

Environment of the committed vectors: Python 3.12.3, numpy 2.3.5, scipy 1.18.1, opencv 4.13.0.
"""
import argparse
import contextlib
import io
import os
import pickle
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402
from oracle import hiccup_oracle as orc  # noqa: E402


def _cases():
    import cv2
    ref = refshim.REFERENCE_ROOT
    rng = np.random.default_rng(20261018)
    cases = {}
    cases["tiny16"] = ("jpeg", orc.synthetic_image(16, 16, 11))
    cases["flat32"] = ("jpeg", np.full((32, 32, 3), 128, dtype=np.uint8))
    cases["flat32b"] = ("jpeg", np.full((32, 48, 3), 77, dtype=np.uint8))
    cases["syn64"] = ("jpeg", orc.synthetic_image(64, 64, 12))
    cases["syn48x80"] = ("jpeg", orc.synthetic_image(48, 80, 13))
    cases["noise32x64"] = ("jpeg", rng.integers(0, 256, size=(32, 64, 3), dtype=np.uint8))
    cases["syn26x40"] = ("jpeg", orc.synthetic_image(26, 40, 14))      # not multiples of 16: encode only
    cases["syn33x47"] = ("jpeg", orc.synthetic_image(33, 47, 15))
    cases["syn64x106"] = ("jpeg", orc.synthetic_image(64, 106, 16))
    cases["syn128x160"] = ("jpeg", orc.synthetic_image(128, 160, 17))
    cases["gh256"] = ("jpeg", cv2.imread(os.path.join(ref, "hiccup/test/resources/gh.png")))
    cases["lenna512"] = ("jpeg", cv2.imread(os.path.join(ref, "resources/Lenna.png")))
    cases["w_syn64"] = ("wavelet", orc.synthetic_image(64, 64, 21))
    cases["w_syn48x80"] = ("wavelet", orc.synthetic_image(48, 80, 22))
    cases["w_flat32"] = ("wavelet", np.full((32, 32, 3), 128, dtype=np.uint8))
    cases["w_gh256"] = ("wavelet", cv2.imread(os.path.join(ref, "hiccup/test/resources/gh.png")))
    # wavelet mode at settings other than the defaults (settings.py:12-16): (mode, image, settings)
    cases["ws_l2"] = ("wavelet", orc.synthetic_image(64, 96, 31), dict(levels=2, multiplier=2, threshold=3, quality_factor=0.6, haar=True))
    cases["ws_l5"] = ("wavelet", orc.synthetic_image(64, 96, 32), dict(levels=5, multiplier=0.5, threshold=5, quality_factor=0.9, haar=False))
    cases["ws_l4"] = ("wavelet", orc.synthetic_image(50, 38, 33), dict(levels=4, multiplier=1, threshold=0, quality_factor=0.5, haar=False))
    return cases


def _record(codec):
    rec = {"dc": [], "rle": []}
    orig_dc, orig_rle = codec.differential_coding, codec.run_length_coding

    def dc(blocks):
        out = orig_dc(blocks)
        rec["dc"].append(np.array([int(v) for v in out], dtype=np.int32))
        return out

    def rle(arr, max_len=0xF):
        out = orig_rle(arr, max_len=max_len)
        rec["rle"].append((np.array([r.length for r in out], dtype=np.int32),
                           np.array([int(r.value) for r in out], dtype=np.int32)))
        return out

    codec.differential_coding, codec.run_length_coding = dc, rle
    return rec, (orig_dc, orig_rle)


def run_case(name, mode, rgb, cfg=None):
    import json
    import hiccup.compression as compression
    import hiccup.codec as codec
    import hiccup.hicimage as hicimage
    import hiccup.model as model
    import hiccup.settings as rsettings
    out = {"mode": np.array(mode), "rgb": rgb}
    saved = (rsettings.WAVELET, rsettings.WAVELET_NUM_LEVELS, rsettings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER,
             rsettings.WAVELET_THRESHOLD, rsettings.WAVELET_QUALITY_FACTOR)
    if cfg is not None:
        out["settings"] = np.array(json.dumps(cfg))
        rsettings.WAVELET = model.Wavelet.HAAR if cfg["haar"] else model.Wavelet.DAUBECHIE
        rsettings.WAVELET_NUM_LEVELS = cfg["levels"]
        rsettings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER = cfg["multiplier"]
        rsettings.WAVELET_THRESHOLD = cfg["threshold"]
        rsettings.WAVELET_QUALITY_FACTOR = cfg["quality_factor"]
    try:
        _run_case(name, mode, rgb, out, compression, codec, hicimage)
    finally:
        (rsettings.WAVELET, rsettings.WAVELET_NUM_LEVELS, rsettings.WAVELET_SUBBAND_QUANTIZATION_MULTIPLIER,
         rsettings.WAVELET_THRESHOLD, rsettings.WAVELET_QUALITY_FACTOR) = saved


def _run_case(name, mode, rgb, out, compression, codec, hicimage):
    t0 = time.time()
    rec, orig = _record(codec)
    try:
        if mode == "jpeg":
            comp = compression.jpeg_compression(rgb)
            for ch, arr in comp.as_dict.items():
                assert arr.dtype == np.int32
                out["coef_" + ch] = arr
            hi = codec.jpeg_encode(comp)
        else:
            comp = compression.wavelet_compression(rgb)
            for ch, bands in comp.as_dict.items():
                for i, b in enumerate(bands):
                    assert b.dtype == np.int32
                    out["band_%s_%d" % (ch, i)] = b
            hi = codec.wavelet_encode(comp)
    finally:
        codec.differential_coding, codec.run_length_coding = orig
    for i, ch in enumerate(("lum", "cr", "cb")):
        if mode == "jpeg":
            out["dc_" + ch] = rec["dc"][i]
        out["rle_len_" + ch], out["rle_val_" + ch] = rec["rle"][i]
    stream = hi.byte_stream()
    out["hic"] = np.frombuffer(pickle.dumps(stream), dtype=np.uint8)
    print("  %s: encode %.1fs, %d payloads, %d bytes" % (name, time.time() - t0, len(stream),
                                                         sum(len(b) for b in stream)), flush=True)
    # decode through the reference's own reader
    t0 = time.time()
    hi2 = hicimage.HicImage.from_bytes(stream)
    try:
        with contextlib.redirect_stdout(io.StringIO()):      # codec.py:406-408 prints the AC array
            dec = codec.jpeg_decode(hi2) if mode == "jpeg" else codec.wavelet_decode(hi2)
        if mode == "jpeg":
            same = dec == comp
        else:   # CompressedImage.__eq__ (model.py:68-74) cannot compare ragged sub-band lists
            same = all(np.array_equal(a, b) for ch in ("lum", "cr", "cb")
                       for a, b in zip(dec.as_dict[ch], comp.as_dict[ch]))
        if not same:
            raise RuntimeError("reference round trip is not the identity")
        rgb_out = compression.jpeg_decompression(dec) if mode == "jpeg" else compression.wavelet_decompression(dec)
        out["rgb_out"] = rgb_out
        out["decode_error"] = np.array("")
    except (AssertionError, RuntimeError, ValueError, IndexError) as e:
        # (level counts the reference's decoder mis-reads -- 1 and 4, codec.py:182-189 -- end here)
        out["decode_error"] = np.array(type(e).__name__)
    print("  %s: decode %.1fs (%s)" % (name, time.time() - t0, str(out["decode_error"]) or "ok"), flush=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*")
    args = ap.parse_args()
    settings = refshim.install()
    settings.DEBUG = False
    cases = _cases()
    for name, case in cases.items():
        if args.only and name not in args.only:
            continue
        mode, rgb = case[0], case[1]
        print(name, mode, rgb.shape, flush=True)
        run_case(name, mode, rgb, case[2] if len(case) > 2 else None)


if __name__ == "__main__":
    main()

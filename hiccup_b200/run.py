"""File-level driver, drop-in for reference hiccup/run.py:13-43 (the caller of the hot path; SURVEY 8(f)
rank 1) plus a directory mode that feeds the batched GPU path.

    compress(path, output, style)      run.py:18-29   image file -> <output>/<name>.<STYLE>-hic
    decompress(path)                   run.py:32-43   .hic file -> pixels (the reference shows them in a window;
                                                       here they are returned, and written with `save=`)
    compress_many(paths, output, ...)  additive: same-shaped images are encoded as one GPU batch and their files
                                       written by the library's host threads (batch.write_files)
    decompress_many(paths, ...)        additive: same-shaped `.hic` files are read and decoded as one GPU batch

Like the reference, `cv2.imread` delivers BGR and the codec treats channel 0 as R (run.py:19), so files
are interchangeable with the reference's in both directions.  Command line: python -m hiccup_b200.run
-c IMG... | -d HIC... [-s JPEG|HIC] [-o OUT] [-r], the subset of bin/belch.py:26-49 that does not need a display.
"""
import argparse
import os

from hiccup_b200 import codec, compression, hicimage, model


def img_name(path, c):
    return os.path.split(path)[-1] + ".%s-hic" % c.value


def _read(path):
    import cv2
    rgb = cv2.imread(path)
    if rgb is None:
        raise RuntimeError("cannot read image %r" % (path,))
    return rgb


def compress(path, output, c=model.Compression.JPEG, restarts=False):
    """restarts=True appends restart records to the file (an extension entry the reference's reader ignores;
    this package's decoder then skips the synchronisation passes of the Huffman decode)."""
    rgb = _read(path)
    if c == model.Compression.JPEG:
        hi = codec.jpeg_encode(compression.jpeg_compression(rgb), restarts=restarts)
    elif c == model.Compression.HIC:
        hi = codec.wavelet_encode(compression.wavelet_compression(rgb), restarts=restarts)
    else:
        raise RuntimeError("Unknown compression type")
    out = os.path.join(output, img_name(path, c))
    hi.write_file(out)
    return out


def decompress(path, save=None):
    hi = hicimage.HicImage.from_file(path)
    if hi.hic_type == model.Compression.JPEG:
        rgb = compression.jpeg_decompression(codec.jpeg_decode(hi))
    elif hi.hic_type == model.Compression.HIC:
        rgb = compression.wavelet_decompression(codec.wavelet_decode(hi))
    else:
        raise RuntimeError("Unknown compression type")
    if save is not None:
        import cv2
        cv2.imwrite(save, rgb)
    return rgb


def compress_many(paths, output, c=model.Compression.JPEG, max_batch=256, restarts=False):
    """Encode many files; images of one shape go through the batched codec together (one launch of
    every kernel per group) and every `.hic` file equals what compress() writes for it."""
    import numpy as np
    from hiccup_b200.batch import DctBatchCodec, WaveletBatchCodec
    groups = {}
    for p in paths:
        rgb = _read(p)
        groups.setdefault(rgb.shape, []).append((p, rgb))
    written = []
    for shape, items in groups.items():
        h, w = shape[:2]
        for a in range(0, len(items), max_batch):
            part = items[a:a + max_batch]
            cls = DctBatchCodec if c == model.Compression.JPEG else WaveletBatchCodec
            bc = cls(len(part), h, w)
            try:
                enc = bc.encode(np.stack([rgb for _, rgb in part]))
                outs = [os.path.join(output, img_name(p, c)) for p, _ in part]
                if restarts:
                    for out, hi in zip(outs, bc.hic_images(enc)):
                        codec.add_restart_records(hi).write_file(out)
                else:
                    bc.write_files(enc, outs)
                written += outs
            finally:
                bc.close()
    return written


def _peek(raw):
    """(mode, (h, w) of the image) of a `.hic` file's bytes, without touching its tables."""
    entries = hicimage.loads(raw)
    mode = model.Compression(hicimage.PlainStringP.from_bytes(entries[0]).payload)
    if mode == model.Compression.JPEG:
        shape = tuple(int(v) for v in hicimage.TupP.from_bytes(entries[19]).numbers)
    else:
        big = tuple(int(v) for v in hicimage.TupP.from_bytes(entries[14]).numbers)      # cD_1: half the image each way
        shape = (2 * big[0], 2 * big[1])
    return mode, shape


def decompress_many(paths, save_dir=None, max_batch=256):
    """Decode many `.hic` files; files of one mode and image shape go through the batched codec together.  Returns the
    pixel arrays in the order of `paths` (equal to what decompress() returns for each); save_dir: also write them as PNG."""
    from hiccup_b200.batch import DctBatchCodec, WaveletBatchCodec
    groups, out = {}, [None] * len(paths)
    for i, p in enumerate(paths):
        with open(p, "rb") as f:
            raw = f.read()
        groups.setdefault(_peek(raw), []).append((i, raw))
    for (mode, (h, w)), items in groups.items():
        cls = DctBatchCodec if mode == model.Compression.JPEG else WaveletBatchCodec
        for a in range(0, len(items), max_batch):
            part = items[a:a + max_batch]
            bc = cls(len(part), h, w)
            try:
                pixels = bc.decode(bc.streams_from_files([raw for _, raw in part]))
                for (i, _), rgb in zip(part, pixels):
                    out[i] = rgb.copy()
            finally:
                bc.close()
    if save_dir is not None:
        import cv2
        for p, rgb in zip(paths, out):
            cv2.imwrite(os.path.join(save_dir, os.path.split(p)[-1] + ".png"), rgb)
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description="hiccup image compression on the GPU (no GUI)")
    ap.add_argument("--compress", "-c", metavar="IMG_PATH", nargs="+")
    ap.add_argument("--decompress", "-d", metavar="HIC", nargs="+")
    ap.add_argument("--compression", "-s", metavar="STYLE", default=model.Compression.HIC.value,
                    choices=[model.Compression.HIC.value, model.Compression.JPEG.value])
    ap.add_argument("--output", "-o", metavar="OUT", default=".")
    ap.add_argument("--restarts", "-r", action="store_true",
                    help="append restart records to the .hic files (ignored by the reference's reader; faster decode here)")
    args = ap.parse_args(argv)
    style = model.Compression(args.compression)
    if args.compress:
        for out in (compress_many(args.compress, args.output, style, restarts=args.restarts) if len(args.compress) > 1
                    else [compress(args.compress[0], args.output, style, restarts=args.restarts)]):
            print(out)
    elif args.decompress:
        if len(args.decompress) > 1:
            decompress_many(args.decompress, save_dir=args.output)
        else:
            decompress(args.decompress[0], save=os.path.join(args.output, os.path.split(args.decompress[0])[-1] + ".png"))
        for p in args.decompress:
            print(os.path.join(args.output, os.path.split(p)[-1] + ".png"))
    else:
        raise RuntimeError("Illegal state")


if __name__ == "__main__":
    main()

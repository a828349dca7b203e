"""Host cost of the `.hic` container step for one 640x426 image, no GPU involved: code tables and framed bit strings of
the sizes a C2 image has (synthetic rows: 6 000 of them, codes of 2..24 bits, 150 KB of bits) -> HicImage ->
byte_stream() (what write_file dumps) and back through HicImage.from_bytes.
Prints one JSON line; `paths` says which of the native whole-table / native row / pure pickle paths calibrated here.

    python tools/hic_container_cpu.py [--reps 50]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiccup_b200 import hicimage             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--force", choices=["table", "rows", "pickle"], default="table",
                    help="which writer / reader to time: the whole-table native calls, the per-row native calls, or pickle alone")
    args = ap.parse_args()
    h, w = 426, 640
    rng = np.random.default_rng(1)
    arrs, framed = [], []
    for k, rows in enumerate([1717, 837, 851, 797, 596, 576, 15, 15, 15]):     # DC x3, run-length values x3, zero counts x3
        sym = rng.permutation(np.arange(-(rows // 2), rows - rows // 2)).astype(np.int32)
        lens = rng.integers(2, 25, rows).astype(np.uint8)
        codes = rng.integers(0, 1 << 24, rows, dtype=np.uint64) & ((np.uint64(1) << lens.astype(np.uint64)) - np.uint64(1))
        arrs.append((sym, lens, codes))
        framed.append(bytes([3]) + rng.integers(0, 256, [9000, 4000, 4000, 45000, 20000, 20000, 25000, 12000, 12000][k], dtype=np.uint8).tobytes())
    nat = hicimage._native()
    paths = {"rows": bool(nat.ok), "table": bool(nat.table_ok)}
    if args.force != "table":
        nat.table_ok = False
    if args.force == "pickle":
        nat.ok = False
        hicimage.ROWS.ok = False

    def write():
        tables = [hicimage.PayloadStringP.from_arrays(*arrs[k], k < 3) for k in range(9)]
        if args.force == "pickle":
            tables = [hicimage.PayloadStringP.from_rows(t.rows) for t in tables]
        bits = [hicimage.BitStringP.from_framed(f) for f in framed]
        hi = hicimage.HicImage.jpeg_image(tables + bits + [hicimage.TupP(h, w), hicimage.TupP(h // 2, w // 2)])
        return hi.byte_stream()

    stream = write()

    def read():
        hi = hicimage.HicImage.from_bytes(stream)
        return [p.arrays() for p in hi.payloads[:9]]

    read()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        write()
    t_write = (time.perf_counter() - t0) / args.reps
    t0 = time.perf_counter()
    for _ in range(args.reps):
        read()
    t_read = (time.perf_counter() - t0) / args.reps
    print(json.dumps({"image": "%dx%d synthetic" % (w, h), "table_rows": [int(a[0].size) for a in arrs],
                      "hic_bytes": sum(len(b) for b in stream), "path": args.force, "calibrated": paths,
                      "write_ms_per_image": round(t_write * 1e3, 3), "read_ms_per_image": round(t_read * 1e3, 3),
                      "cores": 1, "where": "host only (no GPU)"}))


if __name__ == "__main__":
    main()

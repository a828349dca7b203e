"""Host cost of the `.hic` container step for one 640x426 image, no GPU involved: the code tables and framed bit strings
of the oracle's encode -> HicImage -> byte_stream() (what write_file dumps) and back through HicImage.from_bytes.
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
from oracle import hiccup_oracle as orc      # noqa: E402  (input generator only)
from hiccup_b200 import hicimage             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--force", choices=["table", "rows", "pickle"], default="table",
                    help="which writer / reader to time: the whole-table native calls, the per-row native calls, or pickle alone")
    args = ap.parse_args()
    h, w = 426, 640
    enc = orc.jpeg_encode(orc.jpeg_compression(orc.synthetic_image(h, w, 5)))
    arrs = []
    for rows in enc["tables"]:
        arrs.append((np.array([int(a) for a, _ in rows], np.int32), np.array([len(b) for _, b in rows], np.uint8),
                     np.array([int(b, 2) for _, b in rows], np.uint64)))
    framed = [orc.padded_bits_to_bytes(b) for b in enc["bits"]]
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

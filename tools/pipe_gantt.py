"""Host-side phase timeline of PipelinedCodec.round_trip, using only its own synchronisation points (development aid)."""
import sys, os, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from hiccup_b200 import _lib
from hiccup_b200.batch import PipelinedCodec

def main():
    n, h, w = 1024, 426, 640
    chunk, slots = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 8)
    _lib.require_device()
    host, k1 = bench.pinned_array(_lib, (n, h, w, 3))
    base = bench.synthetic_batch(32, h, w, 2000)
    for i in range(n):
        host[i] = base[i % 32]
    out, k2 = bench.pinned_array(_lib, (n, 2 * (h // 2), 2 * (w // 2), 3))
    pipe = PipelinedCodec(n, h, w, chunk=chunk, slots=slots)
    for _ in range(2):
        pipe.round_trip(host, out)
    tr = []
    reps = 3
    t = time.perf_counter()
    pipe.round_trip(host, out, repeat=reps, trace=tr)
    total = (time.perf_counter() - t) * 1e3
    tr.sort(key=lambda e: e[2])
    names = ["wait_in", "h2d", "encode", "decode", "wait_out", "d2h"]
    agg = dict.fromkeys(names, 0.0)
    print("total %.1f ms for %d batches (%.1f ms each)" % (total, reps, total / reps))
    for e in tr:
        slot, v = e[0], e[1]
        ts = [(x - t) * 1e3 for x in e[2:]]
        d = [ts[i + 1] - ts[i] for i in range(6)]
        for k, x in zip(names, d):
            agg[k] += x
        print("slot %d visit %2d  start %6.1f | " % (slot, v, ts[0]) + "  ".join("%s %5.2f" % (k, x) for k, x in zip(names, d)) + " | end %6.1f" % ts[6])
    print("sums (ms):", {k: round(x, 1) for k, x in agg.items()})

if __name__ == "__main__":
    main()

"""Device-side timeline of PipelinedCodec.round_trip: which kernels ran when, on which stream (development aid)."""
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
    _lib.profile_enable(True)
    _lib.profile_report()
    t = time.perf_counter()
    pipe.round_trip(host, out, repeat=2)
    total = (time.perf_counter() - t) * 1e3
    tl = _lib.profile_timeline()
    _lib.profile_enable(False)
    print("wall %.1f ms for 2 batches; %d spans" % (total, len(tl)))
    # union of busy intervals
    iv = sorted((a, b) for _, _, a, b in tl)
    busy, cur_a, cur_b = 0.0, None, None
    for a, b in iv:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                busy += cur_b - cur_a
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    if cur_b is not None:
        busy += cur_b - cur_a
    print("span of spans %.1f ms, union of kernel intervals %.1f ms, sum of kernel intervals %.1f ms" % (
        max(b for _, _, a, b in tl) - min(a for _, _, a, b in tl), busy, sum(b - a for _, _, a, b in tl)))
    agg = {}
    for name, sid, a, b in tl:
        e = agg.setdefault(name, [0.0, 0])
        e[0] += b - a
        e[1] += 1
    for name, (ms, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("  %-28s total %8.2f ms over %4d launches (%.3f each)" % (name, ms, cnt, ms / cnt))
    if len(sys.argv) > 3:
        for name, sid, a, b in sorted(tl, key=lambda e: e[2]):
            if a < float(sys.argv[3]):
                print("%8.3f %8.3f  s%-2d %s" % (a, b, sid, name))

if __name__ == "__main__":
    main()

"""Does the interpreter's thread switch interval matter to PipelinedCodec.round_trip? (development aid)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from hiccup_b200 import _lib
from hiccup_b200.batch import PipelinedCodec

def main():
    n, h, w = 1024, 426, 640
    _lib.require_device()
    host, k1 = bench.pinned_array(_lib, (n, h, w, 3))
    base = bench.synthetic_batch(32, h, w, 2000)
    for i in range(n):
        host[i] = base[i % 32]
    out, k2 = bench.pinned_array(_lib, (n, 2 * (h // 2), 2 * (w // 2), 3))
    for chunk, slots in [(128, 8), (64, 8), (64, 16)]:
        pipe = PipelinedCodec(n, h, w, chunk=chunk, slots=slots)
        for _ in range(2):
            pipe.round_trip(host, out)
        for si in (0.005, 0.0005, 0.00005):
            sys.setswitchinterval(si)
            _lib.sync()
            reps = 5
            t = time.perf_counter()
            pipe.round_trip(host, out, repeat=reps)
            dts = (time.perf_counter() - t) / reps
            print("chunk %3d slots %2d switch %.5f: streamed %.1f ms -> %.0f MP/s" % (chunk, slots, si, dts * 1e3, n * h * w / 1e6 / dts), flush=True)
        pipe.close()

if __name__ == "__main__":
    main()

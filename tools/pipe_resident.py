"""PipelinedCodec with and without the bulk copies (development aid)."""
import sys, os, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
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
    for chunk, slots in [(128, 8), (128, 4), (256, 4), (256, 2), (512, 2), (1024, 1), (64, 16)]:
        pipe = PipelinedCodec(n, h, w, chunk=chunk, slots=slots)
        for _ in range(2):
            pipe.round_trip(host, out)
        res = []
        for resident in (False, True):
            _lib.sync()
            reps = 4
            t = time.perf_counter()
            pipe.round_trip(host, out, repeat=reps, resident=resident)
            res.append((time.perf_counter() - t) / reps * 1e3)
        print("chunk %4d slots %2d: host-to-host %.1f ms, resident %.1f ms per batch" % (chunk, slots, res[0], res[1]), flush=True)
        pipe.close()

if __name__ == "__main__":
    main()

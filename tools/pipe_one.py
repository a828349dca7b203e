"""One configuration of PipelinedCodec.round_trip, streamed: ms per 1024-image batch (development aid)."""
import sys, os, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from hiccup_b200 import _lib
from hiccup_b200.batch import PipelinedCodec
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
_lib.sync()
reps = 6
t = time.perf_counter()
pipe.round_trip(host, out, repeat=reps)
dt = (time.perf_counter() - t) / reps * 1e3
print("chunk %d slots %d: %.1f ms per batch" % (chunk, slots, dt), flush=True)

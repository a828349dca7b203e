"""Device-resident C2 step through PipelinedCodec.device_steps for several chunkings (development aid):
how much of the latency-bound stretches of one chunk hides under another chunk's kernels?"""
import sys, os, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from hiccup_b200 import _lib
from hiccup_b200.batch import PipelinedCodec, DctBatchCodec

def main():
    n, h, w = 1024, 426, 640
    _lib.require_device()
    base = bench.synthetic_batch(64, h, w, 2000)
    rgb = np.concatenate([base] * (n // 64))
    reps = 8
    codec = DctBatchCodec(n, h, w)
    codec.upload(rgb)
    for _ in range(3):
        codec.encode_device(); codec.decode_device()
    _lib.sync()
    t = time.perf_counter()
    for _ in range(reps):
        codec.encode_device(); codec.decode_device()
    _lib.sync()
    print("one codec, one stream: %.2f ms per step" % ((time.perf_counter() - t) * 1e3 / reps), flush=True)
    codec.close()
    for slots in (4, 8, 16, 32):
        if n % slots:
            continue
        pipe = PipelinedCodec(n, h, w, chunk=n // slots, slots=slots)
        pipe.upload_resident(rgb)
        pipe.device_steps(3)
        _lib.sync()
        t = time.perf_counter()
        pipe.device_steps(reps)
        _lib.sync()
        print("%d slots x %d images: %.2f ms per step" % (slots, n // slots, (time.perf_counter() - t) * 1e3 / reps), flush=True)
        pipe.close()

if __name__ == "__main__":
    main()

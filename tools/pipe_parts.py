"""Pipelined path with parts switched off, to find what bounds it (development aid)."""
import sys, os, time, threading
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from hiccup_b200 import _lib
from hiccup_b200.batch import PipelinedCodec

def run(pipe, host, out, copies, kernels, reps=4):
    chunk = pipe.chunk
    def work(slot):
        codec = pipe.codecs[slot]
        for c in range(slot, pipe.n_chunks, pipe.slots):
            a, b = c * chunk, (c + 1) * chunk
            if copies:
                codec.upload(host[a:b])
            if kernels:
                codec.encode_device()
                codec.decode_device()
            if copies:
                cnt = chunk * codec.out_h * codec.out_w * 3
                codec.d_out.download(np.uint8, cnt, codec.stream, out=out[a:b].reshape(-1))
            _lib.sync(codec.stream)
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        th = [threading.Thread(target=work, args=(s,)) for s in range(pipe.slots)]
        for x in th: x.start()
        for x in th: x.join()
        best = min(best, time.perf_counter() - t)
    return best * 1e3

def main():
    n, h, w = 1024, 426, 640
    _lib.require_device()
    host, k1 = bench.pinned_array(_lib, (n, h, w, 3))
    base = bench.synthetic_batch(32, h, w, 2000)
    for i in range(n):
        host[i] = base[i % 32]
    out, k2 = bench.pinned_array(_lib, (n, 2 * (h // 2), 2 * (w // 2), 3))
    for chunk, slots in [(128, 3), (128, 6), (64, 8)]:
        pipe = PipelinedCodec(n, h, w, chunk=chunk, slots=slots)
        for _ in range(2):
            pipe.round_trip(host, out)
        print("chunk %d slots %d: copies only %.1f ms | kernels only %.1f ms | both %.1f ms" % (
            chunk, slots, run(pipe, host, out, True, False), run(pipe, host, out, False, True), run(pipe, host, out, True, True)), flush=True)
        pipe.close()

if __name__ == "__main__":
    main()

"""Where one Lenna-sized call through the drop-in functions spends its time (development aid)."""
import cProfile, pstats, pickle, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hiccup_b200 import codec, compression, hicimage

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lenna512.npz"))
rgb = np.ascontiguousarray(g["rgb"])

def call():
    comp = compression.jpeg_compression(rgb)
    hic = codec.jpeg_encode(comp)
    back = hicimage.HicImage.from_bytes(hic.byte_stream())
    return compression.jpeg_decompression(codec.jpeg_decode(back))

for _ in range(3):
    call()
t = time.perf_counter(); call(); print("one call: %.1f ms" % ((time.perf_counter() - t) * 1e3))
pr = cProfile.Profile(); pr.enable(); call(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)

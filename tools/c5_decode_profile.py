"""Where the decode of one 16384^2 image spends its time (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, bench_bands
from hiccup_b200 import _lib, bands
from hiccup_b200.batch import DctBatchCodec

size = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
_lib.require_device()
image = bench_bands.big_synthetic(size)
cuts = bands.plan_bands(size, 1)
worker = bands.BandWorker(0, 1, size, size, cuts[0], cuts[1], device=0)
worker.load(image)
res = bands.run_local([worker])[0]
t = time.perf_counter(); res = bands.run_local([worker])[0]; print("encode (1 band) %.1f ms" % ((time.perf_counter() - t) * 1e3))
dec = DctBatchCodec(1, size, size)
staging = _lib.PinnedBuffer(bands.stitch_layout(res["all_bits"], 9)[3] + (1 << 20))
for rep in range(3):
    t0 = time.perf_counter()
    enc = bands.to_encoded_streams(res, size, size, out=staging.array(np.uint8))
    t1 = time.perf_counter()
    if rep == 2:
        _lib.profile_enable(True); _lib.profile_report()
    dec.decode_resident(enc)
    _lib.sync()
    t2 = time.perf_counter()
    out = dec.fetch()
    t3 = time.perf_counter()
    print("rep %d: stitch %.1f ms, decode_resident %.1f ms, fetch %.1f ms" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
for k, (ms, n) in sorted(_lib.profile_report().items(), key=lambda kv: -kv[1][0]):
    print("  %-28s %9.3f ms x %d" % (k, ms / max(n, 1), n))
print("nbits per stream:", [int(x) for x in enc.nbits])

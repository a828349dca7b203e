"""Device-resident C2-style step: ms per step (profiling off), then the per-kernel table (development aid).

    python tools/step_time.py [n h w [reps]]      # env: HIC_REPLAY_WIDE, HIC_REPLAY_LOOK, HIC_ENTROPY_SERIAL ...
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from hiccup_b200 import _lib
from hiccup_b200.batch import DctBatchCodec

def main():
    n, h, w = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (1024, 426, 640)))
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    _lib.require_device()
    codec = DctBatchCodec(n, h, w)
    base = bench.synthetic_batch(min(n, 64), h, w, 2000)
    rgb = np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]
    codec.upload(rgb)
    def step():
        codec.encode_device()
        codec.decode_device()
    for _ in range(3):
        step()
    _lib.sync()
    t = time.perf_counter()
    for _ in range(reps):
        step()
    _lib.sync()
    ms = (time.perf_counter() - t) * 1e3 / reps
    t = time.perf_counter()
    for _ in range(reps):
        codec.encode_device()
    _lib.sync()
    ms_enc = (time.perf_counter() - t) * 1e3 / reps
    tag = " ".join("%s=%s" % (k, os.environ[k]) for k in sorted(os.environ) if k.startswith("HIC_"))
    print("STEP %.3f ms  (encode alone %.3f ms)  [%s]  payload %d bytes" % (ms, ms_enc, tag, int(codec.encoder.total_bytes)), flush=True)
    if "--profile" in sys.argv:
        os.environ["HIC_ENTROPY_SERIAL"] = "1"
        step()
        _lib.sync()
        _lib.profile_enable(True)
        _lib.profile_report()
        for _ in range(2):
            step()
        _lib.sync()
        for k, (ms_k, launches) in sorted(_lib.profile_report().items(), key=lambda kv: -kv[1][0]):
            print("  %-28s %8.4f ms x %d" % (k, ms_k / max(launches, 1), launches // 2))

if __name__ == "__main__":
    main()

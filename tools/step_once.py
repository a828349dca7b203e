"""One device-resident encode+decode step of a batch (development aid: the ncu target).

    python tools/step_once.py [n h w [reps]]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from hiccup_b200 import _lib
from hiccup_b200.batch import DctBatchCodec

def main():
    n, h, w = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (1024, 426, 640)))
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    _lib.require_device()
    codec = DctBatchCodec(n, h, w)
    base = bench.synthetic_batch(min(n, 64), h, w, 2000)
    rgb = np.concatenate([base] * ((n + len(base) - 1) // len(base)))[:n]
    codec.upload(rgb)
    _lib.profile_enable(True)
    for _ in range(reps):
        codec.encode_device()
        codec.decode_device()
    _lib.sync()
    rows = np.asarray(codec.encoder.rows).reshape(n, 3, 3)       # [image, channel, kind]
    print("alphabet sizes (mean / max over images), channels lum cr cb x kinds dc value length:")
    print(np.round(rows.mean(axis=0)).astype(int).tolist(), rows.max(axis=0).tolist())
    for k, (ms, launches) in sorted(_lib.profile_report().items(), key=lambda kv: -kv[1][0]):
        print("%-28s %8.4f ms x %d" % (k, ms / max(launches, 1), launches))

if __name__ == "__main__":
    main()

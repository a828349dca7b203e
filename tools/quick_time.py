"""Quick device-resident timing of the transform kernels (development aid; bench.py is the contract)."""
import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hiccup_b200 import _lib

def main():
    n, h, w = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (1024, 426, 640)))
    lib = _lib.load(); _lib.require_device()
    g = _lib.geometry(h, w)
    dev = torch.device("cuda:0")
    import bench
    base = bench.synthetic_batch(min(n, 16), h, w, 2000)
    reps = (n + len(base) - 1) // len(base)
    rgb = torch.from_numpy(np.concatenate([base] * reps)[:n].copy()).to(dev).contiguous()
    blocks = n * g.blocks_per_image
    coef = torch.empty(blocks * 64, dtype=torch.int16, device=dev)
    cap = _lib.tie_capacity(n, h, w)
    ties = torch.empty(cap * 16, dtype=torch.uint8, device=dev)
    stats = torch.zeros(4, dtype=torch.int32, device=dev)
    yp = torch.empty(n * h * w, dtype=torch.uint8, device=dev)
    crp = torch.empty(n * g.hc * g.wc, dtype=torch.uint8, device=dev)
    cbp = torch.empty(n * g.hc * g.wc, dtype=torch.uint8, device=dev)
    out = torch.empty(n * g.out_h * g.out_w * 3, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def fwd():
        _lib.check(lib.hic_dct_forward(rgb.data_ptr(), n, h, w, coef.data_ptr(), ties.data_ptr(), cap, stats.data_ptr(), st))
    def inv():
        _lib.check(lib.hic_dct_inverse(coef.data_ptr(), n, h, w, yp.data_ptr(), crp.data_ptr(), cbp.data_ptr(), out.data_ptr(), ties.data_ptr(), blocks, stats.data_ptr(), st))
    for name, fn, bpp in (("forward", fwd, 6.0), ("inverse", inv, 6.0)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        mp = n * h * w / 1e6
        print("%s: %.3f ms  %.1f MP/s  %.1f GB/s algorithmic (%.1f%% of 6548.8)  stats=%s" % (
            name, ms, mp / ms * 1e3, mp * 1e6 * bpp / ms / 1e6, mp * bpp / ms / 6548.8 * 100 * 1e0, stats.tolist()))
main()

"""Where the host-to-host batch path spends its time (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from hiccup_b200 import _lib, entropy
from hiccup_b200.batch import DctBatchCodec

def main():
    n, h, w = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (1024, 426, 640)))
    _lib.require_device()
    codec = DctBatchCodec(n, h, w)
    host, keep = bench.pinned_array(_lib, (n, h, w, 3))
    base = bench.synthetic_batch(min(n, 32), h, w, 2000)
    for i in range(n):
        host[i] = base[i % len(base)]
    T = {}
    def tick(name, t0):
        _lib.sync()
        T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    reps = 3
    for rep in range(reps + 1):
        if rep == 1:
            T.clear()
        t = time.perf_counter(); codec.upload(host); tick("enc.h2d_rgb", t)
        t = time.perf_counter(); out = codec.encode_device(); tick("enc.device", t)
        enc = codec.encoder
        nbytes = int(enc.total_bytes)
        t = time.perf_counter()
        if codec._h_data is None or codec._h_data.nbytes < nbytes:
            codec._h_data = _lib.PinnedBuffer(nbytes + nbytes // 4)
        data = out.download(np.uint8, nbytes, None, out=codec._h_data.array(np.uint8, nbytes)); tick("enc.d2h_bits", t)
        t = time.perf_counter()
        rows = int(enc.total_rows)
        if codec._h_tab is None or codec._h_tab[1].size < rows:
            cap = rows + rows // 4 + 1024
            codec._h_tab_mem = [_lib.PinnedBuffer(8 * enc.n_streams), _lib.PinnedBuffer(4 * cap), _lib.PinnedBuffer(8 * cap)]
            codec._h_tab = (codec._h_tab_mem[0].array(np.uint32), codec._h_tab_mem[1].array(np.int32), codec._h_tab_mem[2].array(np.uint64))
        index, sym, packed = enc.tables_packed(None, out=codec._h_tab); tick("enc.tables", t)
        res = entropy.EncodedStreams(codec.layout, index, enc.nsym.copy(), enc.nbits.copy(), enc.byte_off.copy(),
                                     enc.byte_len.copy(), sym, packed, data)
        t = time.perf_counter()
        dec = codec.decoder
        idx = np.ascontiguousarray(res.index, np.uint32)
        _lib.check(dec.lib.hic_decode_set_tables_packed(dec.plan, idx.ctypes.data, res.symbols.ctypes.data, res.packed.ctypes.data,
                                                        int(res.symbols.size), None)); tick("dec.set_tables", t)
        t = time.perf_counter()
        need = data.nbytes + 16
        if dec._in is None or dec._in.nbytes < need:
            dec._in = _lib.DeviceBuffer(need + need // 4)
        dec._in.upload(data); tick("dec.h2d_bits", t)
        t = time.perf_counter(); dec.run(dec._in.ptr, res.byte_off, res.nbits, codec.d_coef_dec.ptr); tick("dec.entropy", t)
        t = time.perf_counter()
        _lib.check(codec.lib.hic_dct_inverse(codec.d_coef_dec.ptr, n, h, w, codec.d_y.ptr, codec.d_cr.ptr, codec.d_cb.ptr,
                                             codec.d_out.ptr, codec.d_ties.ptr, codec.blocks, codec.d_stats.ptr, None)); tick("dec.inverse", t)
        t = time.perf_counter()
        count = n * codec.out_h * codec.out_w * 3
        if codec._h_out is None:
            codec._h_out = _lib.PinnedBuffer(count)
        codec.d_out.download(np.uint8, count, None, out=codec._h_out.array(np.uint8, count)); tick("dec.d2h_rgb", t)
    tot = sum(T.values()) / reps
    for k, v in T.items():
        print("%-16s %8.2f ms" % (k, v / reps))
    print("total %.2f ms -> %.1f MP/s; bits %.1f MB, rows %d" % (tot, n * h * w / 1e6 / (tot / 1e3), nbytes / 1e6, enc.total_rows))

if __name__ == "__main__":
    main()

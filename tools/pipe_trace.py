"""Phase timeline of the pipelined host-to-host path (development aid)."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from hiccup_b200 import _lib, entropy
from hiccup_b200.batch import PipelinedCodec

def main():
    n, h, w = 1024, 426, 640
    chunk, slots = int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 3
    _lib.require_device()
    host, k1 = bench.pinned_array(_lib, (n, h, w, 3))
    base = bench.synthetic_batch(32, h, w, 2000)
    for i in range(n):
        host[i] = base[i % 32]
    out, k2 = bench.pinned_array(_lib, (n, 2 * (h // 2), 2 * (w // 2), 3))
    pipe = PipelinedCodec(n, h, w, chunk=chunk, slots=slots)
    for _ in range(2):
        pipe.round_trip(host, out)
    ev = []
    t0 = time.perf_counter()
    def work(slot):
        codec = pipe.codecs[slot]
        st = codec.stream
        def mark(c, name, ta):
            _lib.sync(st)
            ev.append((slot, c, name, (ta - t0) * 1e3, (time.perf_counter() - t0) * 1e3))
        for c in range(slot, pipe.n_chunks, pipe.slots):
            a, b = c * chunk, (c + 1) * chunk
            t = time.perf_counter(); codec.upload(host[a:b]); mark(c, "h2d", t)
            t = time.perf_counter(); codec._forward(); mark(c, "fwd", t)
            t = time.perf_counter(); codec.encoder.symbolize(codec.d_coef.ptr, st); mark(c, "sym", t)
            t = time.perf_counter(); codec.encoder.build_codes(st, on_device=True); mark(c, "build", t)
            t = time.perf_counter(); o = codec.encoder.pack(st); mark(c, "pack", t)
            t = time.perf_counter()
            enc = codec.encoder
            nbytes = int(enc.total_bytes)
            data = o.download(np.uint8, nbytes, st, out=codec._h_data.array(np.uint8, nbytes))
            index, sym, packed = enc.tables_packed(st, out=codec._h_tab)
            e = entropy.EncodedStreams(codec.layout, index, enc.nsym.copy(), enc.nbits.copy(), enc.byte_off.copy(), enc.byte_len.copy(), sym, packed, data)
            mark(c, "d2h_bits", t)
            t = time.perf_counter(); codec.decoder.decode_streams(e, codec.d_coef_dec.ptr, st); mark(c, "dec", t)
            t = time.perf_counter(); codec._inverse(); mark(c, "inv", t)
            t = time.perf_counter()
            cnt = chunk * codec.out_h * codec.out_w * 3
            codec.d_out.download(np.uint8, cnt, st, out=out[a:b].reshape(-1)); mark(c, "d2h_out", t)
    th = [threading.Thread(target=work, args=(s,)) for s in range(slots)]
    for t in th: t.start()
    for t in th: t.join()
    total = (time.perf_counter() - t0) * 1e3
    ev.sort(key=lambda e: e[3])
    for slot, c, name, a, b in ev:
        print("%7.2f %7.2f  slot %d chunk %2d  %-9s %6.2f ms" % (a, b, slot, c, name, b - a))
    agg = {}
    for slot, c, name, a, b in ev:
        agg[name] = agg.get(name, 0) + (b - a)
    print("total %.1f ms; per-phase sums:" % total, {k: round(v, 1) for k, v in agg.items()})

if __name__ == "__main__":
    main()

"""A band worker built on the ORACLE (CPU), standing in for the GPU BandWorker so that the host-side
stitching logic of hiccup_b200/bands.py can be tested without a GPU, in one process and across
torch.distributed (gloo) ranks.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

from hiccup_b200 import bands
from oracle import hiccup_oracle as orc


def _symbols_with_seam(ac, carry, more_after, closes):
    """What the device emits for one band: its slice of the whole channel's run-length list."""
    if not more_after and not np.any(ac != 0):
        # nothing but zeros from here to the end of the stream: the closing band owns the single (0, 0)
        return (np.zeros(1, np.int64), np.zeros(1, np.int64)) if closes else (np.zeros(0, np.int64), np.zeros(0, np.int64))
    virt = np.concatenate([np.zeros(carry, ac.dtype), ac, np.ones(1 if more_after else 0, ac.dtype)])
    lengths, values = orc.run_length(virt)
    lengths, values = list(lengths), list(values)
    drop_front = carry // 15                      # fillers inside the carried run belong to the bands above
    if more_after:
        lengths, values = lengths[:-1], values[:-1]          # the sentinel's own symbol
    elif not closes and lengths and lengths[-1] == 0 and values[-1] == 0:
        lengths, values = lengths[:-1], values[:-1]          # trailing zeros: the closing band says (0, 0)
    return np.array(lengths[drop_front:], np.int64), np.array(values[drop_front:], np.int64)


class OracleBandWorker:
    def __init__(self, band, n_bands, planes, r0, r1):
        """planes: the whole image's oracle coefficient planes; the band owns luminance rows [r0, r1)."""
        self.band, self.n_bands = band, n_bands
        h = planes["lum"].shape[0]
        hc = planes["cr"].shape[0]
        c0, c1 = r0 // 2, min(r1, 2 * (h // 2)) // 2
        c1 = hc if r1 >= h else c1
        rows = {"lum": (r0, r1), "cr": (c0, c1), "cb": (c0, c1)}
        self.zz = {}
        for ch in orc.CHANNELS:
            a, b = rows[ch]
            self.zz[ch] = orc.blocks_zigzag(planes[ch][a:b])

    def steps(self):
        first, last, length, last_dc = [], [], [], []
        for ch in orc.CHANNELS:
            ac = self.zz[ch][:, 1:].reshape(-1)
            nz = np.flatnonzero(ac)
            first.append(int(nz[0]) if nz.size else -1)
            last.append(int(nz[-1]) if nz.size else -1)
            length.append(int(ac.size))
            last_dc.append(int(self.zz[ch][-1, 0]))
        edges = yield ("all_gather", dict(first_nz=first, last_nz=last, length=length, last_dc=last_dc))
        seam = bands.seam_state(edges, self.band)
        streams = []
        for c, ch in enumerate(orc.CHANNELS):
            carry, prev_dc, more_after, closes = seam[c]
            dc = self.zz[ch][:, 0].astype(np.int64)
            diffs = dc.copy()
            diffs[1:] = dc[1:] - dc[:-1]
            diffs[0] = dc[0] - prev_dc
            lengths, values = _symbols_with_seam(self.zz[ch][:, 1:].reshape(-1).astype(np.int64), carry, more_after, closes)
            streams += [diffs, values, lengths]
        hists, nsym = [], []
        for sy in streams:
            if sy.size:
                uniq, first_idx, counts = np.unique(sy, return_index=True, return_counts=True)
            else:
                uniq, first_idx, counts = np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int64)
            hists.append((uniq.astype(np.int32), counts.astype(np.uint32), first_idx.astype(np.uint32)))
            nsym.append(int(sy.size))
        gathered = yield ("all_gather", dict(hists=hists, nsym=nsym))
        tables = bands.build_tables(bands.merge_histograms([m["hists"] for m in gathered], [m["nsym"] for m in gathered]),
                                    bands._host_huffman)
        all_bits = [bands.band_bits(tables, m["hists"]) for m in gathered]
        start_bit, _ = bands.bit_layout(all_bits, self.band)
        index, syms, packed = tables
        mine = []
        for s, sy in enumerate(streams):
            a, n = int(index[s, 0]), int(index[s, 1])
            lut = {int(v): format(int(pk & bands.CODE_MASK), "0%db" % int(pk >> np.uint64(58)))
                   for v, pk in zip(syms[a:a + n].tolist(), packed[a:a + n])}
            bits = "0" * int(start_bit[s]) + "".join(lut[int(v)] for v in sy.tolist())
            assert len(bits) - int(start_bit[s]) == int(all_bits[self.band][s])
            bits += "0" * ((-len(bits)) % 8)
            mine.append(int(bits, 2).to_bytes(len(bits) // 8, "big") if bits else b"")
        final = yield ("gather", dict(bytes=mine))
        if final is None:
            return None
        return dict(tables=tables, all_bits=all_bits, band_bytes=[m["bytes"] for m in final])

    def close(self):
        pass

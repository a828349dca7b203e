"""Row-band sharding of one large image's DCT-mode encode (BASELINE config 5).

The image is cut into bands of rows (multiples of 16, so luminance and chroma blocks align); every
band runs the transform and entropy kernels on its own GPU, and only small metadata crosses between
them -- there is no data-path collective:

    1. every band: K1 on its rows (plus 16 rows of halo each side for the chroma pyramid), E1 pass 1
       -> first / last non-zero position per channel, last DC value per channel
    2. all-gather of those few integers; every band derives its seam state (zero run carried in, previous
       DC, whether a later band holds a non-zero, whether it closes the stream)
    3. every band: E1 pass 2 with the seam state -> exactly its slice of the whole image's symbol lists,
       and their histograms with first-occurrence indices
    4. all-gather of the histograms; every band merges them (counts add, first occurrences offset by the
       symbols of the bands above), replays the reference's heapq Huffman construction once (host) and
       gets the same nine code tables
    5. every band: bit-packs its symbols with the shared codes at the bit phase its slice starts at in the
       stitched string
    6. gather of the band strings; the root ORs them together at their byte offsets and adds the
       pad-count byte (iohelper.py:35-48)

The result is byte-identical to encoding the whole image on one GPU (and to the reference).  The host
logic below is written against a tiny communicator (`all_gather`, `gather`) so that it runs in one process
over several bands (LocalComm, bands executed one after another) or one process per GPU under
torch.distributed (DistComm; gloo or NCCL object collectives carry the metadata).
"""
import ctypes

import numpy as np

CHANNELS = 3
KIND_DC, KIND_VALUE, KIND_LENGTH = 0, 1, 2
CODE_MASK = np.uint64((1 << 58) - 1)


# ------------------------------------------------------------------------------------------------
# planning
# ------------------------------------------------------------------------------------------------
def plan_bands(h, n_bands, align=16):
    """Row boundaries [r_0 = 0, r_1, ..., r_K = h] with every interior boundary a multiple of `align`;
    bands may be empty only if the image has fewer than K aligned rows groups."""
    units = (h + align - 1) // align
    n_bands = max(1, min(int(n_bands), units))
    cuts = [min(h, align * ((units * b) // n_bands)) for b in range(n_bands)] + [h]
    if len(cuts) > 2 and h - cuts[-2] < 2:          # a one-row band has no chroma row of its own
        del cuts[-2]
    return cuts


def band_slice(h, r0, r1, halo=16):
    """Rows of the source image band [r0, r1) needs: 16 rows of halo each side (one chroma block row),
    so the band's blocks are whole block rows of the slice and K1 runs unchanged."""
    return max(0, r0 - halo), min(h, r1 + halo)


# ------------------------------------------------------------------------------------------------
# seam state, histogram merge, bit offsets: pure functions (tests/test_bands_host.py)
# ------------------------------------------------------------------------------------------------
def seam_state(edges, b):
    """edges: per band a dict(first_nz[3], last_nz[3], length[3], last_dc[3]).  Returns band b's
    [(carry_zeros, prev_dc, more_after, closes_stream)] per channel (codec.run_length_coding and
    differential_coding run over the WHOLE channel, codec.py:47-99)."""
    out = []
    for c in range(CHANNELS):
        carry = 0
        for e in edges[:b]:
            if e["last_nz"][c] >= 0:
                carry = e["length"][c] - 1 - e["last_nz"][c]
            else:
                carry += e["length"][c]
        prev_dc = edges[b - 1]["last_dc"][c] if b > 0 else 0
        more_after = any(e["last_nz"][c] >= 0 for e in edges[b + 1:])
        out.append((int(carry), int(prev_dc), int(more_after), int(b == len(edges) - 1)))
    return out


def merge_histograms(band_hists, band_nsym):
    """band_hists[b][s]: (symbols, counts, first_local) arrays of band b's symbol stream s (s = channel*3+kind);
    band_nsym[b][s]: symbols of that stream in band b.  Returns per stream (symbols, counts) ordered by
    first occurrence in the stitched stream (utils.group_by order, utils.py:83-96).
    Dense accumulators over the symbol range of the stream (symbols are int16): a handful of numpy calls per
    stream whatever the number of bands."""
    n_streams = len(band_hists[0])
    merged = []
    for s in range(n_streams):
        syms, cnts, firsts, base = [], [], [], 0
        for b, hists in enumerate(band_hists):
            sym, cnt, first = hists[s]
            if len(sym):
                syms.append(np.asarray(sym, np.int64))
                cnts.append(np.asarray(cnt, np.int64))
                firsts.append(np.asarray(first, np.int64) + base)
            base += int(band_nsym[b][s])
        if not syms:
            merged.append((np.zeros(0, np.int32), np.zeros(0, np.uint32)))
            continue
        sym, cnt, first = np.concatenate(syms), np.concatenate(cnts), np.concatenate(firsts)
        lo = int(sym.min())
        idx = sym - lo
        size = int(idx.max()) + 1
        total = np.bincount(idx, weights=cnt.astype(np.float64), minlength=size).astype(np.int64)   # counts < 2^53: exact
        gfirst = np.full(size, np.iinfo(np.int64).max, np.int64)
        np.minimum.at(gfirst, idx, first)
        present = np.flatnonzero(total)
        order = present[np.argsort(gfirst[present])]          # (first occurrences are distinct positions: no ties)
        merged.append(((order + lo).astype(np.int32), total[order].astype(np.uint32)))
    return merged


def build_tables(merged, huffman_build):
    """Codes of every stream from the merged histograms.  huffman_build(freqs uint32[n]) -> (lens uint8[n],
    codes uint64[n]) must replay the reference's heapq construction (hic_huffman_build_host).
    Returns the packed table layout: index (n_streams, 2), symbols int32, packed uint64."""
    index, syms, packed, pos = [], [], [], 0
    # the nine constructions are independent C calls (ctypes drops the GIL): a few threads take them side by side
    built = list(_pool().map(lambda sc: huffman_build(sc[1]) if sc[0].size else None, merged))
    for (sym, cnt), lc in zip(merged, built):
        if sym.size:
            lens, codes = lc
            syms.append(sym)
            packed.append((lens.astype(np.uint64) << np.uint64(58)) | (codes & CODE_MASK))
        index.append((pos, sym.size))
        pos += sym.size
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return np.array(index, np.uint32).reshape(-1, 2), cat(syms, np.int32), cat(packed, np.uint64)


_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=3)
    return _POOL


def band_bits(tables, hists):
    """Coded bits of one band's streams under the shared tables: sum of count * code length."""
    return all_band_bits(tables, [hists])[0]


def all_band_bits(tables, band_hists):
    """band_bits for every band at once: one dense symbol -> code length table per stream, shared by the bands."""
    index, syms, packed = tables
    out = [np.zeros(len(h), np.uint64) for h in band_hists]
    for s in range(len(band_hists[0])):
        a, n = int(index[s, 0]), int(index[s, 1])
        if not n:
            assert all(not len(h[s][0]) for h in band_hists), "band symbols but an empty merged table"
            continue
        tab_sym = syms[a:a + n].astype(np.int64)
        lo = int(tab_sym.min())
        lut = np.zeros(int(tab_sym.max()) - lo + 1, np.int64)
        lut[tab_sym - lo] = (packed[a:a + n] >> np.uint64(58)).astype(np.int64)
        for b, h in enumerate(band_hists):
            sym, cnt, _ = h[s]
            if not len(sym):
                continue
            i = np.asarray(sym, np.int64) - lo
            assert i.min() >= 0 and i.max() < lut.size and lut[i].all(), "band symbol missing from the merged table"
            out[b][s] = int((np.asarray(cnt, np.int64) * lut[i]).sum())
    return out


def bit_layout(all_bits, b):
    """all_bits[b'][s]: coded bits of stream s in band b'.  Band b's slice of stream s starts at bit
    8 + sum of the bands above (8 = the pad-count byte).  Returns (start_bit[s] in 0..7, first byte[s])."""
    above = np.zeros(len(all_bits[0]), np.uint64)
    for bb in all_bits[:b]:
        above += bb
    pos = above + np.uint64(8)
    return (pos % np.uint64(8)).astype(np.uint32), (pos // np.uint64(8)).astype(np.int64)


def stitch_layout(all_bits, n_streams):
    """Byte layout of the stitched payloads in one buffer: per stream (offset (4-byte aligned), framed
    length, payload bits), and the buffer size (with 16 readable bytes of slack for the decoder)."""
    off, length, nbits, pos = [], [], [], 0
    for s in range(n_streams):
        total = int(sum(int(bb[s]) for bb in all_bits))
        pad = 8 - (total % 8)
        n = 1 + (total + pad) // 8
        off.append(pos)
        length.append(n)
        nbits.append(total)
        pos += (n + 3) & ~3
    return off, length, nbits, pos + 16


def stitch_into(buf, band_bytes, all_bits, n_streams):
    """Write the framed payload of every stream (iohelper.padded_bs_2_bytes layout) into `buf` (uint8, zeroed
    by this call) at stitch_layout's offsets.  band_bytes[b][s]: the raw bytes band b packed for stream s,
    its bits pre-shifted to their phase; neighbours share at most their first / last byte, which are ORed."""
    off, length, nbits, size = stitch_layout(all_bits, n_streams)
    assert buf.size >= size
    firsts = [bit_layout(all_bits, b)[1] for b in range(len(band_bytes))]
    for s in range(n_streams):
        view = buf[off[s]:off[s] + length[s]]
        view[:] = 0
        view[0] = 8 - (nbits[s] % 8)
        for b, per_band in enumerate(band_bytes):
            chunk = per_band[s] if isinstance(per_band[s], np.ndarray) else np.frombuffer(per_band[s], np.uint8)
            n = chunk.size
            if not n:
                continue
            a = int(firsts[b][s])
            if n > 2:
                view[a + 1:a + n - 1] = chunk[1:n - 1]           # the interior belongs to this band alone
            view[a] |= chunk[0]
            if n > 1:
                view[a + n - 1] |= chunk[n - 1]
    buf[size - 16:size] = 0
    return off, length, nbits


def stitch(band_bytes, all_bits, n_streams):
    """The framed payloads as byte strings."""
    size = stitch_layout(all_bits, n_streams)[3]
    buf = np.zeros(size, np.uint8)
    off, length, _ = stitch_into(buf, band_bytes, all_bits, n_streams)
    return [buf[o:o + n].tobytes() for o, n in zip(off, length)]


# ------------------------------------------------------------------------------------------------
# communicators
# ------------------------------------------------------------------------------------------------
class DistComm:
    """One process per band under torch.distributed.  The ranks of one box exchange their metadata through a
    mailbox in shared memory (a slot per rank and turn, a sequence word per rank: an all-gather is one pickle, one
    write and a short spin -- ~0.1 ms where a gloo object collective over eight ranks takes 2 ms, three times per
    encode); anything that does not fit a slot, and every exchange when `mailbox=False`, goes through the
    group's object collectives."""
    MAIL_SLOT = 4 << 20

    def __init__(self, group=None, mailbox=True):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self._mail, self._seq = None, 0
        if mailbox:
            self._open_mailbox()

    def _open_mailbox(self):
        import os
        token = "%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "0"))
        self._mail_path = "/dev/shm/hic_band_%s_mail.bin" % token
        size = 4096 + 2 * self.size * self.MAIL_SLOT
        try:
            if self.rank == 0:
                with open(self._mail_path, "wb") as f:
                    f.truncate(size)
            self.dist.barrier(group=self.group)
            self._mail = np.memmap(self._mail_path, dtype=np.uint8, mode="r+", shape=(size,))
            self._mail_seq = self._mail[:8 * self.size].view(np.int64)
            self.dist.barrier(group=self.group)
        except OSError:
            self._mail = None

    def _mail_slot(self, turn, rank):
        a = 4096 + (turn * self.size + rank) * self.MAIL_SLOT
        return self._mail[a:a + self.MAIL_SLOT]

    def all_gather(self, obj):
        import pickle
        import time
        if self._mail is not None:
            raw = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
            fits = len(raw) + 8 <= self.MAIL_SLOT
            k, turn = self._seq, self._seq & 1
            slot = self._mail_slot(turn, self.rank)
            if fits:
                slot[8:8 + len(raw)] = np.frombuffer(raw, np.uint8)
            slot[:8].view(np.int64)[0] = len(raw) if fits else -1
            self._mail_seq[self.rank] = k + 1                    # published: the payload was written before (x86 store order)
            self._seq += 1
            deadline = time.perf_counter() + 120.0
            while int(self._mail_seq.min()) < k + 1:
                if time.perf_counter() > deadline:
                    raise RuntimeError("band mailbox: a rank did not arrive within 120 s")
            sizes = [int(self._mail_slot(turn, r)[:8].view(np.int64)[0]) for r in range(self.size)]
            if all(n >= 0 for n in sizes):
                return [pickle.loads(self._mail_slot(turn, r)[8:8 + n].tobytes()) for r, n in enumerate(sizes)]
            # somebody's contribution did not fit: everybody saw that, everybody takes the collective
        out = [None] * self.size
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def gather(self, obj, root=0):
        """Gather to `root`.  Byte strings (the packed band payloads, the one bulky item) do not go through
        pickled object collectives: every rank of the box copies its strings once into a shared-memory file
        it keeps mapped (/dev/shm), the root maps the other ranks' files and gets VIEWS of them -- no second
        copy, no file reads -- and only the lengths travel as metadata.  Two files per rank alternate, so the
        views the root received stay valid until the gather after the next one."""
        if isinstance(obj, dict) and isinstance(obj.get("bytes"), list):
            lens_mine = [int(len(b)) for b in obj["bytes"]]
            offs_mine = self._offsets_in_staging(obj["bytes"])
            if offs_mine is None:                       # the strings live elsewhere: one copy into the shared file
                self._turn = getattr(self, "_turn", 0) ^ 1
                mm = self._shared_file(self.rank, self._turn, sum(lens_mine), create=True)
                pos, offs_mine = 0, []
                for b, n in zip(obj["bytes"], lens_mine):
                    mm[pos:pos + n] = np.frombuffer(b, np.uint8) if isinstance(b, (bytes, bytearray)) else b
                    offs_mine.append(pos)
                    pos += n
            self._staged = None
            # (the all-gather also orders every rank's writes before the root's reads)
            meta = self.all_gather((lens_mine, offs_mine))
            if self.rank != root:
                return None
            out = []
            for r, (ls, os_) in enumerate(meta):
                need = max([o + n for o, n in zip(os_, ls)] + [1])
                raw = self._shared_file(r, self._turn, need, create=False)
                out.append(dict(bytes=[raw[o:o + n] for o, n in zip(os_, ls)]))
            return out
        out = [None] * self.size if self.rank == root else None
        self.dist.gather_object(obj, out, dst=root, group=self.group)
        return out

    def staging(self, nbytes):
        """`nbytes` of this rank's shared-memory file for the NEXT gather of band strings, page-locked where CUDA
        is available: the device->host copy of the packed strings lands where the root will read them, and the
        gather itself moves only offsets and lengths."""
        self._turn = getattr(self, "_turn", 0) ^ 1
        mm = self._shared_file(self.rank, self._turn, max(int(nbytes), 1), create=True)
        self._staged = mm
        return mm[:nbytes]

    def _offsets_in_staging(self, parts):
        """Offsets of the byte strings inside the staging view handed out by staging(), or None."""
        mm = getattr(self, "_staged", None)
        if mm is None:
            return None
        base, offs = mm.ctypes.data, []
        for b in parts:
            if not isinstance(b, np.ndarray):
                return None
            if b.size == 0:
                offs.append(0)
                continue
            o = b.ctypes.data - base
            if o < 0 or o + b.size > mm.size:
                return None
            offs.append(int(o))
        return offs

    def _register(self, arr):
        """Page-lock a mapping for CUDA (best effort: without a device, or if the driver refuses, it stays pageable)."""
        try:
            from hiccup_b200 import _lib
            if _lib.load().hic_host_register(arr.ctypes.data, arr.size) == 0:
                self._pinned = getattr(self, "_pinned", {})
                self._pinned[arr.ctypes.data] = True
        except Exception:
            pass

    def _unregister(self, arr):
        if getattr(self, "_pinned", {}).pop(arr.ctypes.data, None):
            from hiccup_b200 import _lib
            _lib.load().hic_host_unregister(arr.ctypes.data)

    def _shared_file(self, rank, turn, nbytes, create):
        """uint8 view of rank `rank`'s shared-memory file number `turn`, at least `nbytes` long.  The owner
        creates and grows it (powers of two, so a growing band rarely remaps); readers remap when the file
        on disk is longer than their mapping."""
        import os
        if not hasattr(self, "_maps"):
            self._maps = {}
            self._token = "%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "0"))
        path = "/dev/shm/hic_band_%s_%d_%d.bin" % (self._token, rank, turn)
        cur = self._maps.get((rank, turn))
        if create:
            if cur is None or cur.size < nbytes:
                size = 1 << max(20, int(nbytes - 1).bit_length())
                with open(path, "ab") as f:
                    f.truncate(size)
                if cur is not None:
                    self._unregister(cur)
                cur = np.memmap(path, dtype=np.uint8, mode="r+", shape=(size,))
                self._maps[(rank, turn)] = cur
                self._register(cur)
        elif cur is None or cur.size < nbytes:
            if cur is not None:
                self._unregister(cur)
            # (mapped writable although only read: CUDA page-locks writable mappings without a special flag)
            cur = np.memmap(path, dtype=np.uint8, mode="r+", shape=(os.path.getsize(path),))
            self._maps[(rank, turn)] = cur
            self._register(cur)
        return cur

    def close(self):
        """Unmap and remove this rank's shared-memory files."""
        import os
        for (rank, turn), mm in list(getattr(self, "_maps", {}).items()):
            self._unregister(mm)
            del mm
            if rank == self.rank:
                try:
                    os.remove("/dev/shm/hic_band_%s_%d_%d.bin" % (self._token, rank, turn))
                except OSError:
                    pass
        self._maps = {}
        if getattr(self, "_mail", None) is not None:
            self._mail_seq = None
            self._mail = None
            if self.rank == 0:
                try:
                    os.remove(self._mail_path)
                except OSError:
                    pass


def run_local(workers):
    """Drive K band workers in one process (bands one after another on the current device, or on the
    devices the workers were created for).  Each worker's steps() is a generator that yields
    ("all_gather" | "gather", contribution) and receives the gathered list (None off the root of a
    gather).  Returns the workers' return values."""
    gens = [w.steps() for w in workers]
    msgs = [next(g) for g in gens]
    results = [None] * len(gens)
    done = False
    while not done:
        kind = msgs[0][0]
        gathered = [m[1] for m in msgs]
        new = []
        for i, g in enumerate(gens):
            try:
                new.append(g.send(gathered if (kind == "all_gather" or i == 0) else None))
            except StopIteration as stop:
                results[i] = stop.value
                done = True
        msgs = new
    return results


# ------------------------------------------------------------------------------------------------
# the band worker (GPU)
# ------------------------------------------------------------------------------------------------
def _host_huffman(freqs):
    from hiccup_b200 import _lib
    freqs = np.ascontiguousarray(freqs, np.uint32)
    lens, codes = np.empty(freqs.size, np.uint8), np.empty(freqs.size, np.uint64)
    _lib.check(_lib.load().hic_huffman_build_host(freqs.ctypes.data, freqs.size, lens.ctypes.data, codes.ctypes.data))
    return lens, codes


class BandWorker:
    """One band of an H x W image on one GPU.  `steps()` is the SPMD program: each `yield` is an
    all-gather (the value yielded is this band's contribution, the value sent back is the list over
    bands)."""

    def __init__(self, band, n_bands, h, w, r0, r1, device=None, value_bins=8192, stream=None):
        from hiccup_b200 import _lib, entropy
        self._lib, self.device = _lib, device
        self.band, self.n_bands, self.h, self.w, self.r0, self.r1 = band, n_bands, h, w, r0, r1
        self.s0, self.s1 = band_slice(h, r0, r1)
        self.stream = stream
        if device is not None:
            _lib.check(_lib.load().hic_set_device(int(device)))
        hs = self.s1 - self.s0
        self.g = _lib.geometry(hs, w)                      # geometry of the slice K1 runs on
        g = self.g
        skip_l, skip_c = (r0 - self.s0) // 8, (r0 - self.s0) // 16
        rows_l = -(-(r1 - r0) // 8)
        hc_band = min(r1, 2 * (h // 2)) // 2 - r0 // 2        # this band's rows of the (h // 2)-row chroma planes
        rows_c = -(-hc_band // 8)
        lay = _lib.StreamLayout()
        lay.n_images, lay.skip_first, lay.blocks_per_image = 1, 1, g.blocks_per_image
        nb = [rows_l * g.nbx_l, rows_c * g.nbx_c, rows_c * g.nbx_c]
        off = [skip_l * g.nbx_l, g.nb_l + skip_c * g.nbx_c, g.nb_l + g.nb_c + skip_c * g.nbx_c]
        for c in range(3):
            lay.nb[c], lay.block_off[c], lay.len[c] = nb[c], off[c], 63 * nb[c]
        self.layout, self.nb, self.block_off = lay, nb, off
        self.d_rgb = _lib.DeviceBuffer(hs * w * 3)
        self.d_coef = _lib.DeviceBuffer(g.blocks_per_image * 128)
        self.tie_capacity = _lib.tie_capacity(1, hs, w)
        self.d_ties = _lib.DeviceBuffer(self.tie_capacity * _lib.TIE_RECORD_BYTES)
        self.d_stats = _lib.DeviceBuffer(4 * _lib.TIE_STATS)
        self.encoder = entropy.EntropyEncoder(lay, value_bins)
        self.rgb, self._pinned, self._h_bytes = None, None, None
        self.staging = None          # optional callable(nbytes) -> uint8 array for the packed strings (DistComm.staging)

    def load(self, image):
        """image: the whole H x W x 3 host array (only this band's slice is uploaded)."""
        assert image.shape == (self.h, self.w, 3) and image.dtype == np.uint8
        hs = self.s1 - self.s0
        if self._pinned is None:
            self._pinned = self._lib.PinnedBuffer(hs * self.w * 3)
        self.rgb = self._pinned.array(np.uint8, hs * self.w * 3).reshape(hs, self.w, 3)
        self.rgb[...] = image[self.s0:self.s1]

    def _use_device(self):
        if self.device is not None:
            self._lib.check(self._lib.load().hic_set_device(int(self.device)))

    def steps(self):
        import time
        _lib, lib, st, enc, g = self._lib, self._lib.load(), self.stream, self.encoder, self.g
        self.trace = trace = [("start", time.perf_counter())]       # (phase that just ended, host clock): where a step's time goes
        mark = lambda name: trace.append((name, time.perf_counter()))
        self._use_device()
        self.d_rgb.upload(self.rgb, st)
        _lib.check(lib.hic_dct_forward(self.d_rgb.ptr, 1, self.s1 - self.s0, self.w, self.d_coef.ptr, self.d_ties.ptr,
                                       self.tie_capacity, self.d_stats.ptr, st))
        first_nz, last_nz = enc.scan(self.d_coef.ptr, st)
        last_dc = [int(self.d_coef.download(np.int16, 1, st, offset=128 * (self.block_off[c] + self.nb[c] - 1))[0])
                   for c in range(3)]
        mark("upload + K1 + scan")
        edges = yield ("all_gather", dict(first_nz=first_nz.tolist(), last_nz=last_nz.tolist(),
                                          length=[63 * n for n in self.nb], last_dc=last_dc))
        mark("exchange: edges")
        self._use_device()
        enc.emit(self.d_coef.ptr, seam_state(edges, self.band), st)
        index, entries, nsym_rl = enc.histograms(st)
        hists, nsym = [], []
        for s in range(9):
            a, n = int(index[s, 0]), int(index[s, 1])
            e = entries[a:a + n]
            hists.append((e[:, 0].copy(), e[:, 1].astype(np.uint32), e[:, 2].astype(np.uint32)))
            nsym.append(self.nb[s // 3] if s % 3 == KIND_DC else int(nsym_rl[s // 3]))
        mark("emit + histograms")
        gathered = yield ("all_gather", dict(hists=hists, nsym=nsym))
        mark("exchange: histograms")
        self._use_device()
        tables = build_tables(merge_histograms([m["hists"] for m in gathered], [m["nsym"] for m in gathered]), _host_huffman)
        all_bits = all_band_bits(tables, [m["hists"] for m in gathered])
        start_bit, _ = bit_layout(all_bits, self.band)
        mark("merge + Huffman tables (host)")
        enc.set_codes(tables[0], tables[1], tables[2], np.array(nsym, np.uint32), all_bits[self.band], start_bit, st)
        out = enc.pack(st)
        nbytes = int(enc.total_bytes)
        if self.staging is not None:                 # (a communicator's shared, page-locked memory: no copy at the gather)
            host = self.staging(nbytes)
        else:
            if self._h_bytes is None or self._h_bytes.nbytes < nbytes:
                if self._h_bytes is not None:
                    self._h_bytes.free()
                self._h_bytes = _lib.PinnedBuffer(nbytes + nbytes // 4 + 4096)
            host = self._h_bytes.array(np.uint8, nbytes)
        data = out.download(np.uint8, nbytes, st, out=host) if nbytes else np.zeros(0, np.uint8)
        # views of the pinned staging (valid until this worker's next pack)
        mine = [data[int(enc.byte_off[s]):int(enc.byte_off[s]) + int(enc.byte_len[s])] for s in range(9)]
        self.stats = self.d_stats.download(np.uint32, _lib.TIE_STATS, st)
        mark("pack + download")
        final = yield ("gather", dict(bytes=mine))               # the strings go to the root only
        mark("exchange: band strings")
        if final is None:
            return None
        return dict(tables=tables, all_bits=all_bits, band_bytes=[m["bytes"] for m in final])

    def close(self):
        self.encoder.close()
        for b in (self.d_rgb, self.d_coef, self.d_ties, self.d_stats, self._pinned, self._h_bytes):
            if b is not None:
                b.free()
        self._pinned = self._h_bytes = None


def assemble(result, h, w):
    """The reference's HicImage (codec.jpeg_encode's return value) from the stitched bands."""
    from hiccup_b200 import hicimage
    index, syms, packed = result["tables"]
    payloads = stitch(result["band_bytes"], result["all_bits"], 9)
    tables, bits = [], []
    for kind in range(3):
        for c in range(3):
            s = c * 3 + kind
            a, n = int(index[s, 0]), int(index[s, 1])
            pk = packed[a:a + n]
            tables.append(hicimage.PayloadStringP.from_arrays(syms[a:a + n], (pk >> np.uint64(58)).astype(np.uint8), pk & CODE_MASK,
                                                              kind == KIND_DC))
            bits.append(hicimage.BitStringP.from_framed(payloads[s]))
    return hicimage.HicImage.jpeg_image(tables + bits + [hicimage.TupP(h, w), hicimage.TupP(h // 2, w // 2)])


def to_encoded_streams(result, h, w, out=None):
    """The stitched bands as an entropy.EncodedStreams (what DctBatchCodec(1, h, w).decode takes).
    out: optional uint8 staging array (ideally page-locked) of at least stitch_layout(...)[3] bytes."""
    from hiccup_b200 import _lib, entropy
    index, syms, packed = result["tables"]
    size = stitch_layout(result["all_bits"], 9)[3]
    buf = np.zeros(size, np.uint8) if out is None else out[:size]
    off, length, nbits = stitch_into(buf, result["band_bytes"], result["all_bits"], 9)
    return entropy.EncodedStreams(_lib.layout_dct(1, h, w), index, np.zeros(9, np.uint32), np.array(nbits, np.uint64),
                                  np.array(off, np.uint64), np.array(length, np.uint64), syms, packed, buf[:size - 16])


def upload_stitched(result, d_data, stream=None):
    """The stitched framed payloads (stitch_into's layout) assembled in the DEVICE buffer `d_data` straight from the
    band strings: every band's interior bytes are one host->device copy to their place, and only the bytes two
    neighbours share (and the pad-count byte of each stream) are combined on the host.  Saves the host pass over
    the whole coded image that stitch_into() makes before an upload.  Returns (byte offsets, payload bits)."""
    from hiccup_b200 import _lib
    lib = _lib.load()
    all_bits, band_bytes = result["all_bits"], result["band_bytes"]
    off, length, nbits, size = stitch_layout(all_bits, 9)
    assert d_data.nbytes >= size
    _lib.check(lib.hic_memset(d_data.ptr, 0, size, stream))
    firsts = [bit_layout(all_bits, b)[1] for b in range(len(band_bytes))]
    shared = {}
    for s in range(9):
        shared[off[s]] = 8 - (nbits[s] % 8)
        for b, per_band in enumerate(band_bytes):
            chunk = per_band[s] if isinstance(per_band[s], np.ndarray) else np.frombuffer(per_band[s], np.uint8)
            n = chunk.size
            if not n:
                continue
            a = off[s] + int(firsts[b][s])
            if n > 2:
                _lib.check(lib.hic_memcpy_h2d(d_data.ptr + a + 1, chunk.ctypes.data + 1, n - 2, stream))
            shared[a] = shared.get(a, 0) | int(chunk[0])
            if n > 1:
                shared[a + n - 1] = shared.get(a + n - 1, 0) | int(chunk[n - 1])
    keys = sorted(shared)
    vals = np.array([shared[k] for k in keys], np.uint8)
    for i, k in enumerate(keys):
        _lib.check(lib.hic_memcpy_h2d(d_data.ptr + k, vals.ctypes.data + i, 1, stream))
    _lib.sync(stream)                  # `vals` and the band strings may go
    return np.array(off, np.uint64), np.array(nbits, np.uint64)


# ------------------------------------------------------------------------------------------------
# row-band sharded DECODE: the entropy decode of one image does not shard (the format has no restart points, so the
# symbol streams are one chain per channel), but everything after it does.  The root decodes the stitched payloads
# into the image's coefficient blocks, hands every band its block rows (plus one chroma block row of halo each side,
# for the 5-tap pyrUp) GPU to GPU, and every rank runs K7 / K8 on its rows and downloads THEM -- the decoded pixels
# (3 bytes per pixel, five times the coded image) leave through every GPU's PCIe link instead of the root's alone.
# The exchange itself is the caller's (bench_bands.py uses NCCL send / recv over NVLink); this is the bookkeeping.
# ------------------------------------------------------------------------------------------------
def band_block_ranges(g, s0, s1):
    """Block index ranges [first, last) of the rows [s0, s1) (multiples of 16) of an image with geometry g inside its
    coefficient buffer: (luminance, Cr, Cb)."""
    l0, l1 = (s0 // 8) * g.nbx_l, min(-(-s1 // 8), g.nby_l) * g.nbx_l
    c0, c1 = (s0 // 16) * g.nbx_c, min(-(-(s1 // 2) // 8), g.nby_c) * g.nbx_c
    return [(l0, l1), (g.nb_l + c0, g.nb_l + c1), (g.nb_l + g.nb_c + c0, g.nb_l + g.nb_c + c1)]


class BandDecoder:
    """One band's share of a sharded decode: a DctBatchCodec for the band's slice (its rows + halo) whose coefficient
    buffer the caller fills with band_block_ranges() of the whole image's blocks; inverse() runs K7 / K8 on the
    slice and fetch_rows() downloads the band's own rows."""

    def __init__(self, h, w, r0, r1, device=None):
        from hiccup_b200 import _lib
        from hiccup_b200.batch import DctBatchCodec
        self.h, self.w, self.r0, self.r1 = h, w, r0, r1
        self.s0, self.s1 = band_slice(h, r0, r1)
        self.g_image = _lib.geometry(h, w)
        self.codec = DctBatchCodec(1, self.s1 - self.s0, w, device=device)
        self.g = self.codec.g
        self.src = band_block_ranges(self.g_image, self.s0, self.s1)          # in the image's buffer
        self.dst = [(0, self.g.nb_l), (self.g.nb_l, self.g.nb_l + self.g.nb_c),
                    (self.g.nb_l + self.g.nb_c, self.g.nb_l + 2 * self.g.nb_c)]                 # in the slice's buffer
        for (a, b), (c, d) in zip(self.src, self.dst):
            assert b - a == d - c, "band blocks %r do not fill the slice planes %r" % (self.src, self.dst)

    @property
    def coef_ptr(self):
        return self.codec.d_coef_dec.ptr

    def inverse(self):
        self.codec._inverse()

    def fetch_rows(self, out_rows):
        """out_rows: uint8 array of the band's rows, (r1 - r0) x out_w x 3 (ideally page-locked)."""
        ow = self.codec.out_w
        n = (self.r1 - self.r0) * ow * 3
        assert out_rows.size == n and out_rows.flags.c_contiguous
        self.codec.d_out.download(np.uint8, n, self.codec.stream, offset=(self.r0 - self.s0) * ow * 3, out=out_rows.reshape(-1))

    def close(self):
        self.codec.close()


def encode_banded(image, n_bands, devices=None, value_bins=8192):
    """Encode one image as `n_bands` row bands in this process (devices: optional list of CUDA device
    indices, one per band, cycled; default: the current device for all).  Returns the HicImage."""
    image = np.ascontiguousarray(image, np.uint8)
    h, w = image.shape[:2]
    cuts = plan_bands(h, n_bands)
    workers = []
    for b in range(len(cuts) - 1):
        dev = None if not devices else devices[b % len(devices)]
        wk = BandWorker(b, len(cuts) - 1, h, w, cuts[b], cuts[b + 1], device=dev, value_bins=value_bins)
        wk.load(image)
        workers.append(wk)
    try:
        results = run_local(workers)
        return assemble(results[0], h, w)        # before close(): the band strings are views of pinned staging
    finally:
        for wk in workers:
            wk.close()


def encode_banded_dist(image, comm, device=None, value_bins=8192, worker=None):
    """SPMD form: rank r of `comm` encodes band r.  Every rank passes the same image array (only its slice
    is touched).  Returns the HicImage on rank 0, None elsewhere."""
    h, w = image.shape[:2]
    cuts = plan_bands(h, comm.size)
    if len(cuts) - 1 != comm.size:
        raise ValueError("image of %d rows cannot be cut into %d bands of 16-row multiples" % (h, comm.size))
    wk = worker or BandWorker(comm.rank, comm.size, h, w, cuts[comm.rank], cuts[comm.rank + 1], device=device,
                              value_bins=value_bins)
    wk.load(image)
    gen = wk.steps()
    kind, msg = next(gen)
    result = None
    try:
        while True:
            kind, msg = gen.send(comm.all_gather(msg) if kind == "all_gather" else comm.gather(msg, 0))
    except StopIteration as stop:
        result = stop.value
    try:
        return assemble(result, h, w) if comm.rank == 0 else None
    finally:
        if worker is None:
            wk.close()

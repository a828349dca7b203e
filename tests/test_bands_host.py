"""Host-side logic of row-band sharding (hiccup_b200/bands.py) on the CPU: the seam state, histogram
merge, shared Huffman tables, bit offsets and byte stitching must reproduce the whole-image encode of
the oracle byte for byte.  The per-band device work is stood in for by tests/band_oracle_worker.py.
Also covers the N > 1 path across two torch.distributed (gloo) processes."""
import os
import sys

import numpy as np
import pytest

from hiccup_b200 import bands
from oracle import hiccup_oracle as orc
from tests.band_oracle_worker import OracleBandWorker


def _whole(planes):
    enc = orc.jpeg_encode(planes)
    return ([[(int(a), b) for a, b in t] for t in enc["tables"]], [orc.padded_bits_to_bytes(b) for b in enc["bits"]])


def _check(hic, planes):
    tables, bits = _whole(planes)
    stream = hic.byte_stream()
    for i in range(9):
        assert [(int(a), b) for a, b in hic.payloads[i].rows] == tables[i], "table %d" % i
        assert stream[10 + i] == bits[i], "bit string %d" % i


def _image(kind, h, w, seed):
    if kind == "synthetic":
        return orc.synthetic_image(h, w, seed)
    if kind == "flat":                           # all-zero AC everywhere: only carried runs and one (0, 0)
        return np.full((h, w, 3), 90, np.uint8)
    if kind == "half":                           # detail in the top half only: bands below are all zero
        img = np.full((h, w, 3), 90, np.uint8)
        img[:h // 2] = orc.synthetic_image(h // 2, w, seed)
        return img
    if kind == "bottom":                         # detail in the bottom band only: a long carried run
        img = np.full((h, w, 3), 33, np.uint8)
        img[3 * h // 4:] = orc.synthetic_image(h - 3 * h // 4, w, seed)
        return img
    raise ValueError(kind)


def test_plan_bands():
    assert bands.plan_bands(16384, 8) == [2048 * i for i in range(9)]
    assert bands.plan_bands(100, 3) == [0, 32, 64, 100]
    assert bands.plan_bands(17, 2) == [0, 17]
    assert bands.plan_bands(40, 8) == [0, 16, 32, 40]
    for h in (16, 33, 250, 1000):
        for k in (1, 2, 3, 5, 8):
            cuts = bands.plan_bands(h, k)
            assert cuts[0] == 0 and cuts[-1] == h and all(b > a for a, b in zip(cuts, cuts[1:]))
            assert all(c % 16 == 0 for c in cuts[1:-1])


@pytest.mark.parametrize("kind,shape,k", [("synthetic", (96, 80), 2), ("synthetic", (130, 72), 3), ("synthetic", (64, 64), 4),
                                          ("flat", (64, 48), 3), ("half", (128, 64), 4), ("bottom", (128, 64), 4),
                                          ("synthetic", (50, 34), 2), ("half", (96, 40), 3)])
def test_stitched_bands_equal_whole_image(kind, shape, k):
    rgb = _image(kind, shape[0], shape[1], 7)
    planes = orc.jpeg_compression(rgb)
    cuts = bands.plan_bands(shape[0], k)
    workers = [OracleBandWorker(b, len(cuts) - 1, planes, cuts[b], cuts[b + 1]) for b in range(len(cuts) - 1)]
    results = bands.run_local(workers)
    assert all(r is None for r in results[1:])
    _check(bands.assemble(results[0], shape[0], shape[1]), planes)


def test_seam_state_rules():
    edges = [dict(first_nz=[3, -1, 0], last_nz=[10, -1, 62], length=[63, 63, 63], last_dc=[5, 6, 7]),
             dict(first_nz=[-1, -1, 1], last_nz=[-1, -1, 1], length=[63, 63, 63], last_dc=[8, 9, 10]),
             dict(first_nz=[0, -1, -1], last_nz=[0, -1, -1], length=[63, 63, 63], last_dc=[1, 2, 3])]
    assert bands.seam_state(edges, 0) == [(0, 0, 1, 0), (0, 0, 0, 0), (0, 0, 1, 0)]
    assert bands.seam_state(edges, 1) == [(52, 5, 1, 0), (63, 6, 0, 0), (0, 7, 0, 0)]
    assert bands.seam_state(edges, 2) == [(52 + 63, 8, 0, 1), (126, 9, 0, 1), (61, 10, 0, 1)]


def _gloo_worker(rank, world, port, shape, seed, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rgb = orc.synthetic_image(shape[0], shape[1], seed)
        planes = orc.jpeg_compression(rgb)
        comm = bands.DistComm()
        cuts = bands.plan_bands(shape[0], world)
        wk = OracleBandWorker(rank, world, planes, cuts[rank], cuts[rank + 1])
        wk.load = lambda image: None
        hic = bands.encode_banded_dist(rgb, comm, worker=wk)
        if rank == 0:
            _check(hic, planes)
            with open(out_path, "w") as f:
                f.write("ok")
        else:
            assert hic is None
        dist.barrier()
        comm.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_band_encode(tmp_path):
    """world_size 2 over gloo: every rank owns one band, metadata travels by object collectives."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "done")
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_gloo_worker, args=(2, port, (96, 64), 3, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shared_memory_band_files_grow_and_remap(monkeypatch):
    """The band strings travel through mapped /dev/shm files: the owner grows its file, a reader that mapped
    the shorter file remaps, and views taken on one turn survive the other turn's write."""
    monkeypatch.setenv("MASTER_PORT", "test%d" % os.getpid())
    owner, reader = object.__new__(bands.DistComm), object.__new__(bands.DistComm)
    owner.rank, reader.rank = 1, 0
    try:
        a = owner._shared_file(1, 0, 1000, create=True)
        a[:1000] = np.arange(1000, dtype=np.uint8)
        view = reader._shared_file(1, 0, 1000, create=False)[:1000]
        assert np.array_equal(view, np.arange(1000, dtype=np.uint8))
        big = (3 << 20) + 17                                   # beyond the first power-of-two size
        b = owner._shared_file(1, 0, big, create=True)
        assert b.size >= big and np.array_equal(b[:1000], np.arange(1000, dtype=np.uint8))     # growing keeps the content
        b[big - 1] = 99
        assert reader._shared_file(1, 0, big, create=False)[big - 1] == 99
        other = owner._shared_file(1, 1, 500, create=True)      # the alternate file of the same rank
        other[:500] = 7
        assert np.array_equal(view, np.arange(1000, dtype=np.uint8))
    finally:
        owner.close()
        reader.close()
    assert not [f for f in os.listdir("/dev/shm") if ("test%d" % os.getpid()) in f]


def test_staged_strings_are_gathered_without_a_copy(monkeypatch):
    """DistComm.staging(): strings written into the rank's shared file are announced by offset and length only;
    the root's views alias the owner's memory.  (One process plays both ranks; the all-gather is stubbed.)"""
    monkeypatch.setenv("MASTER_PORT", "stg%d" % os.getpid())
    owner, root = object.__new__(bands.DistComm), object.__new__(bands.DistComm)
    owner.rank, owner.size, root.rank, root.size = 1, 2, 0, 2
    try:
        box = {}
        owner.all_gather = lambda meta: box.setdefault("owner", meta) and None
        st = owner.staging(300)
        st[:300] = np.arange(300, dtype=np.uint8)
        parts = [st[0:100], st[100:100], st[104:300]]
        assert owner.gather(dict(bytes=parts), root=0) is None
        lens, offs = box["owner"]
        assert lens == [100, 0, 196] and offs == [0, 0, 104]
        # the root's side of the same exchange: its own strings (copied: they are plain bytes) + the owner's views
        root.all_gather = lambda meta: [meta, box["owner"]]
        got = root.gather(dict(bytes=[b"abc", b"", b"xy"]), root=0)
        assert [bytes(b) for b in got[0]["bytes"]] == [b"abc", b"", b"xy"]
        assert [bytes(b) for b in got[1]["bytes"]] == [bytes(p) for p in parts]
        st[5] = 200                                              # a view, not a copy
        assert got[1]["bytes"][0][5] == 200
        # strings that are not in the staging area still travel (by one copy)
        box.clear()
        owner.all_gather = lambda meta: box.setdefault("owner", meta) and None
        owner.gather(dict(bytes=[np.arange(10, dtype=np.uint8)]), root=0)
        assert box["owner"][0] == [10]
    finally:
        owner.close()
        root.close()
    assert not [f for f in os.listdir("/dev/shm") if ("stg%d" % os.getpid()) in f]


@pytest.mark.parametrize("h,w,k", [(16384, 16384, 8), (2048, 512, 5), (426, 640, 4), (1080, 1920, 8), (96, 80, 2)])
def test_band_block_ranges_fill_the_slice_geometry(h, w, k):
    """The sharded decode hands band b the block rows of its slice (rows + halo): three contiguous ranges of the image's
    coefficient buffer that must be exactly the three planes of the slice's own geometry, and the bands' own rows
    (without halo) must tile the image."""
    from hiccup_b200 import _lib
    g = _lib.geometry(h, w)
    cuts = bands.plan_bands(h, k)
    covered = 0
    for b in range(len(cuts) - 1):
        s0, s1 = bands.band_slice(h, cuts[b], cuts[b + 1])
        gs = _lib.geometry(s1 - s0, w)
        (l0, l1), (r0, r1), (b0, b1) = bands.band_block_ranges(g, s0, s1)
        assert (l1 - l0, r1 - r0, b1 - b0) == (gs.nb_l, gs.nb_c, gs.nb_c)
        assert 0 <= l0 < l1 <= g.nb_l and g.nb_l <= r0 < r1 <= g.nb_l + g.nb_c and g.nb_l + g.nb_c <= b0 < b1 <= g.blocks_per_image
        assert l0 == (s0 // 8) * g.nbx_l and r0 - g.nb_l == (s0 // 16) * g.nbx_c == b0 - g.nb_l - g.nb_c
        assert s0 % 16 == 0 and (s1 % 16 == 0 or s1 == h) and s0 <= cuts[b] and cuts[b + 1] <= s1
        covered += cuts[b + 1] - cuts[b]
    assert covered == h

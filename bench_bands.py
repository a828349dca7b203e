#!/usr/bin/env python
"""bench_bands.py -- BASELINE config 5: one 16384x16384 RGB image, DCT mode, row-band sharded over the
GPUs of one box (hiccup_b200/bands.py), then decoded.

    python bench_bands.py [--steps K] [--warmup W] [--size 16384]                       # 1 GPU = 1 band
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench_bands.py --gpus N                                          # N bands

A step = encode of the whole image as N bands (one per rank; the only cross-rank traffic is the
metadata of bands.py over a gloo group -- no data-path collective) + decode of the stitched `.hic`
stream on rank 0 (the format has no restart offsets, so the entropy decode of ONE image does not shard;
DESIGN.md section 7).  Prints one JSON line on rank 0: encode MP/s (max over ranks), decode MP/s, and the
combined encode+decode MP/s.  `strong` scaling: the image is fixed, the bands shrink with N.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (shared helpers: synthetic images, stdout protection)


def big_synthetic(size, tile=2048):
    """size x size x 3: a `tile`-sized synthetic image (SURVEY 8(d) recipe) repeated; generating 805 MB
    of bicubic noise directly would take minutes of host time."""
    base = bench.synthetic_image(min(tile, size), min(tile, size), 5000)
    reps = -(-size // base.shape[0])
    return np.ascontiguousarray(np.tile(base, (reps, reps, 1))[:size, :size])


def main():
    bench.claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--check", action="store_true", help="compare the stitched stream with a one-band encode on rank 0")
    ap.add_argument("--shard-decode", action="store_true",
                    help="inverse transform and pixel download of every band on its own GPU (coefficient bands over NCCL)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    torch.cuda.set_device(local_rank)
    from hiccup_b200 import _lib, bands, codec, compression
    _lib.check(_lib.load().hic_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        meta = dist.new_group(backend="gloo")            # host-side metadata exchange
        comm = bands.DistComm(meta)
    size = args.size
    image = big_synthetic(size)
    cuts = bands.plan_bands(size, world)
    worker = bands.BandWorker(rank, world, size, size, cuts[rank], cuts[rank + 1], device=local_rank)
    worker.load(image)
    if dist is not None:
        worker.staging = comm.staging          # the packed strings go straight into the shared, page-locked file

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def encode_step():
        if world == 1:
            res = bands.run_local([worker])[0]
        else:
            gen = worker.steps()
            kind, msg = next(gen)
            res = None
            try:
                while True:
                    kind, msg = gen.send(comm.all_gather(msg) if kind == "all_gather" else comm.gather(msg, 0))
            except StopIteration as stop:
                res = stop.value
        return res

    decoder, staging = None, None

    def decode_step(res):
        """The band strings go to the device at their places in the stitched payloads (bands.upload_stitched: only the
        bytes two bands share pass through host code), entropy decode + inverse transform, pixels to a host buffer."""
        nonlocal decoder, staging
        from hiccup_b200.batch import DctBatchCodec
        if decoder is None:
            decoder = DctBatchCodec(1, size, size)
            staging = _lib.DeviceBuffer(bands.stitch_layout(res["all_bits"], 9)[3] + (1 << 20))
        off, nbits = bands.upload_stitched(res, staging)
        index, syms, packed = res["tables"]
        decoder.decoder.decode_device_data(index, syms, packed, staging.ptr, off, nbits, decoder.d_coef_dec.ptr)
        decoder._inverse()
        return decoder.fetch()

    # ---- sharded decode: the root's entropy decode, then every band's K7 / K8 and download on its own GPU ----
    class _Cai:          # a raw device pointer as a torch byte tensor (NCCL send / recv wants tensors, and has no int16)
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}

    shard = bool(args.shard_decode and world > 1)
    band_dec, image_coef, shared_out, shared_map = None, None, None, None
    if shard:
        band_dec = bands.BandDecoder(size, size, cuts[rank], cuts[rank + 1], device=local_rank)
        path = "/dev/shm/hic_bands_out_%s.bin" % os.environ.get("MASTER_PORT", "0")
        nbytes = size * size * 3
        if rank == 0:
            with open(path, "wb") as f:
                f.truncate(nbytes)
        dist.barrier()
        shared_map = np.memmap(path, dtype=np.uint8, mode="r+", shape=(nbytes,))
        _lib.load().hic_host_register(shared_map.ctypes.data, nbytes)
        shared_out = shared_map.reshape(size, size, 3)
        dist.barrier()

    def decode_sharded(res):
        nonlocal decoder, staging, image_coef
        from hiccup_b200.batch import DctBatchCodec
        g = band_dec.g_image
        mine = torch.as_tensor(_Cai(band_dec.coef_ptr, band_dec.g.blocks_per_image * 128), device="cuda")     # 128 bytes per block
        ops = []
        if rank == 0:
            if decoder is None:
                decoder = DctBatchCodec(1, size, size)
                staging = _lib.DeviceBuffer(bands.stitch_layout(res["all_bits"], 9)[3] + (1 << 20))
            off, nbits = bands.upload_stitched(res, staging)
            index, syms, packed = res["tables"]
            decoder.decoder.decode_device_data(index, syms, packed, staging.ptr, off, nbits, decoder.d_coef_dec.ptr)
            whole = torch.as_tensor(_Cai(decoder.d_coef_dec.ptr, g.blocks_per_image * 128), device="cuda")
            for r in range(world):
                s0, s1 = bands.band_slice(size, cuts[r], cuts[r + 1])
                for (a, b), (c, d) in zip(bands.band_block_ranges(g, s0, s1), band_dec.dst if r == 0 else [(0, 0)] * 3):
                    if r == 0:
                        mine[128 * c:128 * d].copy_(whole[128 * a:128 * b])
                    else:
                        ops.append(dist.P2POp(dist.isend, whole[128 * a:128 * b], r))
        else:
            for c, d in band_dec.dst:
                ops.append(dist.P2POp(dist.irecv, mine[128 * c:128 * d], 0))
        if ops:
            for w_ in dist.batch_isend_irecv(ops):
                w_.wait()
        torch.cuda.synchronize()
        band_dec.inverse()
        band_dec.fetch_rows(shared_out[cuts[rank]:cuts[rank + 1]])

    res = None
    for _ in range(args.warmup):
        res = encode_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = encode_step()
    barrier()
    t_enc = (time.perf_counter() - t0) / args.steps
    # where the last encode step's time went on this rank (host clock between the phases of BandWorker.steps), max over ranks
    phases = [(b[0], (b[1] - a[1]) * 1e3) for a, b in zip(worker.trace, worker.trace[1:])]
    if dist is not None:
        t = torch.tensor([t_enc] + [p[1] for p in phases], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_enc = float(t[0].item())
        phases = [(p[0], float(v)) for p, v in zip(phases, t[1:].tolist())]
    t_dec, same, out = None, None, None
    hic = None
    if rank == 0:
        out = decode_step(res)
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps // 2)):
            out = decode_step(res)
        torch.cuda.synchronize()
        t_dec = (time.perf_counter() - t0) / max(1, args.steps // 2)
        hic = bands.assemble(res, size, size)
        if args.check:
            whole = bands.encode_banded(image, 1)
            same = whole.byte_stream() == hic.byte_stream()
            # decoded pixels: a 512x512 window against the oracle's encode+decode of that window alone (exact away
            # from the window's border, where the chroma pyramids see other neighbours)
            from oracle import hiccup_oracle as orc
            y0, x0, win, m = (size // 2 // 16) * 16, (size // 2 // 16) * 16 - 512, 512, 32
            y0, x0 = max(0, min(y0, size - win)), max(0, min(x0, size - win))
            ref = orc.jpeg_decompression(orc.jpeg_compression(image[y0:y0 + win, x0:x0 + win]))
            same = bool(same and np.array_equal(out[0][y0 + m:y0 + win - m, x0 + m:x0 + win - m], ref[m:-m, m:-m]))
    barrier()
    t_dec_shard, shard_same = None, None
    if shard:
        decode_sharded(res)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps // 2)):
            decode_sharded(res)
            barrier()
        t_dec_shard = (time.perf_counter() - t0) / max(1, args.steps // 2)
        t = torch.tensor([t_dec_shard], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dec_shard = float(t.item())
        if rank == 0 and out is not None:
            shard_same = bool(np.array_equal(shared_out, out[0]))
    barrier()
    if rank == 0:
        mp = size * size / 1e6
        nbytes = sum(len(b) for b in hic.byte_stream())
        line = {
            "metric": bench.METRIC, "value": mp / (t_enc + t_dec), "unit": bench.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (t_enc + t_dec) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%dx%d RGB single image, DCT mode, row-band sharded across %d B200 (encode), decoded on rank 0"
                                   % (size, size, world),
                       "bands": [int(b - a) for a, b in zip(cuts, cuts[1:])], "parallelism": "row bands, host-side stitching, no collective",
                       "timing": "host wall clock per step incl. host<->device copies, metadata exchange and stitching; max over ranks",
                       "hic_bytes": nbytes},
            "encode": {"value": mp / t_enc, "unit": bench.UNIT, "ms": t_enc * 1e3,
                       "phases_ms_max_over_ranks": {k: round(v, 2) for k, v in phases}},
            "decode": {"value": mp / t_dec, "unit": bench.UNIT, "ms": t_dec * 1e3, "where": "rank 0 only"},
            "decode_sharded": None if t_dec_shard is None else {
                "value": mp / t_dec_shard, "unit": bench.UNIT, "ms": t_dec_shard * 1e3, "step_ms": (t_enc + t_dec_shard) * 1e3,
                "equals_rank0_decode": shard_same,
                "where": "entropy decode on rank 0, coefficient bands to their ranks over NCCL, K7 / K8 and the pixel download on every rank"},
            "matches_one_band_encode": same,
        }
        bench.emit(line)
    if shard:
        _lib.load().hic_host_unregister(shared_map.ctypes.data)
        if rank == 0:
            try:
                os.remove("/dev/shm/hic_bands_out_%s.bin" % os.environ.get("MASTER_PORT", "0"))
            except OSError:
                pass
    if dist is not None:
        dist.barrier()
        comm.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

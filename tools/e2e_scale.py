"""Host-to-host C2 step of PipelinedCodec.round_trip on N ranks of one box, swept over what the host
side can change: slots, chunk size, rank -> core placement, how waiting threads wait, and the kind of
page-locked memory (development aid; run under torch.distributed.run, gloo for the max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/e2e_scale.py [--configs "slots=8,chunk=128;slots=3,chunk=128,pin=1;..."] [--steps 4]

Keys of a configuration: slots, chunk, pin (0 none | 1 disjoint core sets in rank order | 2 reversed),
block (1 = blocking waits), mem (cuda = cudaMallocHost | thp = THP-backed mmap + cudaHostRegister),
copy_only (1 = the same copies with no kernels: the floor of this exact pattern, per rank and slot).
Rank 0 prints one JSON line per configuration: ms per step (max over ranks), aggregate MP/s.
"""
import argparse
import ctypes
import json
import mmap
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from hiccup_b200 import _lib
from hiccup_b200.batch import PipelinedCodec

N, H, W = 1024, 426, 640


def thp_pinned(nbytes):
    """Anonymous 2 MB-aligned mapping advised MADV_HUGEPAGE, touched, then registered with CUDA."""
    two = 2 << 20
    size = (nbytes + two - 1) // two * two
    m = mmap.mmap(-1, size + two, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
    base = (addr + two - 1) // two * two
    libc = ctypes.CDLL(None, use_errno=True)
    libc.madvise(ctypes.c_void_p(base), ctypes.c_size_t(size), 14)         # MADV_HUGEPAGE
    arr = np.frombuffer(m, dtype=np.uint8, count=size, offset=base - addr)
    arr[:] = 0
    rt = ctypes.CDLL("libcudart.so.12")
    rc = rt.cudaHostRegister(ctypes.c_void_p(base), ctypes.c_size_t(size), 0)
    if rc != 0:
        raise RuntimeError("cudaHostRegister -> %d" % rc)
    return arr[:nbytes], m


def parse(text):
    out = []
    for part in text.split(";"):
        part = part.strip()
        if not part:
            continue
        cfg = dict(slots=8, chunk=128, pin=0, block=0, mem="cuda", copy_only=0)
        for kv in part.split(","):
            k, v = kv.split("=")
            cfg[k.strip()] = v.strip() if k.strip() == "mem" else int(v)
        out.append(cfg)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="slots=8,chunk=128;slots=4,chunk=128;slots=3,chunk=128;slots=2,chunk=128;"
                                         "slots=8,chunk=128,pin=1;slots=4,chunk=128,pin=1;slots=8,chunk=128,pin=2;"
                                         "slots=8,chunk=128,block=1;slots=8,chunk=128,mem=thp;slots=8,chunk=128,mem=thp,pin=1;"
                                         "slots=8,chunk=128,copy_only=1;slots=8,chunk=128,copy_only=1,mem=thp")
    ap.add_argument("--steps", type=int, default=4)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    _lib.check(_lib.load().hic_set_device(local))
    ncpu = os.cpu_count() or 1
    all_cores = set(range(ncpu))
    base = bench.synthetic_batch(32, H, W, 2000 + 100 * rank)
    buffers = {}

    def host_buffers(mem, pin):
        key = (mem, pin)
        if key not in buffers:
            out_shape = (N, 2 * (H // 2), 2 * (W // 2), 3)
            if mem == "thp":
                a, k1 = thp_pinned(N * H * W * 3)
                b, k2 = thp_pinned(int(np.prod(out_shape)))
                a, b = a.reshape(N, H, W, 3), b.reshape(out_shape)
            else:
                a, k1 = bench.pinned_array(_lib, (N, H, W, 3))
                b, k2 = bench.pinned_array(_lib, out_shape)
            for i in range(N):
                a[i] = base[i % 32]
            buffers[key] = (a, b, k1, k2)
        return buffers[key][:2]

    def sync_max(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for cfg in parse(args.configs):
        if cfg["pin"]:
            per = max(1, ncpu // world)
            slot = rank if cfg["pin"] == 1 else world - 1 - rank
            os.sched_setaffinity(0, set(range(slot * per % ncpu, slot * per % ncpu + per)))
        else:
            os.sched_setaffinity(0, all_cores)
        host, out = host_buffers(cfg["mem"], cfg["pin"])
        err = None
        ms = float("nan")
        try:
            if cfg["copy_only"]:
                ms = copy_only(host, out, cfg, args.steps, dist if world > 1 else None)
            else:
                pipe = PipelinedCodec(N, H, W, chunk=cfg["chunk"], slots=cfg["slots"], device=local, blocking_sync=bool(cfg["block"]))
                if not cfg["block"]:
                    _lib.check(_lib.load().hic_set_blocking_sync(0))
                pipe.round_trip(host, out)
                pipe.round_trip(host, out)
                if world > 1:
                    dist.barrier()
                t = time.perf_counter()
                pipe.round_trip(host, out, repeat=args.steps)
                _lib.sync()
                ms = (time.perf_counter() - t) * 1e3 / args.steps
                pipe.close()
        except Exception as e:          # report and go on with the next configuration
            err = repr(e)
        ms = sync_max(ms)
        if rank == 0:
            line = dict(cfg)
            line.update(n_gpus=world, cpus=ncpu, ms_per_step=round(ms, 2), aggregate_mps=round(world * N * H * W / 1e6 / (ms / 1e3), 0) if ms == ms else None,
                        error=err)
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def copy_only(host, out, cfg, steps, dist):
    """The bulk copies of round_trip with no kernels: per chunk one upload of its images, one download of a
    payload-sized piece and one download of its pixels, on `slots` streams driven by `slots` threads with the
    same two gates."""
    import threading
    chunk, slots = cfg["chunk"], cfg["slots"]
    n_chunks = N // chunk
    in_b, out_b = chunk * H * W * 3, chunk * out.shape[1] * out.shape[2] * 3
    pay_b = 156_322_772 // n_chunks + 4_000_000 // n_chunks
    streams = [_lib.stream_create() for _ in range(slots)]
    d_in = [_lib.DeviceBuffer(in_b) for _ in range(slots)]
    d_out = [_lib.DeviceBuffer(out_b) for _ in range(slots)]
    d_pay = [_lib.DeviceBuffer(pay_b) for _ in range(slots)]
    h_pay = [_lib.PinnedBuffer(pay_b) for _ in range(slots)]
    gate_in, gate_out = threading.Lock(), threading.Lock()
    lib = _lib.load()
    flat_in, flat_out = host.reshape(-1), out.reshape(-1)

    def work(slot, repeat):
        _lib.check(lib.hic_set_device(int(os.environ.get("LOCAL_RANK", "0"))))
        st = streams[slot]
        for v in range(slot, repeat * n_chunks, slots):
            c = v % n_chunks
            with gate_in:
                d_in[slot].upload(flat_in[c * in_b:(c + 1) * in_b], st)
                _lib.sync(st)
            d_pay[slot].download(np.uint8, pay_b, st, out=h_pay[slot].array(np.uint8, pay_b))
            with gate_out:
                d_out[slot].download(np.uint8, out_b, st, out=flat_out[c * out_b:(c + 1) * out_b])

    def run(repeat):
        th = [threading.Thread(target=work, args=(s, repeat)) for s in range(slots)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    run(1)
    if dist is not None:
        dist.barrier()
    t = time.perf_counter()
    run(steps)
    _lib.sync()
    ms = (time.perf_counter() - t) * 1e3 / steps
    for b in d_in + d_out + d_pay + h_pay:
        b.free()
    for st in streams:
        _lib.stream_destroy(st)
    return ms


if __name__ == "__main__":
    main()

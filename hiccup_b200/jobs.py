"""A batch job shared by the ranks (one process per GPU) of one box -- additive: the reference is one
process on one image (run.py:18-43).

north_star partitions work "across the 8 GPUs of one box by image ... using host-side offset stitching
rather than NCCL".  Which image goes to which GPU is the runtime's choice, and on a real box the GPUs'
paths to host memory are NOT alike: measured with tools/pcie_floor.cu on a 4-GPU B200 VM, GPU 0 moves
a C2 step's bytes (0.84 GB in, 1.06 GB out) in 38 ms while GPUs 1-3, which share one path (another
socket or switch uplink: ~60 GB/s for all three together), need 93 ms each.  A fixed share per rank
finishes when the slowest link does.  So the job keeps every rank's images and results in shared,
page-locked host memory and hands chunks out from one ticket counter: a rank takes the next chunk
when one of its pipeline slots is free, whoever owns the images -- faster links take more chunks and
all links stay busy until the job is done.

    job = SharedJob("c2", rank, world, n_per_rank, h, w, out_h, out_w, barrier=dist.barrier)
    job.inputs[rank][...] = my images                    # every rank fills its own part
    job.reset(); pipe.run_job(job)                       # PipelinedCodec.run_job (batch.py)
    job.outputs[r]                                       # decoded pixels of rank r's images, on every rank
    job.close()

Memory: /dev/shm files mapped by every rank and registered with CUDA (hic_host_register), so bulk
copies to and from them are asynchronous and full rate in every process; the ticket counter is a word
of another shared mapping (hic_ticket_take = one atomic fetch-add).  No GPU-to-GPU traffic at all.
"""
import ctypes
import mmap
import os

import numpy as np

from hiccup_b200 import _lib


class SharedJob:
    CTL_BYTES = 4096

    def __init__(self, name, rank, world, n, h, w, out_h, out_w, barrier=None, base="/dev/shm", register=True):
        """n images of h x w per rank.  `barrier`: a callable that synchronises the ranks (None when world
        is 1).  Collective: every rank must construct the job with the same arguments.  `register=False`
        leaves the mappings pageable (host-logic tests without a GPU)."""
        if register:
            _lib.require_device()
        self._register = bool(register)
        self.lib = _lib.load()
        self.rank, self.world, self.n = int(rank), int(world), int(n)
        self.in_shape = (self.n, int(h), int(w), 3)
        self.out_shape = (self.n, int(out_h), int(out_w), 3)
        self._barrier = barrier if barrier is not None else (lambda: None)
        self._base = os.path.join(base, "hic_job_%s" % name)
        self._maps, self._registered, self._owned = [], [], []
        in_bytes, out_bytes = int(np.prod(self.in_shape)), int(np.prod(self.out_shape))
        try:
            self._create(self._path("in", self.rank), in_bytes)
            self._create(self._path("out", self.rank), out_bytes)
            if self.rank == 0:
                self._create(self._path("ctl", 0), self.CTL_BYTES)
            self._barrier()
            self.inputs = [self._map(self._path("in", r), in_bytes).reshape(self.in_shape) for r in range(self.world)]
            self.outputs = [self._map(self._path("out", r), out_bytes).reshape(self.out_shape) for r in range(self.world)]
            ctl = self._map(self._path("ctl", 0), self.CTL_BYTES, register=False)
            self._ctl_addr = ctl.ctypes.data
            self._ctl = ctl.view(np.int64)
            self._barrier()
        except Exception:
            self.close()
            raise

    def _path(self, kind, r):
        return "%s_%s%d" % (self._base, kind, r)

    def _create(self, path, nbytes):
        fd = os.open(path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
        try:
            os.ftruncate(fd, nbytes)
            os.posix_fallocate(fd, 0, nbytes)          # fail now, not with a SIGBUS later, if /dev/shm is too small
        finally:
            os.close(fd)
        self._owned.append(path)

    def _map(self, path, nbytes, register=True):
        fd = os.open(path, os.O_RDWR)
        try:
            m = mmap.mmap(fd, nbytes, flags=mmap.MAP_SHARED, prot=mmap.PROT_READ | mmap.PROT_WRITE)
        finally:
            os.close(fd)
        arr = np.frombuffer(m, dtype=np.uint8, count=nbytes)
        self._maps.append((m, arr))
        if register and self._register:
            _lib.check(self.lib.hic_host_register(arr.ctypes.data, nbytes))
            self._registered.append(arr.ctypes.data)
        return arr

    # ---- tickets ---------------------------------------------------------------------------------
    def reset(self):
        """Rewind the ticket counter (collective: all ranks call it; returns after everybody has)."""
        self._barrier()
        if self.rank == 0:
            self._ctl[0] = 0
        self._barrier()

    def take(self):
        first = ctypes.c_int64(0)
        _lib.check(self.lib.hic_ticket_take(self._ctl_addr, 1, ctypes.byref(first)))
        return int(first.value)

    def close(self):
        for addr in self._registered:
            self.lib.hic_host_unregister(addr)
        self._registered = []
        self.inputs, self.outputs, self._ctl = [], [], None
        maps, self._maps = self._maps, []
        for m, arr in maps:
            del arr
        for m, _ in maps:
            try:
                m.close()
            except BufferError:          # a view is still alive somewhere: the mapping goes with the process
                pass
        try:
            self._barrier()
        except Exception:
            pass
        for path in self._owned:
            try:
                os.unlink(path)
            except OSError:
                pass
        self._owned = []

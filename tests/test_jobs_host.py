"""Host logic of the box-wide job queue (hiccup_b200/jobs.py): shared mappings and the ticket counter across
two torch.distributed (gloo) ranks, no GPU (register=False leaves the mappings pageable)."""
import os

import numpy as np


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from hiccup_b200 import jobs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        job = jobs.SharedJob("t%d" % port, rank, world, 6, 4, 8, 4, 8, barrier=dist.barrier, register=False)
        job.inputs[rank][...] = rank + 1                       # every rank fills its own images
        job.reset()
        assert all(int(job.inputs[r][0, 0, 0, 0]) == r + 1 for r in range(world))      # and sees the others'
        # the by-image partition of PipelinedCodec.run_job, with a copy instead of the codec: chunks of 2 images
        per_rank, chunk, repeat = 3, 2, 3
        total = repeat * world * per_rank
        mine = []
        while True:
            v = job.take()
            if v >= total:
                break
            u = v % (world * per_rank)
            owner, c = u // per_rank, u % per_rank
            job.outputs[owner][c * chunk:(c + 1) * chunk] = job.inputs[owner][c * chunk:(c + 1) * chunk] + 10
            mine.append(v)
        dist.barrier()
        for r in range(world):
            assert np.all(job.outputs[r] == r + 11)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            every = sorted(v for part in gathered for v in part)
            assert every == list(range(total)), every           # each ticket taken exactly once, box-wide
            with open(os.path.join(out_dir, "ok"), "w") as f:
                f.write("ok")
        job.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_job_queue(tmp_path):
    import torch.multiprocessing as mp
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "ok")).read() == "ok"
    assert not [f for f in os.listdir("/dev/shm") if ("hic_job_t%d" % port) in f]


def test_ticket_counter_is_atomic_across_threads():
    import ctypes
    import threading
    from hiccup_b200 import _lib
    lib = _lib.load()
    counter = np.zeros(8, np.int64)
    seen = [[] for _ in range(8)]

    def take(i):
        first = ctypes.c_int64(0)
        for _ in range(2000):
            _lib.check(lib.hic_ticket_take(counter.ctypes.data, 1, ctypes.byref(first)))
            seen[i].append(first.value)

    threads = [threading.Thread(target=take, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert sorted(v for part in seen for v in part) == list(range(16000)) and int(counter[0]) == 16000

#!/usr/bin/env python
"""bench.py -- encode+decode megapixels/sec of the hiccup DCT hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--images n --height h --width w]

A step = one full pass of the hot path over one batch of synthetic images: encode (K1 fused
colour/pyrDown/DCT/quantise/zigzag + float64 tie fix-up, E1 run-length symbols + histograms, E2 host
Huffman construction, E3 bit packing) followed by decode (D1 Huffman decode, D2 run-length expand,
D3 DC prefix sum, K7 dequantise/IDCT + fix-up, K8 pyrUp/colour).  Default workload: BASELINE.json
configs[1], a batch of 1024 synthetic 640x426 RGB images on one B200.  With N GPUs every rank
processes its own batch (images are independent; no data-path collective) -> weak scaling.

`value`  = whole-job MP/s with the batch resident in HBM when the timed region starts.
`e2e`    = the same metric through the public host API (DctBatchCodec.encode / .decode) with the
           batch in pinned host memory: host->device and device->host copies inside the timed region.
`--impl reference` times the reference's CPU path (the oracle port: the reference itself cannot
finish one 640x426 image's entropy stage in minutes) on all host cores.
"""
import argparse
import ctypes
import json
import os
import pickle
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the pipelined host-to-host path runs many CUDA streams; give them separate hardware queues
# (must be set before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "encode+decode megapixels/sec"
UNIT = "MP/s"


def synthetic_image(h, w, seed):
    """uint8 noise at 1/16 resolution, bicubic upsample, + N(0, 4), clipped."""
    import cv2
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, size=(max(2, h // 16), max(2, w // 16), 3), dtype=np.uint8)
    big = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC).astype(np.float32)
    big += rng.normal(0.0, 4.0, size=big.shape).astype(np.float32)
    return np.clip(np.rint(big), 0, 255).astype(np.uint8)


def synthetic_batch(n, h, w, seed0, out=None):
    out = np.empty((n, h, w, 3), np.uint8) if out is None else out
    for i in range(n):
        out[i] = synthetic_image(h, w, seed0 + i)
    return out


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                clk, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax = mx
            if t0 <= t <= t1:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:          # region shorter than the sampling period: use every sample we have
            for t, line in self.rows:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def pinned_array(lib_mod, shape, dtype=np.uint8):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    lib_mod.check(lib_mod.load().hic_host_alloc(ctypes.byref(p), nbytes))
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p


def kernel_bytes(name, P, B, s_ac, bits_bytes, rows, n_ss, bins, mode):
    """Algorithmic bytes one launch of `name` moves: P pixels, B 64-element
    blocks, s_ac run-length symbols, bits_bytes Huffman-coded payload bytes, rows code-table rows,
    n_ss symbol streams, bins histogram bins per stream."""
    coef = 128.0 * B
    n_dc = B if mode == "dct" else 0.0
    sym = 3.0 * s_ac + 2.0 * n_dc
    table = {
        "forward_kernel": 6.0 * P,                      # 3 B/pixel RGB in, 1.5 samples x int16 out
        "wavelet_forward_kernel": 9.0 * P,              # 3 B/pixel in, 3 samples x int16 out
        "wavelet_inverse_kernel": 9.0 * P,
        "rle_tile_summary_kernel": coef + 2.0 * n_dc,           # (whole images: the DC differences leave from here too)
        "rle_emit_kernel": coef + sym,
        "compact_kernel": 8.0 * n_ss * bins + 12.0 * rows,
        "huffman_sort_kernel": 20.0 * rows,
        "huffman_replay_kernel": 8.0 * rows,            # one span over the 21 concurrent tier launches
        "huffman_codes_kernel": 28.0 * rows,
        "pack_tile_bits_kernel": sym,
        "pack_emit_kernel": sym + bits_bytes,
        "build_tables_kernel": 12.0 * rows + 32768.0 * n_ss,
        "huffman_sync_kernel": bits_bytes,
        "huffman_resync_kernel": bits_bytes,
        "huffman_write_kernel": bits_bytes + sym,
        "expand_tile_sum_kernel": 1.0 * s_ac,
        "expand_scatter_kernel": 3.0 * s_ac + 2.0 * s_ac + 2.0 * n_dc,
        "dc_tile_sum_kernel": 2.0 * n_dc,
        "dc_prefix_kernel": 4.0 * n_dc,
        "inverse_kernel": coef + 1.5 * P,
        "upsample_colour_kernel": 4.5 * P,
    }
    return table.get(name)


CONFIGS = {
    # BASELINE.json configs[0..3]; c5 (16384^2 band-sharded over 8 GPUs) is bench_bands.py.
    # c1 is the reference's own CPU-runnable case: ONE 512x512 image through the eight drop-in functions (run_c1).
    "c1": dict(mode="dct", images=1, height=512, width=512, label="resources/Lenna.png 512x512 RGB, DCT mode encode->decode round trip through the drop-in functions (%d image per call)"),
    "c2": dict(mode="dct", images=1024, height=426, width=640, label="batch of %d synthetic 640x426 RGB images per GPU, DCT mode encode+decode"),
    "c3": dict(mode="dct", images=256, height=2160, width=3840, label="batch of %d synthetic 4K (3840x2160) RGB images per GPU, DCT + Huffman encode+decode"),
    "c4": dict(mode="wavelet", images=1, height=4320, width=7680, label="%d synthetic 8K (7680x4320) RGB image per GPU, wavelet mode encode+decode"),
}


def load_traffic(config):
    """{kernel: DRAM bytes per launch} from the newest profiles/rNN_traffic_<config>.json, and that file's name
    (None, None-ish when there is no capture for this workload: the numbers are then null, not stale)."""
    import glob
    if config is None:
        return {}, None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]_traffic_%s.json" % config)))
    if not files:
        return {}, None
    with open(files[-1]) as f:
        raw = json.load(f)
    traffic = {k: v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in raw.items() if isinstance(v, dict) and "dram_read_bytes" in v}
    if "upsample_colour_vec_kernel" in traffic:
        traffic["upsample_colour_kernel"] = traffic["upsample_colour_vec_kernel"]
    return traffic, "profiles/" + os.path.basename(files[-1]) + " (ncu --set full, one launch each; captured at the commit that added the file)"


def _oracle_round_trip(rgb, mode):
    from oracle import hiccup_oracle as orc
    if mode == "dct":
        return orc.jpeg_decompression(orc.jpeg_decode(orc.jpeg_encode(orc.jpeg_compression(rgb))))
    return orc.wavelet_decompression(orc.wavelet_decode(orc.wavelet_encode(orc.wavelet_compression(rgb))))


def cpu_baseline_sample(h, w, n_images, seed0, mode):
    """The oracle port (numpy restatement of the reference) on one host core."""
    imgs = [synthetic_image(h, w, seed0 + i) for i in range(n_images)]
    t0 = time.perf_counter()
    for rgb in imgs:
        _oracle_round_trip(rgb, mode)
    dt = time.perf_counter() - t0
    return n_images * h * w / 1e6 / dt, dt


def _ref_worker(args):
    h, w, seed, mode = args
    out = _oracle_round_trip(synthetic_image(h, w, seed), mode)
    return int(out[0, 0, 0])


def resolve_config(args):
    cfg = dict(CONFIGS[args.config])
    for key in ("images", "height", "width", "mode"):
        if getattr(args, key) is not None:
            cfg[key] = getattr(args, key)
    cfg["workload"] = cfg["label"] % cfg["images"] if "%d" in cfg["label"] else cfg["label"]
    if any(getattr(args, k) is not None for k in ("images", "height", "width", "mode")):
        cfg["workload"] = "%d synthetic %dx%d RGB image(s) per GPU, %s mode encode+decode" % (
            cfg["images"], cfg["width"], cfg["height"], cfg["mode"])
    return cfg


def cpu_sample_shape(h, w):
    """A bounded CPU sample: whole images up to ~1 MP, else a 1024x1024 corner crop (multiple of 16)."""
    if h * w <= (1 << 20):
        return h, w, "whole %dx%d images" % (w, h)
    ch, cw = min(h, 1024), min(w, 1024)
    return ch, cw, "%dx%d corner crops of the %dx%d images" % (cw, ch, w, h)


def run_reference(args):
    """Reference arm: the reference's CPU algorithm (oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cfg = resolve_config(args)
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)
    h, w, what = cpu_sample_shape(cfg["height"], cfg["width"])
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for i in range(args.warmup):
            pool.map(_ref_worker, [(h, w, 5000 + i * per_step + j, cfg["mode"]) for j in range(per_step)])
        t0 = time.perf_counter()
        for i in range(args.steps):
            pool.map(_ref_worker, [(h, w, 9000 + i * per_step + j, cfg["mode"]) for j in range(per_step)])
        dt = time.perf_counter() - t0
    value = args.steps * per_step * h * w / 1e6 / dt
    sample = "%d per step, %s, oracle port (numpy restatement of the reference), one process per core" % (per_step, what)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def run_c1(args):
    """BASELINE.json configs[0]: Lenna 512x512 through the reference's own call sequence (run.py:18-43):
    jpeg_compression -> jpeg_encode -> .hic bytes -> HicImage.from_bytes -> jpeg_decode -> jpeg_decompression,
    one image per call, host arrays in and out.  A latency workload: the line reports ms per call and per function.
    The image and the unmodified reference's outputs for it come from tests/golden/lenna512.npz (written by
    tests/golden/gen_golden.py from /root/reference/resources/Lenna.png; the reference needs 274 s for this)."""
    import pickle
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from hiccup_b200 import _lib, codec, compression, hicimage
    from hiccup_b200.batch import DctBatchCodec
    _lib.check(_lib.load().hic_set_device(local_rank))
    g = np.load(os.path.join(ROOT, "tests", "golden", "lenna512.npz"))
    rgb, want_hic, want_out = np.ascontiguousarray(g["rgb"]), pickle.loads(g["hic"].tobytes()), g["rgb_out"]
    h, w = rgb.shape[:2]
    names = ("jpeg_compression", "jpeg_encode", "to_bytes", "from_bytes", "jpeg_decode", "jpeg_decompression")
    per_fn = {k: 0.0 for k in names}

    def call(timed):
        t = [time.perf_counter()]
        comp = compression.jpeg_compression(rgb); t.append(time.perf_counter())
        hic = codec.jpeg_encode(comp); t.append(time.perf_counter())
        stream = hic.byte_stream(); t.append(time.perf_counter())
        back = hicimage.HicImage.from_bytes(stream); t.append(time.perf_counter())
        planes = codec.jpeg_decode(back); t.append(time.perf_counter())
        out = compression.jpeg_decompression(planes); t.append(time.perf_counter())
        if timed:
            for k, a, b in zip(names, t, t[1:]):
                per_fn[k] += (b - a) * 1e3
        return stream, out

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        stream, out = call(False)
    same_bytes = stream == want_hic
    same_pixels = bool(np.array_equal(out, want_out))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    t_start = time.time()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call(True)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    barrier()
    # device-resident arm: the same image as a batch of one on the batched codec (K1 .. K8 + entropy, no copies)
    bc = DctBatchCodec(1, h, w)
    bc.upload(rgb[None])
    for _ in range(args.warmup):
        bc.encode_device(); bc.decode_device()
    _lib.sync()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    pairs = []
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        bc.encode_device(); bc.decode_device()
        eb.record()
        pairs.append((ea, eb))
    torch.cuda.synchronize()
    ms_dev = sum(a.elapsed_time(b) for a, b in pairs) / args.steps
    _lib.profile_enable(True)
    _lib.profile_report()
    for _ in range(args.steps):
        bc.encode_device(); bc.decode_device()
    _lib.sync()
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end)
    if dist is not None:
        t = torch.tensor([ms_e2e, ms_dev], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e, ms_dev = float(t[0].item()), float(t[1].item())
    peak, peak_src = measured_peak()
    k1 = prof.get("forward_kernel", (0.0, 1))
    k1_ms = k1[0] / max(k1[1], 1)
    dominant = max(prof.items(), key=lambda kv: kv[1][0])[0]
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        _oracle_round_trip(rgb, "dct")
        dt = time.perf_counter() - t0
        cpu = {"value": h * w / 1e6 / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "this one image, full encode+decode through the oracle port (%.2f s); the unmodified reference needs 274 s for it (BASELINE.md)" % dt}
    if rank == 0:
        mp = h * w / 1e6
        line = {
            "metric": METRIC, "value": world * mp / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "Lenna (tests/golden/lenna512.npz)",
            "config": {"workload": CONFIGS["c1"]["label"] % 1, "mode": "dct", "images_per_gpu": 1, "height": h, "width": w,
                       "parallelism": "one image per call per GPU, %d GPU(s), no collective" % world,
                       "l2": "L2 flushed (256 MB device fill) before every timed device-resident step",
                       "note": "a latency workload: 0.26 MP per call cannot fill a B200; the per-call floor is launch latency and host work"},
            "roofline": {"kernel": dominant, "bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                         "peak_source": peak_src, "note": "launch-latency-bound at this size; see north_star_kernel",
                         "north_star_kernel": {"kernel": "forward_kernel", "ms_per_launch": round(k1_ms, 4), "algorithmic_bytes": 6.0 * h * w,
                                               "achieved": round(6.0 * h * w / k1_ms / 1e6, 1) if k1_ms else None, "peak": peak, "unit": "GB/s",
                                               "frac": round(6.0 * h * w / k1_ms / 1e6 / peak, 4) if k1_ms else None}},
            "kernels": {k: {"ms_per_launch": round(v[0] / max(v[1], 1), 4), "launches_per_step": v[1] / args.steps} for k, v in prof.items()},
            "cpu_baseline": cpu, "clocks": clocks,
            "e2e": {"value": world * mp / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(rgb.nbytes + 4 * 1.5 * h * w * 2),
                    "d2h_bytes_per_step": int(4 * 1.5 * h * w * 2 + sum(len(b) for b in stream) + out.nbytes),
                    "api": "compression.jpeg_compression -> codec.jpeg_encode -> HicImage.byte_stream -> HicImage.from_bytes -> codec.jpeg_decode -> "
                           "compression.jpeg_decompression, numpy arrays in and out, one call after another",
                    "ms_per_function": {k: round(v / args.steps, 3) for k, v in per_fn.items()}},
            "gpu_launches": int(sum(v[1] for v in prof.values())),
            "parity": {"checked_images": 1, "mismatches": int(not (same_bytes and same_pixels)),
                       "against": "the UNMODIFIED reference's .hic payloads and decoded pixels for this image (tests/golden/lenna512.npz)",
                       "hic_bytes_equal": bool(same_bytes), "pixels_equal": same_pixels},
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not (same_bytes and same_pixels):
        sys.stderr.write("bench.py: PARITY FAILURE on Lenna\n")
        return 3
    return 0


def run_e2e(args, cfg, codec, pipe, host_rgb, rank, local_rank, world, dist, barrier):
    """The host-to-host arm.  The public batched call is PipelinedCodec: chunks of the batch flow through
    concurrent slots so H2D, kernels and D2H overlap; inputs and outputs are page-locked host arrays and every
    step copies the whole batch in and the compressed streams, code tables and decoded pixels out.

    One GPU: PipelinedCodec.round_trip(host_rgb, host_out).  Several GPUs: the ranks' batches form ONE job
    in shared page-locked host memory (hiccup_b200/jobs.py) and every rank's pipeline takes chunks from one
    ticket counter (PipelinedCodec.run_job) -- by-image partition with the share of each GPU decided by how
    fast its path to host memory turns out to be, because on a multi-GPU box those paths are not alike.

    Beside the real thing the same call runs with copy_only=True (identical copies, threads and gates, no
    kernels): `copy_floor_ms` is what PCIe and host memory allow this very pattern on this box at this N."""
    import torch
    from hiccup_b200 import _lib
    from hiccup_b200.batch import PipelinedCodec
    from hiccup_b200.jobs import SharedJob
    mode = cfg["mode"]
    n, h, w = cfg["images"], cfg["height"], cfg["width"]
    pixels = n * h * w
    steps = args.steps
    chunk = pipe.chunk
    table_bytes = [0]

    def on_encoded(first, e):
        table_bytes[0] += int(e.symbols.nbytes + e.packed.nbytes + e.index.nbytes)

    def max_over_ranks(*vals):
        if dist is None:
            return vals
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return tuple(float(v) for v in t.tolist())

    def sum_over_ranks(*vals):
        if dist is None:
            return vals
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return tuple(float(v) for v in t.tolist())

    def timed(fn):
        barrier()
        t = time.perf_counter()
        out = fn()
        _lib.sync()
        ms = (time.perf_counter() - t) * 1e3          # this rank's wall clock from the common start to its own end
        barrier()
        return ms, out

    job = None
    extra = {}
    if world > 1:
        job = SharedJob("bench%s" % os.environ.get("MASTER_PORT", "0"), rank, world, n, h, w, pipe.out_shape[1], pipe.out_shape[2],
                        barrier=dist.barrier)
        job.inputs[rank][...] = host_rgb
        host_in, host_out = job.inputs[rank], job.outputs[rank]
        for _ in range(2):
            job.reset()
            pipe.run_job(job)
        job.reset()
        ms, (payload, chunks_mine) = timed(lambda: pipe.run_job(job, on_encoded=on_encoded, repeat=steps))
        job.reset()
        ms_floor, _ = timed(lambda: pipe.run_job(job, repeat=steps, copy_only=True))
        # the fixed partition of round 1 (every rank its own 1/N of the images), and its floor
        for _ in range(2):
            pipe.round_trip(host_in, host_out)
        ms_static, _ = timed(lambda: pipe.round_trip(host_in, host_out, repeat=steps))
        ms_static_floor, _ = timed(lambda: pipe.round_trip(host_in, host_out, repeat=steps, copy_only=True))
        ms, ms_floor, ms_static, ms_static_floor = max_over_ranks(ms, ms_floor, ms_static, ms_static_floor)
        payload_job, tables_job = sum_over_ranks(float(payload), float(table_bytes[0]))
        shares = [0.0] * world
        shares[rank] = float(chunks_mine)
        shares = [int(v) for v in sum_over_ranks(*shares)]
        per_step_chunks = world * (n // chunk)
        h2d = int(world * host_rgb.nbytes)
        d2h = int(payload_job + tables_job / steps + world * host_out.nbytes)
        api = ("PipelinedCodec.run_job(SharedJob, repeat=steps): the %d ranks' batches are one job in shared page-locked host "
               "memory; chunks of %d images are taken from one ticket counter by whichever rank has a free slot (8 slots per "
               "rank), steps streamed back to back; each chunk is decoded from the payload its encoder left on the device "
               "(downloaded as a result, not uploaded again)" % (world, chunk))
        extra = {"chunks_taken_per_rank": shares, "chunks_per_step": per_step_chunks,
                 "static_partition": {"ms_per_step": ms_static / steps, "value": world * pixels / 1e6 / (ms_static / steps / 1e3),
                                      "copy_floor_ms": ms_static_floor / steps,
                                      "what": "round 1's fixed share per rank (PipelinedCodec.round_trip on its own batch)"}}
        ms_drained = None
    else:
        host_out, _keep_out = pinned_array(_lib, pipe.out_shape)
        host_in = host_rgb
        for _ in range(2):
            pipe.round_trip(host_in, host_out)
        # (a) one batch per call, pipeline drained between steps
        ms_drained, _ = timed(lambda: [pipe.round_trip(host_in, host_out) for _ in range(steps)])
        # (b) the K steps as a stream of batches through the same pipeline (no drain in between); every step
        # still copies its whole batch in and its streams, tables and pixels out
        ms, payload = timed(lambda: pipe.round_trip(host_in, host_out, on_encoded=on_encoded, repeat=steps))
        ms_floor, _ = timed(lambda: pipe.round_trip(host_in, host_out, repeat=steps, copy_only=True))
        # (c) the same stream of batches with every chunk decoded from the HOST copy of its payload and tables (uploaded
        # again, as decoding a file would): what the device-resident hand-over of (b) saves
        pipe.round_trip(host_in, host_out, from_device=False)
        ms_reupload, _ = timed(lambda: pipe.round_trip(host_in, host_out, repeat=steps, from_device=False))
        extra = {"decode_from_host_copy": {"ms_per_step": ms_reupload / steps, "value": pixels / 1e6 / (ms_reupload / steps / 1e3),
                                           "what": "round_trip(from_device=False): each chunk's compressed payload and code tables go "
                                                   "back up over PCIe before its decode"}}
        h2d = int(host_rgb.nbytes)
        d2h = int(payload + table_bytes[0] // max(steps, 1) + host_out.nbytes)
        api = ("PipelinedCodec.round_trip(repeat=steps): chunks of %d images over 8 slots, steps streamed back to back; each "
               "chunk is decoded from the compressed payload the encoder left on the device (downloaded to the host as a "
               "result, not uploaded again)" % chunk)
    e2e_value = world * pixels / 1e6 / (ms / steps / 1e3)

    # ---- parity of what the timed path produced (outside the timed region) ----
    # (1) the pipelined / job path against the unchunked codec on this rank's whole batch
    enc_res = codec.encode(host_rgb)
    codec.decode(enc_res)
    mine = codec._h_out.array(np.uint8)[:host_out.size]
    e2e_same = bool(np.array_equal(host_out.reshape(-1), mine))
    # (2) k images of THIS batch against the oracle (the CPU restatement of the reference): code tables, framed
    # bit strings and decoded pixels, bit for bit
    parity = oracle_parity(codec, enc_res, mine.reshape(pipe.out_shape), host_rgb, mode, k=8 if pixels / n <= (1 << 20) else 0,
                           seed=rank)
    if mode == "dct":
        parity.update({"forward_ties": {"flagged_blocks": int(codec.forward_stats[0]), "reevaluated": int(codec.forward_stats[1]),
                                        "changed": int(codec.forward_stats[2])},
                       "inverse_ties": {"flagged_blocks": int(codec.inverse_stats[0]), "reevaluated": int(codec.inverse_stats[1]),
                                        "changed": int(codec.inverse_stats[2])}})
    # (3) what `e2e` stops short of: the `.hic` container objects and their byte strings (one pickle per table row,
    # as the format demands -- host Python, one core, a sample of this batch's images; not in the timed region)
    sample = list(range(0, n, max(1, n // 32)))[:32]
    t0 = time.perf_counter()
    hic_bytes = sum(sum(len(b) for b in hi.byte_stream()) for hi in codec.hic_images(enc_res, images=sample))
    hic_ms = (time.perf_counter() - t0) * 1e3 / len(sample)
    hic_level = {"ms_per_image": round(hic_ms, 3), "images": len(sample), "cores": 1, "bytes_per_image": hic_bytes // len(sample),
                 "what": "batch codec hic_images() + HicImage.byte_stream() on the host: the step from the packed tables and framed "
                         "bit strings `e2e` delivers to the reference's list of pickled payloads (what write_file() dumps)"}
    # (4) the same step for the WHOLE batch through the library's host threads (batch.hic_files / streams_from_files:
    # files byte-identical to (3)'s, checked on the sample), and the files -> pixels direction from those bytes
    try:
        threads = min(len(os.sched_getaffinity(0)), 32)
        files = codec.hic_files(enc_res, threads=threads, reuse=True)
        t0 = time.perf_counter()
        files = codec.hic_files(enc_res, threads=threads, reuse=True)          # (into the codec's own buffer, like its other staging)
        write_ms = (time.perf_counter() - t0) * 1e3
        same_files = all(bytes(files[i]) == pickle.dumps(hi.byte_stream()) for i, hi in zip(sample[:4], codec.hic_images(enc_res, images=sample[:4])))
        back = codec.streams_from_files(files, threads=threads)
        t0 = time.perf_counter()
        back = codec.streams_from_files(files, threads=threads)
        read_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        codec.decode(back)
        decode_ms = (time.perf_counter() - t0) * 1e3
        same_pixels = bool(np.array_equal(codec._h_out.array(np.uint8)[:host_out.size], host_out.reshape(-1)))
        hic_level["batch_files"] = {
            "images": n, "host_threads": threads, "write_ms_per_batch": round(write_ms, 2), "read_ms_per_batch": round(read_ms, 2),
            "write_ms_per_image": round(write_ms / n, 4), "read_ms_per_image": round(read_ms / n, 4),
            "file_bytes_per_batch": int(sum(len(f) for f in files)), "files_equal_python_container": bool(same_files),
            "decode_of_read_files_ms": round(decode_ms, 2), "decoded_pixels_equal": same_pixels,
            "what": "hic_hicfile_pack_files / scan_files + parse_files: every image's whole `.hic` file written into / parsed from "
                    "host memory by the library's threads, one call per batch (this rank's batch; not in the timed region)"}
        del files, back
    except Exception as exc:                                 # a reported figure, never a reason to lose the line
        hic_level["batch_files"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": ms / steps, "api": api,
           "copy_floor_ms": ms_floor / steps, "frac_of_copy_floor": (ms_floor / ms) if ms else None,
           "copy_floor_how": "the same call with copy_only=True: identical bulk copies, slot threads and gates, no kernels, all ranks at once",
           "matches_unchunked": e2e_same, "hic_container": hic_level}
    if ms_drained is not None:
        e2e["drained_value"] = world * pixels / 1e6 / (ms_drained / steps / 1e3)
        e2e["drained_ms_per_step"] = ms_drained / steps
    e2e.update(extra)
    if job is not None:
        job.close()
    pipe.close()
    return e2e, e2e_same, parity


def oracle_parity(codec, enc, decoded, host_rgb, mode, k, seed):
    """k random images of the benchmark's own batch through the oracle: every code table, every framed bit
    string and every decoded pixel must be equal.  Returns {"checked_images", "mismatches", ...}."""
    out = {"checked_images": 0, "mismatches": 0, "against": "oracle port (numpy restatement of the reference), bit-exact"}
    if k <= 0:
        out["note"] = "images above 1 MP: the oracle comparison of whole images lives in tests/test_gpu_full_size.py"
        return out
    from oracle import hiccup_oracle as orc
    rng = np.random.default_rng(77 + seed)
    n = host_rgb.shape[0]
    picks = sorted(set(int(i) for i in rng.choice(n, size=min(k, n), replace=False)))
    bad = []
    for i in picks:
        rgb = np.array(host_rgb[i])
        if mode == "dct":
            planes = orc.jpeg_compression(rgb)
            want = orc.jpeg_encode(planes)
            pixels = orc.jpeg_decompression(orc.jpeg_decode(want))
            kinds = 3
        else:
            planes = orc.wavelet_compression(rgb)
            want = orc.wavelet_encode(planes)
            pixels = orc.wavelet_decompression(orc.wavelet_decode(want))
            kinds = 2
        ok = True
        for kind in range(kinds):
            for c in range(3):
                s = (i * 3 + c) * 3 + (kind if mode == "dct" else kind + 1)
                j = kind * 3 + c
                if enc.framed(s) != orc.padded_bits_to_bytes(want["bits"][j]):
                    ok = False
                if [(int(a), b) for a, b in enc.table(s)] != [(int(a), b) for a, b in want["tables"][j]]:
                    ok = False
        if not np.array_equal(decoded[i], pixels):
            ok = False
        if not ok:
            bad.append(i)
    out["checked_images"] = len(picks)
    out["mismatches"] = len(bad)
    out["images"] = picks
    if bad:
        out["mismatched_images"] = bad
    return out



_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: anything libraries print (NCCL's version banner, ...) goes to
    stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default=None, choices=["dct", "wavelet"])
    ap.add_argument("--images", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "c1" and not any(getattr(args, k) is not None for k in ("images", "height", "width", "mode")):
        return run_c1(args)
    cfg = resolve_config(args)
    mode = cfg["mode"]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from hiccup_b200 import _lib
    from hiccup_b200.batch import DctBatchCodec, WaveletBatchCodec
    _lib.check(_lib.load().hic_set_device(local_rank))
    n, h, w = cfg["images"], cfg["height"], cfg["width"]
    codec = (DctBatchCodec if mode == "dct" else WaveletBatchCodec)(n, h, w, stream=None)
    host_rgb, _keep_in = pinned_array(_lib, (n, h, w, 3))
    distinct = n if n * h * w <= 300e6 else min(n, 16)       # big batches repeat a few distinct images
    synthetic_batch(distinct, h, w, 1000 * 2 + rank * n, out=host_rgb[:distinct])
    for i in range(distinct, n):
        host_rgb[i] = host_rgb[i % distinct]
    codec.upload(host_rgb)
    _lib.sync()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        codec.encode_device()
        codec.decode_device()

    # The batch as eight chunks on eight CUDA streams (PipelinedCodec, also the host-to-host arm's object): one
    # chunk's latency-bound stretches -- the serial heapq replays of the Huffman builder, table builds, stream
    # scans -- run under the other chunks' bandwidth-bound kernels.  Same kernels, same results as the single
    # codec (images are independent), everything resident in HBM.
    from hiccup_b200.batch import PipelinedCodec
    n_slots = 8 if (n >= 8 and n % 8 == 0) else 1
    pipe = PipelinedCodec(n, h, w, chunk=n // n_slots, slots=n_slots, mode=mode, device=local_rank)
    overlapped = n_slots > 1
    # host threads that drive the slots: one per slot when this rank has the cores for it (the slots' host threads
    # spin while they wait for the device), fewer when the ranks of the box share them
    slot_threads = max(1, min(n_slots, (os.cpu_count() or n_slots) // max(world, 1) - 1))
    if overlapped:
        pipe.upload_resident(host_rgb)

    # ---- device-resident arm (value) -------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    if overlapped:
        pipe.device_steps(args.warmup, threads=slot_threads)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    # Working sets below ~2x L2 (126 MB) get an L2 flush (a 256 MB device fill) before every timed step,
    # timed step by step so the flush itself stays outside the measurement.
    need_flush = host_rgb.nbytes < 256e6
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if need_flush else None

    def timed_overlapped():
        # the K steps streamed through the eight slots; the events bracket everything because device_steps()
        # returns only after every slot's stream has drained
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        t0 = time.perf_counter()
        pipe.device_steps(args.steps, threads=slot_threads)
        e1.record()
        barrier()
        return max(e0.elapsed_time(e1), 0.0), (time.perf_counter() - t0) * 1e3

    def timed_steps():
        if need_flush:
            pairs = []
            for i in range(args.steps):
                flush.fill_(i & 0xFF)
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
                step_device()
                eb.record()
                pairs.append((ea, eb))
            barrier()
            return sum(a.elapsed_time(b) for a, b in pairs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_device()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    barrier()
    t_start = time.time()
    ms_single = timed_steps()                    # one codec, one stream: what round 1 reported
    ms_overlapped = None
    if overlapped and not need_flush:
        barrier()
        ms_overlapped, _wall = timed_overlapped()
    t_end = time.time()
    clocks = sampler.stop(t_start, t_end)
    # Per-kernel times (the `kernels` table and `roofline`): the same K steps once more with CUDA events
    # around every kernel (hic_profile_*) and the encoder's DC Huffman pass kept on the main stream
    # (HIC_ENTROPY_SERIAL) -- in the timed steps above it runs beside rle_emit on other streams, where an
    # event span around either kernel would also count the time it waited for the other.
    os.environ["HIC_ENTROPY_SERIAL"] = "1"
    step_device()
    barrier()
    _lib.profile_enable(True)
    _lib.profile_report()
    ms_serial = timed_steps()
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    del os.environ["HIC_ENTROPY_SERIAL"]
    if dist is not None:
        t = torch.tensor([ms_overlapped if ms_overlapped is not None else 0.0, ms_single], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_overlapped, ms_single = (float(t[0].item()) if ms_overlapped is not None else None), float(t[1].item())
    # both arms ran the same K steps under the same rules (max over ranks each): the step is the faster one
    use_overlapped = ms_overlapped is not None and ms_overlapped < ms_single
    ms_total = ms_overlapped if use_overlapped else ms_single
    ms_step = ms_total / args.steps
    pixels = n * h * w
    value = world * pixels / 1e6 / (ms_step / 1e3)

    # ---- roofline of the dominant kernel ---------------------------------------------------
    enc = codec.encoder
    s_ac = float(sum(int(enc.nsym[s]) for s in range(1, len(enc.nsym), 3)))
    bits_bytes = float(enc.total_bytes)
    B = float(codec.layout.n_images * codec.layout.blocks_per_image)
    rows, n_ss, bins = float(enc.total_rows), float(enc.n_streams), float(enc.value_bins)
    peak, peak_src = measured_peak()
    kernels = {}
    for name, (ms, launches) in prof.items():
        per = ms / max(launches, 1)
        ab = kernel_bytes(name, float(pixels), B, s_ac, bits_bytes, rows, n_ss, bins, mode)
        kernels[name] = {"ms_per_launch": round(per, 4), "launches_per_step": launches / args.steps,
                         "share_of_step": round(ms / ms_serial, 4) if ms_serial else None,
                         "algorithmic_gbs": round(ab / per / 1e6, 1) if ab and per > 0 else None,
                         "frac_of_peak": round(ab / per / 1e6 / peak, 4) if ab and per > 0 else None}
    dominant = max(prof.items(), key=lambda kv: kv[1][0])[0] if prof else None
    # DRAM traffic per launch: not measurable inside this run (it needs ncu), so it is quoted from the newest
    # committed `ncu --set full` capture of this exact workload, and the line says which file that was
    traffic, traffic_source = load_traffic(args.config if (n, h, w, mode) == tuple(CONFIGS[args.config][k] for k in ("images", "height", "width", "mode")) else None)
    for name in kernels:
        kernels[name]["dram_traffic_bytes"] = traffic.get(name)
    roofline = None
    if dominant:
        k = kernels[dominant]
        roofline = {"kernel": dominant, "bound": "hbm", "achieved": k["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": k["frac_of_peak"], "traffic": traffic.get(dominant), "peak_source": peak_src,
                    "share_of_step": k["share_of_step"],
                    "algorithmic_bytes": kernel_bytes(dominant, float(pixels), B, s_ac, bits_bytes, rows, n_ss, bins, mode),
                    "note": "latency-bound serial heapq replay (DESIGN.md section 5); the HBM-bound kernels are listed under `kernels`"
                            if dominant == "huffman_replay_kernel" else None}
    # the same for the largest kernel that actually moves bytes (the replay's bytes are negligible)
    roofline_hbm = None
    movers = {k: v for k, v in kernels.items() if v["frac_of_peak"] is not None and k not in
              ("huffman_replay_kernel", "huffman_sort_kernel", "huffman_codes_kernel", "build_tables_kernel")}
    if movers:
        name = max(movers, key=lambda k: movers[k]["ms_per_launch"] * movers[k]["launches_per_step"])
        k = movers[name]
        roofline_hbm = {"kernel": name, "bound": "hbm", "achieved": k["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": k["frac_of_peak"], "traffic": traffic.get(name), "share_of_step": k["share_of_step"],
                        "algorithmic_bytes": kernel_bytes(name, float(pixels), B, s_ac, bits_bytes, rows, n_ss, bins, mode)}
    # K1, the kernel north_star sets the 60 % target for, rides inside `roofline` whatever the dominant kernel is
    if roofline is not None:
        roofline["traffic_source"] = traffic_source
        k1_name = "forward_kernel" if mode == "dct" else "wavelet_forward_kernel"
        if k1_name in kernels:
            k = kernels[k1_name]
            roofline["north_star_kernel"] = {
                "kernel": k1_name, "what": "fused colour + pyrDown + DCT + quantise + zigzag (K1)" if mode == "dct" else "fused colour + 3-level db1 + quantise + zigzag (K9)",
                "bound": "hbm", "achieved": k["algorithmic_gbs"], "peak": peak, "unit": "GB/s", "frac": k["frac_of_peak"],
                "ms_per_launch": k["ms_per_launch"], "traffic": traffic.get(k1_name),
                "algorithmic_bytes": kernel_bytes(k1_name, float(pixels), B, s_ac, bits_bytes, rows, n_ss, bins, mode),
                "target_frac": 0.60}
    # the replay span stands for 21 tier launches
    gpu_launches = int(sum(v[1] for v in prof.values()) + 20 * prof.get("huffman_replay_kernel", (0, 0))[1])

    # ---- end-to-end arm (host buffers, copies inside the timed region) ---------------------
    e2e, e2e_same, parity = run_e2e(args, cfg, codec, pipe, host_rgb, rank, local_rank, world, dist, barrier)

    # a parity failure on any rank fails the run (after the line is printed, so that it can be read)
    bad = (0 if e2e_same else 1) + int(parity.get("mismatches", 0))
    if dist is not None:
        t = torch.tensor([float(bad), float(parity.get("checked_images", 0))], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        bad = int(t[0].item())
        parity["checked_images_all_ranks"] = int(t[1].item())
        parity["failures_all_ranks"] = bad

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ch_, cw_, what = cpu_sample_shape(h, w)
        n_cpu = max(2, min(64, int(17.5e6 // (ch_ * cw_))))      # ~10 s of single-core work
        v, dt = cpu_baseline_sample(ch_, cw_, n_cpu, 2000, mode)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d x %s of this workload, full encode+decode through the oracle port "
                         "(numpy restatement of the reference, %.1f s)" % (n_cpu, what, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "mode": mode, "images_per_gpu": n, "distinct_images": distinct, "height": h, "width": w, "parallelism": "by image, %d GPU(s), no collective" % world,
                       "l2": ("L2 flushed (256 MB device fill) before every timed step; steps timed individually"
                              if need_flush else "inputs larger than L2 (%.0f MB RGB per batch), no flush" % (host_rgb.nbytes / 1e6)),
                       "compressed_bytes_per_batch": int(enc.total_bytes), "symbols_per_batch": int(s_ac),
                       "step": ("the batch as %d chunks of %d images on %d CUDA streams driven by %d host threads (PipelinedCodec.device_steps), "
                                "the K steps streamed back to back; a single codec on one stream takes single_stream_ms_per_step"
                                % (n_slots, n // n_slots, n_slots, slot_threads))
                               if use_overlapped else "one codec on one CUDA stream (the faster of the two arms timed in this run)",
                       "single_stream_ms_per_step": ms_single / args.steps,
                       "multi_stream_ms_per_step": (ms_overlapped / args.steps) if ms_overlapped is not None else None},
            "roofline": roofline, "roofline_largest_hbm_kernel": roofline_hbm, "kernels": kernels,
            "kernel_timing": {"how": "CUDA events around every kernel over %d extra steps with the DC Huffman pass serialised "
                                     "(HIC_ENTROPY_SERIAL); shares are of that pass" % args.steps,
                              "serialised_ms_per_step": ms_serial / args.steps},
            "cpu_baseline": cpu, "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": gpu_launches, "parity": parity,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if bad:
        sys.stderr.write("bench.py: PARITY FAILURE -- %s\n" % json.dumps(parity))
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())

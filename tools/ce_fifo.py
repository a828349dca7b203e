"""Latency of a small host->device copy (and of a small memset) while another stream streams bulk copies:
does splitting the bulk copy into pieces let the small one through? (development aid)"""
import torch, time, threading
big = 105 << 20
h_big = torch.empty(big, dtype=torch.uint8).pin_memory()
d_big = torch.empty(big, dtype=torch.uint8, device="cuda")
h_small = torch.empty(64 << 10, dtype=torch.uint8).pin_memory()
d_small = torch.empty(64 << 10, dtype=torch.uint8, device="cuda")
d_set = torch.empty(37 << 20, dtype=torch.uint8, device="cuda")
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
stop = False
def bulk(piece):
    while not stop:
        with torch.cuda.stream(sa):
            for off in range(0, big, piece):
                d_big[off:off + piece].copy_(h_big[off:off + piece], non_blocking=True)
        sa.synchronize()
for piece in (big, 4 << 20, 1 << 20):
    stop = False
    th = threading.Thread(target=bulk, args=(piece,)); th.start()
    time.sleep(0.05)
    lat, lat2, lat3 = [], [], []
    for _ in range(40):
        with torch.cuda.stream(sb):
            t = time.perf_counter(); d_small.copy_(h_small, non_blocking=True); sb.synchronize(); lat.append(time.perf_counter() - t)
            t = time.perf_counter(); d_set.zero_(); sb.synchronize(); lat2.append(time.perf_counter() - t)
            t = time.perf_counter(); h_small.copy_(d_small, non_blocking=True); sb.synchronize(); lat3.append(time.perf_counter() - t)
    stop = True; th.join()
    f = lambda v: "median %.3f ms, max %.3f ms" % (sorted(v)[len(v) // 2] * 1e3, max(v) * 1e3)
    print("bulk H2D in pieces of %d MB: small H2D %s | 37 MB zero_ %s | small D2H %s" % (piece >> 20, f(lat), f(lat2), f(lat3)))

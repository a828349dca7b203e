"""The copy pattern of PipelinedCodec.round_trip without any kernels: what does PCIe alone allow? (development aid)"""
import torch, time, sys
n_chunks, reps = 8, 5
big, small = 128 * 426 * 640 * 3, 20 << 20
h_in = torch.empty(n_chunks * big, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_chunks * big, dtype=torch.uint8).pin_memory()
h_bits = torch.empty(n_chunks * small, dtype=torch.uint8).pin_memory()
streams = [torch.cuda.Stream() for _ in range(n_chunks)]
d_in = [torch.empty(big, dtype=torch.uint8, device="cuda") for _ in range(n_chunks)]
d_out = [torch.empty(big, dtype=torch.uint8, device="cuda") for _ in range(n_chunks)]
d_bits = [torch.empty(small, dtype=torch.uint8, device="cuda") for _ in range(n_chunks)]
def run(order):
    torch.cuda.synchronize(); t = time.perf_counter()
    for r in range(reps):
        for c in range(n_chunks):
            with torch.cuda.stream(streams[c]):
                for op in order:
                    if op == "I": d_in[c].copy_(h_in[c * big:(c + 1) * big], non_blocking=True)
                    if op == "b": h_bits[c * small:(c + 1) * small].copy_(d_bits[c], non_blocking=True)
                    if op == "B": d_bits[c].copy_(h_bits[c * small:(c + 1) * small], non_blocking=True)
                    if op == "O": h_out[c * big:(c + 1) * big].copy_(d_out[c], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e3
for order in ("IbBO", "IO", "I", "O"):
    run(order)
    print(order, "%.1f ms per batch" % run(order))

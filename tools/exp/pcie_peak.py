"""Measured host<->device copy bandwidth of the box (pinned memory): one direction alone and both at once.
The end-to-end leg of bench.py moves 103 MB in and 103 MB out per step, so this is its ceiling."""
import time
import torch

dev = torch.device("cuda")
n = 100 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9


run(True, True, 3)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both at once %.1f GB/s per direction" % (run(True, False), run(False, True), run(True, True)))

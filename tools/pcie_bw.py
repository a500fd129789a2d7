"""Host <-> device copy bandwidth of this box (pinned host memory, 2 GiB blocks): each direction alone and both at once on two streams.
The e2e leg of bench.py moves the whole state both ways every step; this is the ceiling it is measured against."""
import time
import torch
n = 2 << 30
h0 = torch.empty(n, dtype=torch.uint8).pin_memory()
h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
d0 = torch.empty(n, dtype=torch.uint8, device="cuda")
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s0):
                d0.copy_(h0, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s1):
                h1.copy_(d1, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t) / 1e9
run(True, True, 1)
print("H2D alone %.1f GB/s   D2H alone %.1f GB/s   both at once %.1f GB/s per direction" % (run(True, False), run(False, True), run(True, True)))

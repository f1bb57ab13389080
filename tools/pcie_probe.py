"""Host<->device copy bandwidth of the box (pinned memory): H2D, D2H, and both directions at once.
The end-to-end (`e2e`) arm of bench.py moves the whole ensemble both ways every step; this is its ceiling."""
import time
import torch

n = 100 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n / (time.perf_counter() - t0) / 1e9


run(True, True, 2)
print("H2D GB/s %.1f" % run(True, False))
print("D2H GB/s %.1f" % run(False, True))
print("both, per direction GB/s %.1f" % run(True, True))

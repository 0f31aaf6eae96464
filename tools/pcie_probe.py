#!/usr/bin/env python3
"""Host<->device copy ceilings of the box (pinned memory), to put the end-to-end number in context."""
import time
import torch

n = 128 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("both directions at once", both, 2 * n)):
    t = timeit(fn)
    print("%-26s %.1f GB/s (%.3f ms for %d MiB)" % (name, nbytes / t / 1e9, t * 1e3, nbytes >> 20))

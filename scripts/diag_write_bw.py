#!/usr/bin/env python
"""Write-only / read-only / copy bandwidth of the box (torch fill_, sum, copy_ on 4 GiB), for judging the
write-bound first layer against what the memory system gives a pure writer."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("fill  (write only): %.0f GB/s" % (n * 4 / timed(lambda: a.fill_(1.0)) / 1e6))
print("sum   (read only) : %.0f GB/s" % (n * 4 / timed(lambda: a.sum()) / 1e6))
print("copy  (read+write): %.0f GB/s" % (2 * n * 4 / timed(lambda: b.copy_(a)) / 1e6))

"""HBM bandwidth by read/write mix on this GPU (library kernels, for context): pure write (fill), copy (1:1), pure read (sum).
usage: python tools/probes/hbm_mix.py"""
import torch

n = 1 << 30                                      # 1 GiB: far beyond the 126 MB L2
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
af = a.view(torch.float32)


def timed(fn, nbytes, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return nbytes / ms / 1e6


print(f"fill  (write only)      {timed(lambda: a.fill_(7), n):8.1f} GB/s")
print(f"copy  (read + write)    {timed(lambda: b.copy_(a), 2 * n):8.1f} GB/s")
print(f"sum   (read only)       {timed(lambda: af.sum(), n):8.1f} GB/s")

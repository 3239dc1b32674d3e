#!/usr/bin/env python
"""Measures plain HBM write / read / copy bandwidth on the box with torch, for context next to
MEASURED_PEAKS.json (which is a copy figure)."""
import torch
dev = torch.device("cuda:0")
x = torch.empty(1 << 30, dtype=torch.int64, device=dev)   # 8 GiB


def timed(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: x.fill_(7))
print(f"write 8 GiB fill_: {ms * 1e3:.1f} us  {x.numel() * 8 / ms / 1e6:.1f} GB/s")
ms = timed(lambda: x.zero_())
print(f"write 8 GiB zero_: {ms * 1e3:.1f} us  {x.numel() * 8 / ms / 1e6:.1f} GB/s")
y = x.view(torch.uint8)[: 4 << 30].view(torch.int32)
ms = timed(lambda: y.sum())
print(f"read 4 GiB sum(int32): {ms * 1e3:.1f} us  {y.numel() * 4 / ms / 1e6:.1f} GB/s")
z = torch.empty_like(x[: 1 << 29])
ms = timed(lambda: z.copy_(x[: 1 << 29]))
print(f"copy 4 GiB -> 4 GiB: {ms * 1e3:.1f} us  {2 * z.numel() * 8 / ms / 1e6:.1f} GB/s (read+write)")

#!/usr/bin/env python
"""Per-step device time of a scan against the text size (ASCII, m = 16, ~16 hits per MiB): the intercept is the
fixed cost of a call (launch ramp, pipeline prologue, last-CTA tail, expand kernel), the slope the streaming rate.
    python profiles/size_sweep.py"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402

dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
alpha = bmx.synth.ALPHABETS["ascii95"]
big = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
bmx.synth.fill_device(big, 0, 42, alpha)
pat = bmx.synth.pattern_from_stream(16, 42, alpha)
bmx.synth.plant_device(big, pat, bmx.synth.plant_offsets(big.numel(), 16, 8192, 42))
pos = torch.empty(1 << 20, dtype=torch.int64, device=dev)
sc = bmx.Scanner(0)
sc.set_pattern(pat, stream=stream)
sc.set_timing(0)
print(f"{'bytes':>12s} {'positions us':>13s} {'GB/s':>8s} {'count-only us':>14s} {'GB/s':>8s} {'hits':>7s}")
for shift in range(16, 30):
    n = 1 << shift
    text = big[:n]
    row = []
    for mode in ("positions", "count"):
        def run():
            sc.begin(pos if mode == "positions" else None, stream=stream)
            sc.scan(text, 0, stream=stream)
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        reps = 200 if n <= (64 << 20) else 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        cnt, _ = sc.finish(stream=stream)
        us = e0.elapsed_time(e1) / reps * 1e3
        row += [us, n / us / 1e3]
    print(f"{n:12d} {row[0]:13.2f} {row[1]:8.1f} {row[2]:14.2f} {row[3]:8.1f} {cnt:7d}", flush=True)

#!/usr/bin/env python
"""End-to-end throughput of bmx_search (host pointers) for pageable vs pinned host text.
python profiles/e2e_host_memory.py [GiB]"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30))
dev = torch.device("cuda:0")
t = torch.empty(n, dtype=torch.uint8, device=dev)
bmx.synth.fill_device(t, 0, 43, bmx.synth.ALPHABETS["dna"])
pat = bmx.synth.fill_host(12345, 32, 43, bmx.synth.ALPHABETS["dna"]).tobytes()
pageable = t.cpu().numpy()
pinned = torch.from_numpy(pageable).pin_memory()
for name, buf in (("pinned", pinned), ("pageable", pageable)):
    for threads in ((None,) if name == "pinned" else (1, 2, 4, 8, 12, 16, 24, 32)):
        if threads:
            os.environ["BMX_STAGING_THREADS"] = str(threads)
        bmx.search(buf, pat, max_positions=1 << 16)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            c, p = bmx.search(buf, pat, max_positions=1 << 16)
        dt = (time.perf_counter() - t0) / reps
        print(f"{name:9s} staging_threads={threads} {n / dt / 1e9:7.1f} GB/s  ({dt * 1e3:.1f} ms, {c} hits)", flush=True)

#!/usr/bin/env python
"""K patterns in one pass (bmx_search_multi_device) against K single-pattern passes over the same 4 GiB of
device-resident text: wall time per call (both are synchronous calls), counts checked against each other.
    python profiles/multi_pattern_bench.py"""
from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402

dev = torch.device("cuda:0")
n = 4 << 30


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r


for alphabet, m in (("dna", 32), ("ascii95", 16), ("bytes256", 8), ("dna", 12)):
    alpha = bmx.synth.ALPHABETS[alphabet]
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    bmx.synth.fill_device(text, 0, 43, alpha)
    for K in (1, 4, 16, 64):
        pats = [bmx.synth.fill_host(1_000_003 * (k + 1), m, 43, alpha).tobytes() for k in range(K)]   # cut from the text: >= 1 hit each
        ms_multi, res = timed(lambda: bmx.search_multi_device(text, pats, max_positions=4096))
        ms_serial, ser = timed(lambda: [bmx.search_device(text, p, max_positions=4096)[0] for p in pats], reps=2)
        ok = [c for c, _ in res] == list(ser)
        one = ms_serial / K
        print(f"{alphabet:9s} m={m:<3d} K={K:<3d} one pass {ms_multi:8.3f} ms ({n / ms_multi / 1e6:7.1f} GB/s of text, "
              f"{K * n / ms_multi / 1e6:8.1f} GB/s pattern-text)   K passes {ms_serial:8.3f} ms   single {one:6.3f} ms   "
              f"ratio one-pass/single {ms_multi / one:5.2f}   hits {sum(c for c, _ in res)}  counts equal: {ok}", flush=True)
    del text

#!/usr/bin/env python
"""Early-exit first-occurrence query (bmx_find_first_device): wall time per call against where the first
match lies in a 4 GiB device-resident DNA text, next to a full positions scan of the same text.
    python profiles/find_first_bench.py"""
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
text = torch.empty(n, dtype=torch.uint8, device=dev)
bmx.synth.fill_device(text, 0, 77, bmx.synth.ALPHABETS["dna"])
pat = b"ACGTTGCAACGTTGCAGGCCTTAAGGCCTTAAACGT"   # 36 bytes: no natural occurrence in 4 GiB of random DNA
assert bmx.find_first_device(text, pat) == -1


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6, r


us, _ = timed(lambda: bmx.search_device(text, pat, max_positions=16)[0])
print(f"full scan (bmx_search_device, 4 GiB)         {us:9.1f} us per call")
us, r = timed(lambda: bmx.find_first_device(text, pat))
print(f"find_first, no match (whole text scanned)    {us:9.1f} us per call  -> {r}")
for at in (1000, 10 << 20, 100 << 20, 1 << 30, (4 << 30) - 100):
    bmx.synth.plant_device(text, pat, [at])
    us, r = timed(lambda: bmx.find_first_device(text, pat))
    assert r == at, (r, at)
    print(f"find_first, first match at byte {at:<12d} {us:9.1f} us per call")
    text[at:at + len(pat)] = ord("A")   # remove it again (leaves a run of A's, which cannot match the pattern)

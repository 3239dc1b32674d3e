#!/usr/bin/env python
"""The reference's own kind of input at scale: its English fixture (input5L.txt, from
tests/golden/golden.npz) tiled to 1 GiB, searched for the reference's pattern "is" and others.
Every 2 KiB segment has hits here, so this exercises the mid-density path (masks + staged expand).
python profiles/english_bench.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402
from conftest import Golden, load_oracle  # noqa: E402

g = Golden(ROOT / "tests" / "golden" / "golden.npz")
base = np.frombuffer(g.text("input5L"), dtype=np.uint8)
n = 1 << 30
reps = n // base.size + 1
dev = torch.device("cuda:0")
text = torch.from_numpy(np.tile(base, reps)[:n].copy()).to(dev)
stream = torch.cuda.current_stream().cuda_stream
sc = bmx.Scanner(0)
pos = torch.empty(n // 8, dtype=torch.int64, device=dev)
oracle = load_oracle()
for pat in (b"is", b"the", b" ", b"position", b"HACKHACK", b"occurrences starting from", b"e"):
    sc.set_pattern(pat, stream=stream)
    res = []
    for mode in ("positions", "count"):
        def run():
            sc.begin(pos if mode == "positions" else None, stream=stream)
            sc.scan(text, 0, stream=stream)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        cnt, st = sc.finish(stream=stream)
        ms = e0.elapsed_time(e1) / 10
        res.append(f"{mode}: {n / ms / 1e6:7.1f} GB/s ({ms * 1e3:7.1f} us, scan kernel {st['scan_kernel_ms'] * 1e3:6.1f} us)")
    # parity on the first 64 MiB against the oracle
    w = 64 << 20
    sc.begin(pos, stream=stream)
    sc.scan(text[:w], 0, stream=stream)
    c64, _ = sc.finish(stream=stream)
    want = oracle.search_np(text[:w].cpu().numpy(), pat, threads=-1)
    ok = c64 == want.size and np.array_equal(pos[:c64].cpu().numpy(), want)
    print(f"{pat!r:32} {st['variant']:8s} hits={cnt:<10d} parity64MiB={ok}  " + "  |  ".join(res), flush=True)

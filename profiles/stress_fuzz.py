#!/usr/bin/env python
"""One-off stress fuzz on the GPU: random texts of every density (sigma 1..256, n up to a few MiB,
periodic and mixed texts), every variant, odd alignments, truncated capacities -- each case against
the oracle.  python profiles/stress_fuzz.py [cases] [seed]"""
import random
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402
from conftest import load_oracle  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 600
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rnd = random.Random(seed)
rng = np.random.default_rng(seed)
oracle = load_oracle()
dev = torch.device("cuda:0")
bad = 0
for it in range(cases):
    sigma = rnd.choice([1, 2, 2, 3, 4, 4, 8, 26, 95, 256])
    n = rnd.choice([rnd.randint(1, 4000), rnd.randint(4000, 300000), rnd.randint(300000, 5 << 20)])
    kind = rnd.choice(["random", "random", "periodic", "mixed"])
    if kind == "periodic":
        unit = bytes(rnd.randrange(sigma) + (60 if sigma < 190 else 0) for _ in range(rnd.randint(1, 7)))
        text = np.frombuffer((unit * (n // len(unit) + 1))[:n], dtype=np.uint8).copy()
    else:
        text = (rng.integers(0, sigma, size=n, dtype=np.uint8) + (60 if sigma < 190 else 0)).astype(np.uint8)
        if kind == "mixed" and n > 1000:
            a, b = sorted(rnd.sample(range(n), 2))
            text[a:b] = text[a]
    m = rnd.choice([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16, 25, 32, 33, 64, 100, 128, 300, 1500])
    m = min(m, max(1, n))
    if rnd.random() < 0.7 and n >= m:
        o = rnd.randint(0, n - m)
        pat = text[o:o + m].tobytes()
    else:
        pat = bytes(rnd.randrange(sigma) + (60 if sigma < 190 else 0) for _ in range(m))
    want = oracle.search_np(text, pat, threads=4) if n > 100000 else oracle.search(text.tobytes(), pat)
    mis = rnd.randint(0, 20)
    buf = torch.empty(n + mis + 64, dtype=torch.uint8, device=dev)
    td = buf[mis:mis + n]
    td.copy_(torch.from_numpy(text))
    variants = ["auto", "window"] + (["qgram"] if m >= 7 else []) + (["shiftand"] if m <= 32 else [])
    for v in variants:
        cap = rnd.choice([max(n, 1), max(want.size // 2, 1), 1])
        c, pos, _ = bmx.search_device(td, pat, max_positions=cap, variant=v)
        c2, _, _ = bmx.search_device(td, pat, variant=v)
        got = pos.cpu().numpy()
        if c != want.size or c2 != want.size or not np.array_equal(got, want[:cap]):
            bad += 1
            print("MISMATCH", it, kind, sigma, n, m, v, cap, c, c2, want.size, flush=True)
    if it % 100 == 0:
        print(f"case {it}: ok so far (bad={bad})", flush=True)
print("done: cases", cases, "mismatches", bad)
sys.exit(1 if bad else 0)

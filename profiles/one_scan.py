#!/usr/bin/env python
"""A handful of scans of one (text, pattern) pair -- the thing to put under ncu.
    python profiles/one_scan.py english "occurrences starting from" [reps] [multi-K]
    python profiles/one_scan.py dna32 - 5 16        # K = 16 patterns cut from the text, one pass
text: english (the reference's fixture input5L.txt tiled to 1 GiB) | dna32 | ascii16 (1 GiB synthetic)"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402

kind, pat = sys.argv[1], sys.argv[2].encode()
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
K = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda:0")
n = 1 << 30
if kind == "english":
    from conftest import Golden
    base = np.frombuffer(Golden(ROOT / "tests" / "golden" / "golden.npz").text("input5L"), dtype=np.uint8)
    text = torch.from_numpy(np.tile(base, n // base.size + 1)[:n].copy()).to(dev)
else:
    alpha = bmx.synth.ALPHABETS["dna" if kind.startswith("dna") else "ascii95"]
    m = int(kind.lstrip("dnasci"))
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    bmx.synth.fill_device(text, 0, 43, alpha)
    if pat == b"-":
        pat = bmx.synth.fill_host(1_000_003, m, 43, alpha).tobytes()
stream = torch.cuda.current_stream().cuda_stream
if K:
    m = len(pat)
    pats = [text[1_000_003 * (k + 1): 1_000_003 * (k + 1) + m].cpu().numpy().tobytes() for k in range(K)]
    for _ in range(reps):
        res = bmx.search_multi_device(text, pats, max_positions=4096)
    print("multi", K, [c for c, _ in res][:8])
else:
    sc = bmx.Scanner(0)
    sc.set_pattern(pat, stream=stream)
    pos = torch.empty(n // 8, dtype=torch.int64, device=dev)
    for _ in range(reps):
        sc.begin(pos, stream=stream)
        sc.scan(text, 0, stream=stream)
    cnt, st = sc.finish(stream=stream)
    print(pat, cnt, st)

#!/usr/bin/env python
"""Verification knobs side by side in ONE process on one box: the scan of mid-density texts (the reference's English
fixture tiled to 1 GiB; 1 GiB of DNA with short patterns) under every combination of
    BMX_COOP_VERIFY  1: flagged chunks checked by the whole warp, 0: by their own lane with BM skips
    BMX_DENSE_LANES  candidate lanes per segment from which a warp builds all masks right away
(both read by plan_scan at every scan).  Count-only and positions, GB/s of text.
    python profiles/verify_ab.py [coop values] [dense values]      e.g.  1,0  0,3,6,12,33   (0 = the library's default)"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402
from conftest import Golden  # noqa: E402

coops = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "0"]
denses = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "12", "33"]
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
n = 1 << 30
base = np.frombuffer(Golden(ROOT / "tests" / "golden" / "golden.npz").text("input5L"), dtype=np.uint8)
english = torch.from_numpy(np.tile(base, n // base.size + 1)[:n].copy()).to(dev)
dna = torch.empty(n, dtype=torch.uint8, device=dev)
alpha = bmx.synth.ALPHABETS["dna"]
bmx.synth.fill_device(dna, 0, 4321, alpha)
cases = [("english", english, p) for p in (b"position", b"HACKHACK", b"occurrences starting from", b"reference", b"HACKH")]
cases += [("dna", dna, bmx.synth.pattern_from_stream(m, 2000 + m, alpha)) for m in (5, 6, 7, 8, 9, 10, 12)]
sc = bmx.Scanner(0)
pos = torch.empty(n // 8, dtype=torch.int64, device=dev)


def rate(text, positions, reps=10):
    def run():
        sc.begin(pos if positions else None, stream=stream)
        sc.scan(text, 0, stream=stream)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    cnt, _ = sc.finish(stream=stream)
    return n / (e0.elapsed_time(e1) / reps) / 1e6, cnt


print(f"{'text':8s} {'pattern':28s} {'hits':>10s}  " + "  ".join(f"coop{c}/dense{d:>2s} pos | count" for c in coops for d in denses))
for name, text, pat in cases:
    cells, counts = [], set()
    for c in coops:
        for d in denses:
            os.environ["BMX_COOP_VERIFY"] = c
            if d == "0":
                os.environ.pop("BMX_DENSE_LANES", None)
            else:
                os.environ["BMX_DENSE_LANES"] = d
            sc.set_pattern(pat, stream=stream)
            rp, cnt = rate(text, True)
            rc, cnt2 = rate(text, False)
            counts.update((cnt, cnt2))
            cells.append(f"{rp:12.0f} | {rc:5.0f}")
    print(f"{name:8s} {pat!r:28s} {min(counts):>10d}{'' if len(counts) == 1 else ' MISMATCH'}  " + "  ".join(cells), flush=True)

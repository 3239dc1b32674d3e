#!/usr/bin/env python
"""QGRAM on 7 <= m <= 10: per-residue q-gram lengths (default) against one length for all residues
(BMX_QGRAM_UNIFORM=1) and the library's own choice (auto: per-residue when the pattern has <= 4 distinct
bytes and m <= 9).  1 GiB device-resident text, positions written.
    python profiles/short_qgram_bench.py"""
from __future__ import annotations

import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402

dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
n = 1 << 30
sc = bmx.Scanner(0)
pos = torch.empty(1 << 24, dtype=torch.int64, device=dev)
text = torch.empty(n, dtype=torch.uint8, device=dev)
for alpha_name in ("dna", "ascii95", "bytes256"):
    alpha = bmx.synth.ALPHABETS[alpha_name]
    bmx.synth.fill_device(text, 0, 4321, alpha)
    for m in (7, 8, 9, 10, 11):
        pat = bmx.synth.pattern_from_stream(m, 2000 + m, alpha)
        bmx.synth.plant_device(text, pat, bmx.synth.plant_offsets(n, m, 100, m))
        cells, counts = [], []
        for uniform in ("1", "0", "-1"):
            os.environ["BMX_QGRAM_UNIFORM"] = uniform
            sc.set_pattern(pat, variant="qgram", stream=stream)
            for _ in range(3):
                sc.begin(pos, stream=stream)
                sc.scan(text, 0, stream=stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                sc.begin(pos, stream=stream)
                sc.scan(text, 0, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            cnt, _ = sc.finish(stream=stream)
            counts.append(cnt)
            cells.append(n / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9)
        print(f"{alpha_name:9s} m={m:2d} hits={counts[0]:<9d} same={counts[0] == counts[1]}  "
              f"uniform {cells[0]:7.1f} GB/s   per-residue {cells[1]:7.1f} GB/s   auto {cells[2]:7.1f} GB/s", flush=True)

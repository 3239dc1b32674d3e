#!/usr/bin/env python
"""WINDOW filter for m = 1..4 on sparse text (bytes256, 4 GiB): the byte-parallel flag construction (m <= 3)
against the per-position compare (m = 4).    python profiles/short_window_bench.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402

dev = torch.device("cuda:0")
n = 4 << 30
stream = torch.cuda.current_stream().cuda_stream
sc = bmx.Scanner(0)
sc.set_timing(0)
pos = torch.empty(1 << 26, dtype=torch.int64, device=dev)
for alphabet in ("bytes256", "ascii95"):
    alpha = bmx.synth.ALPHABETS[alphabet]
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    bmx.synth.fill_device(text, 0, 44, alpha)
    for m in (1, 2, 3, 4):
        pat = bmx.synth.pattern_from_stream(m, 44, alpha)
        sc.set_pattern(pat, stream=stream)
        out = []
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
            out.append(f"{mode} {n / ms / 1e6:7.1f} GB/s ({ms * 1e3:7.1f} us)")
        print(f"{alphabet:9s} m={m} hits={cnt:<10d} " + "  |  ".join(out), flush=True)
    del text

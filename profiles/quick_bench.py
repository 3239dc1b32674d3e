#!/usr/bin/env python
"""Quick device-resident throughput table over the bench workloads (positions + count-only).
python profiles/quick_bench.py [workload ...]"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402

names = sys.argv[1:] or ["dna_m32_4GiB", "bytes256_m4_4GiB", "bytes256_m16_4GiB", "bytes256_m128_4GiB",
                         "ascii95_m64_shard", "ascii95_m16_64MiB", "aaa_1GiB"]
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
sc = bmx.Scanner(0)
for name in names:
    w = bench.WORKLOADS[name]
    n, m = w["n"], w["m"]
    alpha = bmx.synth.ALPHABETS[w["alphabet"]]
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    bmx.synth.fill_device(text, 0, w["seed"], alpha)
    pat = bench.make_pattern(bmx, w, n)
    bmx.synth.plant_device(text, pat, bench.plant_list(bmx, w, n, 1))
    cap = n if w["alphabet"] == "a" else 1 << 20
    pos = torch.empty(cap, dtype=torch.int64, device=dev)
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
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        cnt, st = sc.finish(stream=stream)
        ms = e0.elapsed_time(e1) / 20
        gbs = n / (ms * 1e-3) / 1e9
        traffic = (n + (8 * cnt if mode == "positions" else 0)) / (ms * 1e-3) / 1e9
        out.append(f"{mode}: {gbs:7.1f} GB/s scanned ({traffic:7.1f} GB/s algorithmic, {ms * 1e3:8.1f} us, scan kernel {st['scan_kernel_ms'] * 1e3:7.1f} us)")
    ok = bench.verify_hits(torch, text, 0, pat, pos, cnt, cap)
    print(f"{name:22s} {st['variant']:8s} hits={cnt:<11d} ok={ok}  " + "  |  ".join(out), flush=True)
    del text, pos

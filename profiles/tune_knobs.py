#!/usr/bin/env python
"""Sweeps the launch knobs of the scan kernel (tile size, pipeline depth, CTAs per SM) on one
B200 for a given workload; prints GB/s per setting.  python profiles/tune_knobs.py [workload]"""
from __future__ import annotations

import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "dna_m32_4GiB"
w = bench.WORKLOADS[name]
dev = torch.device("cuda:0")
n, m = w["n"], w["m"]
alpha = bmx.synth.ALPHABETS[w["alphabet"]]
text = torch.empty(n, dtype=torch.uint8, device=dev)
bmx.synth.fill_device(text, 0, w["seed"], alpha)
pat = bench.make_pattern(bmx, w, n)
bmx.synth.plant_device(text, pat, bench.plant_list(bmx, w, n, 1))
cap = n if w["alphabet"] == "a" else 1 << 20
pos = torch.empty(cap, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
sc = bmx.Scanner(0)
sc.set_pattern(pat, stream=stream)
for mode in ("positions", "count"):
    for tile in (16384, 32768):
        for ctas in (2, 1):
            for stages in (2, 3, 4, 6, 8):
                os.environ.update(BMX_TILE=str(tile), BMX_CTAS_PER_SM=str(ctas), BMX_STAGES=str(stages))
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
                print(f"{name} {mode:9s} tile={tile} ctas/sm={ctas} stages={st['stages']} (asked {stages}) "
                      f"{n / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e9:8.1f} GB/s hits={cnt}", flush=True)

#!/usr/bin/env python
"""Single-process multi-GPU entry point (bmx_mg_search) on all visible GPUs: parity against the oracle on a
64 MiB text with plants across every shard seam, then end-to-end throughput from pinned and pageable host text.
python profiles/mg_check.py [GiB per GPU]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402
from conftest import load_oracle  # noqa: E402

oracle = load_oracle()
mg = bmx.MultiGpu(0)
R = mg.ngpus
print("GPUs:", R)
alpha = bmx.synth.ALPHABETS["ascii95"]
m = 64
pat = bmx.synth.pattern_from_stream(m, 47, alpha)
n = (64 << 20) + 12345
text = bmx.synth.fill_host(0, n, 47, alpha)
per = -(-(-(-n // R)) // 16) * 16
plants = [r * per - d for r in range(1, R) for d in (m // 2, 1, m - 1, m, 0)] + list(bmx.synth.plant_offsets(n, m, 500, 47))
bmx.synth.plant_host(text, pat, plants)
want = oracle.search_np(text, pat, threads=-1)
count, pos, shard = mg.search(text, pat)
print("parity:", count == want.size and np.array_equal(pos, want), "hits", count, "per GPU", shard)
assert count == want.size and np.array_equal(pos, want)

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30)) * R
big = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dev_chunk = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0")
for off in range(0, n, 1 << 30):
    k = min(1 << 30, n - off)
    bmx.synth.fill_device(dev_chunk[:k], off, 47, alpha)
    big[off:off + k].copy_(dev_chunk[:k])
torch.cuda.synchronize()
for name, buf in (("pinned", big), ("pageable", big.numpy().copy())):
    mg.search(buf, pat, max_positions=1 << 16)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        c, p, sh = mg.search(buf, pat, max_positions=1 << 16)
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:9s} {n / (1 << 30):.0f} GiB over {R} GPUs: {n / dt / 1e9:7.1f} GB/s end to end ({dt * 1e3:.1f} ms, {c} hits)", flush=True)

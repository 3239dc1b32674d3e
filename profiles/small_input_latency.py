#!/usr/bin/env python
"""Latency of one search on inputs of the reference's own size (its fixture input5L.txt, 500 KB, pattern "is"):
host-pointer call (H2D + scan + D2H) and device-resident call, against the reference's serial code on the host."""
import ctypes
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx  # noqa: E402
from conftest import Golden  # noqa: E402

g = Golden(ROOT / "tests" / "golden" / "golden.npz")
text = g.text("input5L")
pat = b"is"
dev = torch.device("cuda:0")
td = torch.from_numpy(np.frombuffer(text, dtype=np.uint8).copy()).to(dev)
pos = torch.empty(1 << 16, dtype=torch.int64, device=dev)


def timed(f, reps=200):
    for _ in range(20):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


c, p = bmx.search(text, pat)
print(f"input5L.txt ({len(text)} bytes), pattern 'is': {c} hits")
print(f"bmx_search (host pointers, pageable)      {timed(lambda: bmx.search(text, pat)):8.1f} us per call")
print(f"bmx_search_device (text resident)         {timed(lambda: bmx.search_device(td, pat, pos_out=pos)):8.1f} us per call")
print(f"bmx_search_device, count only             {timed(lambda: bmx.search_device(td, pat)):8.1f} us per call")
so = ROOT / "oracle" / "_ref" / "libref_bm.so"
if so.exists():
    ref = ctypes.CDLL(str(so))
    out = np.zeros(1 << 16, dtype=np.int64)
    cnt = ctypes.c_uint64()
    f = lambda: ref.ref_bm_search(ctypes.c_char_p(text), ctypes.c_int64(len(text)), ctypes.c_char_p(pat), ctypes.c_int32(len(pat)),  # noqa: E731
                                  out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(out.size), ctypes.byref(cnt))
    t0 = time.perf_counter()
    for _ in range(50):
        f()
    print(f"reference serial code on the host (1 core) {(time.perf_counter() - t0) / 50 * 1e6:8.1f} us per call")

import sys
from pathlib import Path
import numpy as np
ROOT = Path("/root/repo")
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx
from conftest import Golden
g = Golden(ROOT / "tests" / "golden" / "golden.npz")
base = np.frombuffer(g.text("input5L"), dtype=np.uint8)
n = 1 << 30
text = torch.from_numpy(np.tile(base, n // base.size + 1)[:n].copy()).cuda()
stream = torch.cuda.current_stream().cuda_stream
sc = bmx.Scanner(0); sc.set_timing(2)
pos = torch.empty(n // 8, dtype=torch.int64, device="cuda")
for pat in (b"is", b"position"):
    sc.set_pattern(pat, stream=stream)
    for mode in ("positions", "count"):
        ks = []
        for _ in range(8):
            sc.begin(pos if mode == "positions" else None, stream=stream)
            sc.scan(text, 0, stream=stream)
            c, st = sc.finish(stream=stream)
            ks.append(st["scan_kernel_ms"] * 1e3)
        print(pat, mode, "scan kernel us", round(float(np.median(ks)), 1), flush=True)

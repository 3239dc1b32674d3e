#!/usr/bin/env python
"""Builds and runs profiles/micro/call_latency.cpp on the reference's own fixture (input5L.txt from
tests/golden/golden.npz, pattern "is") and on a sparse pattern: wall time per C-ABI call without Python."""
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import Golden  # noqa: E402

pkg = ROOT / "parallel_implementation_of_string_matching_algorithms_opencl_b200"
exe = ROOT / "profiles" / "micro" / "call_latency"
subprocess.run(["nvcc", "-O2", "-std=c++17", "-o", str(exe), str(ROOT / "profiles" / "micro" / "call_latency.cpp"),
                f"-I{ROOT / 'include'}", f"-L{pkg}", "-lbmx", f"-Xlinker=-rpath,{pkg}"], check=True)
g = Golden(ROOT / "tests" / "golden" / "golden.npz")
with tempfile.TemporaryDirectory() as d:
    f = Path(d) / "input5L.txt"
    f.write_bytes(g.text("input5L"))
    for pat in ("is", "occurrences starting from", "HACKHACK"):
        subprocess.run([str(exe), str(f), pat], check=True)
        print()

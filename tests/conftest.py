"""Shared fixtures.  `-m "not gpu"` runs on a CPU-only box (oracle, host logic, ABI surface);
`-m gpu` runs the parity tests proper through the C ABI on a B200."""
from __future__ import annotations

import ctypes
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Oracle:
    """ctypes face of oracle/liboracle.so (the CPU restatement) -- the CHECKER, tests only."""

    def __init__(self, lib):
        self.lib = lib

    def search(self, text: bytes, pat: bytes):
        text = bytes(text)
        n, m = len(text), len(pat)
        cap = max(n - m + 1, 1)
        pos = np.zeros(cap, dtype=np.int64)
        cnt = ctypes.c_uint64()
        rc = self.lib.oracle_search(ctypes.c_char_p(text), ctypes.c_int64(n), ctypes.c_char_p(pat), ctypes.c_int32(m),
                                    pos.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(cap), ctypes.byref(cnt))
        assert rc == 0, rc
        return pos[: cnt.value].copy()

    def search_np(self, text: np.ndarray, pat: bytes, threads: int = -1):
        """Multi-threaded windowed oracle on a numpy uint8 buffer (no copy)."""
        n, m = text.size, len(pat)
        cnt = ctypes.c_uint64()
        rc = self.lib.oracle_search_mt(ctypes.c_void_p(text.ctypes.data), ctypes.c_int64(n), ctypes.c_char_p(pat),
                                       ctypes.c_int32(m), None, ctypes.c_int64(0), ctypes.byref(cnt), ctypes.c_int32(threads))
        assert rc == 0, rc
        pos = np.zeros(max(cnt.value, 1), dtype=np.int64)
        rc = self.lib.oracle_search_mt(ctypes.c_void_p(text.ctypes.data), ctypes.c_int64(n), ctypes.c_char_p(pat),
                                       ctypes.c_int32(m), pos.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(pos.size),
                                       ctypes.byref(cnt), ctypes.c_int32(threads))
        assert rc == 0, rc
        return pos[: cnt.value]

    def tables(self, pat: bytes):
        bad = np.zeros(256, dtype=np.int32)
        good = np.zeros(max(len(pat), 1), dtype=np.int32)
        self.lib.oracle_build_bad(ctypes.c_char_p(pat), len(pat), bad.ctypes.data_as(ctypes.c_void_p))
        self.lib.oracle_build_good(ctypes.c_char_p(pat), len(pat), good.ctypes.data_as(ctypes.c_void_p))
        return bad, good[: len(pat)]

    def partition_words(self, text: bytes, nparts: int):
        se = np.zeros(2 * nparts, dtype=np.int32)
        self.lib.oracle_partition_words(ctypes.c_char_p(bytes(text) + b"\0"), nparts, se.ctypes.data_as(ctypes.c_void_p))
        return se

    def search_partitions(self, text: bytes, pat: bytes, se):
        se = np.ascontiguousarray(se, dtype=np.int32)
        ans = np.zeros(se.size // 2, dtype=np.int32)
        rc = self.lib.oracle_search_partitions(ctypes.c_char_p(bytes(text)), ctypes.c_char_p(pat),
                                               se.ctypes.data_as(ctypes.c_void_p), ans.ctypes.data_as(ctypes.c_void_p),
                                               len(pat), ans.size)
        assert rc == 0
        return ans

    def synth_fill(self, offset: int, length: int, seed: int, alphabet: bytes):
        buf = np.zeros(length, dtype=np.uint8)
        self.lib.oracle_synth_fill(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(offset), ctypes.c_int64(length),
                                   ctypes.c_uint64(seed), ctypes.c_char_p(alphabet), len(alphabet))
        return buf


def load_oracle() -> Oracle:
    so = ROOT / "oracle" / "liboracle.so"
    src = ROOT / "oracle" / "bm_oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["bash", str(ROOT / "oracle" / "build_oracle.sh")], check=True, capture_output=True)
    lib = ctypes.CDLL(str(so))
    lib.oracle_fnv1a64_positions.restype = ctypes.c_uint64
    return Oracle(lib)


@pytest.fixture(scope="session")
def oracle() -> Oracle:
    return load_oracle()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own code compiled here (oracle/_ref/libref_bm.so); skip when absent."""
    so = ROOT / "oracle" / "_ref" / "libref_bm.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libref_bm.so not built (needs /root/reference)")
    return ctypes.CDLL(str(so))


class Golden:
    def __init__(self, path: Path):
        z = np.load(path)
        self.z = z
        self.manifest = json.loads(bytes(z["manifest"]).decode())

    def text(self, name: str) -> bytes:
        return bytes(self.z[f"text/{name}"])

    def cases(self):
        for i, c in enumerate(self.manifest["cases"]):
            yield c["text"], bytes.fromhex(c["pattern_hex"]), self.z[f"case/{i}/pos"], c["source"]

    def tables(self):
        for i, t in enumerate(self.manifest["tables"]):
            yield bytes.fromhex(t["pattern_hex"]), self.z[f"table/{i}/bad128"], self.z[f"table/{i}/good"]

    def parts(self):
        for i, p in enumerate(self.manifest["parts"]):
            yield p["text"], bytes.fromhex(p["pattern_hex"]), p["nparts"], self.z[f"part/{i}/se"], self.z[f"part/{i}/ans"]

    def synthetic(self):
        for i, s in enumerate(self.manifest["synthetic"]):
            yield s, bytes.fromhex(s["pattern_hex"]), self.z[f"syn/{i}/pos"]


@pytest.fixture(scope="session")
def golden() -> Golden:
    return Golden(ROOT / "tests" / "golden" / "golden.npz")


@pytest.fixture(scope="session")
def bmx():
    """The product package with libbmx.so built (nvcc cross-compiles without a GPU)."""
    import parallel_implementation_of_string_matching_algorithms_opencl_b200 as pkg
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import build as b

    if not pkg.LIB_PATH.exists():
        b.build()
    pkg._lib.load()
    return pkg


def synth_text(bmx, spec: dict) -> tuple[np.ndarray, bytes]:
    """Rebuild a golden synthetic case: (text, pattern)."""
    alpha = bmx.synth.ALPHABETS[spec["alphabet"]]
    text = bmx.synth.fill_host(0, spec["n"], spec["seed"], alpha)
    pat = bytes.fromhex(spec["pattern_hex"])
    bmx.synth.plant_host(text, pat, bmx.synth.plant_offsets(spec["n"], spec["m"], spec["plants"], spec["seed"]))
    return text, pat

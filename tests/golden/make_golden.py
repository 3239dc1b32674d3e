#!/usr/bin/env python
"""Generates tests/golden/golden.npz from the REFERENCE'S OWN CODE (oracle/_ref, built by
oracle/build_oracle.sh from /root/reference) run on the reference's fixtures and on seeded
synthetic inputs.  Run in the authoring container only (it reads /root/reference); the .npz
travels to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py

Contents (all arrays; `manifest` is a JSON string describing them):
  text/<name>             reference fixture bytes (small fixtures + input5L.txt, compressed)
  case/<i>/pos            int64 ascending match starts
  table/<i>/bad128, good  the reference's own tables (BoyreMoore.cpp:153-190)
  part/<i>/se, ans        word partition + per-process counts (whole unmodified program)
Every case records `source`: "ref" = verbatim reference code (libref_bm.so);
"oracle" = widened restatement (inputs outside the reference's defined domain: bytes >= 0x80).
"""
from __future__ import annotations

import ctypes
import json
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from parallel_implementation_of_string_matching_algorithms_opencl_b200 import synth  # noqa: E402

REF = Path(os.environ.get("BMX_REFERENCE_ROOT", "/root/reference"))
ref = ctypes.CDLL(str(ROOT / "oracle/_ref/libref_bm.so"))
orc = ctypes.CDLL(str(ROOT / "oracle/liboracle.so"))


def run(lib, fn, text: bytes, pat: bytes):
    n, m = len(text), len(pat)
    cap = max(n, 1)
    pos = np.zeros(cap, dtype=np.int64)
    cnt = ctypes.c_uint64()
    rc = getattr(lib, fn)(ctypes.c_char_p(text), ctypes.c_int64(n), ctypes.c_char_p(pat), ctypes.c_int32(m),
                          pos.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(cap), ctypes.byref(cnt))
    assert rc == 0, (fn, rc)
    return pos[: cnt.value].copy()


def legal(text: bytes, pat: bytes) -> bool:
    return bool(ref.ref_bm_legal(ctypes.c_char_p(text), ctypes.c_int64(len(text)), ctypes.c_char_p(pat),
                                 ctypes.c_int32(len(pat))))


def main():
    arrays, manifest = {}, {"texts": {}, "cases": [], "tables": [], "parts": [], "synthetic": []}
    dbg = REF / "BoyreMoore/x64/Debug"
    fixtures = {
        "input2": dbg / "input2.txt", "input3": dbg / "input3.txt", "input4": dbg / "input4.txt",
        "input5": dbg / "input5.txt", "input6": dbg / "input6.txt", "input7": dbg / "input7.txt",
        "input8": dbg / "input8.txt", "input9": dbg / "input9.txt",
        "input5L": REF / "BoyreMoore/BoyreMoore/input5L.txt",
    }
    texts = {}
    for name, path in fixtures.items():
        texts[name] = path.read_bytes()
        arrays[f"text/{name}"] = np.frombuffer(texts[name], dtype=np.uint8)
        manifest["texts"][name] = {"bytes": len(texts[name]), "from": str(path.relative_to(REF))}

    search_pat = (dbg / "input1Search.txt").read_bytes()  # the reference's own pattern: b"is"
    plan = [(name, search_pat) for name in fixtures]
    plan += [("input5L", p) for p in (b"HACKHACK", b"position", b"the", b"e", b" ", b"occurrences starting from",
                                      b"In the picture above we are finding occurrences starting from position 4, but there is a")]
    plan += [("input7", p) for p in (b"there", b"hiHi", b"H", b"stronghiHiHello")]
    plan += [("input8", p) for p in (b"the", b"\xe2\x80\x99", b"and")]
    for name, pat in plan:
        t = texts[name]
        if legal(t, pat):
            pos, source = run(ref, "ref_bm_search", t, pat), "ref"
            assert np.array_equal(pos, run(orc, "oracle_search", t, pat)), (name, pat)
        else:
            pos, source = run(orc, "oracle_search", t, pat), "oracle"
        brute = np.array([i for i in range(len(t) - len(pat) + 1) if t[i:i + len(pat)] == pat], dtype=np.int64) \
            if len(t) <= 8192 else None
        if brute is not None:
            assert np.array_equal(pos, brute), (name, pat)
        i = len(manifest["cases"])
        arrays[f"case/{i}/pos"] = pos
        manifest["cases"].append({"text": name, "pattern_hex": pat.hex(), "count": int(pos.size), "source": source})

    # the reference's own tables
    for pat in (b"is", b"HACKHACK", b"there", b"hiHi", b"abcbab", b"position", b"aaa", b"a", b"abracadabra",
                b"GCAGAGAG", b"aabaabaab", b"xyzzyxyzzy" * 9):
        bad = np.zeros(128, dtype=np.int32)
        good = np.zeros(len(pat) + 1, dtype=np.int32)
        assert ref.ref_bm_build_tables(ctypes.c_char_p(pat), len(pat), bad.ctypes.data_as(ctypes.c_void_p),
                                       good.ctypes.data_as(ctypes.c_void_p)) == 0
        i = len(manifest["tables"])
        arrays[f"table/{i}/bad128"] = bad
        arrays[f"table/{i}/good"] = good[: len(pat)]  # entry 0 is never written by the reference
        manifest["tables"].append({"pattern_hex": pat.hex()})

    # whole unmodified program (BoyreMoore_ref): per-process counts for the 2-way word partition
    exe = ROOT / "oracle/_ref/BoyreMoore_ref"
    for name in ("input5L", "input7", "input6"):
        with tempfile.TemporaryDirectory() as d:
            Path(d, "inputEd.txt").write_bytes(texts[name])
            Path(d, "input1Search.txt").write_bytes(search_pat)
            Path(d, "kernel1.cl").write_bytes((dbg / "kernel1.cl").read_bytes())
            out = subprocess.run([str(exe)], cwd=d, capture_output=True, check=True).stdout.decode("latin-1")
        counts = [int(x) for x in re.findall(r"occurrences by process \d+ is (\d+)", out)][:2]
        se = np.zeros(4, dtype=np.int32)
        orc.oracle_partition_words(ctypes.c_char_p(texts[name] + b"\0"), 2, se.ctypes.data_as(ctypes.c_void_p))
        ans = np.zeros(2, dtype=np.int32)
        assert ref.ref_bm_search_partitions(ctypes.c_char_p(texts[name]), ctypes.c_int64(len(texts[name])),
                                            ctypes.c_char_p(search_pat), se.ctypes.data_as(ctypes.c_void_p),
                                            ans.ctypes.data_as(ctypes.c_void_p), len(search_pat), 2) == 0
        assert list(ans) == counts, (name, list(ans), counts)   # pins the restated partitioner
        i = len(manifest["parts"])
        arrays[f"part/{i}/se"] = se
        arrays[f"part/{i}/ans"] = np.array(counts, dtype=np.int32)
        manifest["parts"].append({"text": name, "pattern_hex": search_pat.hex(), "nparts": 2})

    # seeded synthetic cases (text regenerated from the seed on both sides; only hits are stored)
    syn = [
        ("ascii95", 1 << 20, 16, 42, 100), ("ascii95", 1 << 20, 4, 47, 50), ("ascii95", 300_000, 99, 48, 20),
        ("dna", 1 << 20, 32, 43, 100), ("dna", 1 << 20, 8, 49, 10), ("ascii128", 1 << 20, 16, 50, 64),
        ("ascii128", 1 << 20, 4, 51, 64), ("ascii128", 1 << 20, 99, 52, 64), ("ascii95", 1 << 20, 64, 53, 64),
        ("dna", 70_001, 7, 54, 5), ("dna", 50_000, 3, 55, 0), ("a", 40_000, 3, 56, 0),
    ]
    for alpha_name, n, m, seed, plants in syn:
        alpha = synth.ALPHABETS[alpha_name]
        text = synth.fill_host(0, n, seed, alpha)
        pat = b"a" * m if alpha_name == "a" else synth.pattern_from_stream(m, seed, alpha)
        offs = synth.plant_offsets(n, m, plants, seed)
        synth.plant_host(text, pat, offs)
        tb = text.tobytes()
        assert legal(tb, pat)
        pos = run(ref, "ref_bm_search", tb, pat)
        assert np.array_equal(pos, run(orc, "oracle_search", tb, pat))
        i = len(manifest["synthetic"])
        arrays[f"syn/{i}/pos"] = pos
        manifest["synthetic"].append({"alphabet": alpha_name, "n": n, "m": m, "seed": seed, "plants": plants,
                                      "pattern_hex": pat.hex(), "count": int(pos.size), "source": "ref"})

    arrays["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    out = Path(__file__).with_name("golden.npz")
    np.savez_compressed(out, **arrays)
    print(f"wrote {out} ({out.stat().st_size} bytes): {len(manifest['cases'])} fixture cases, "
          f"{len(manifest['tables'])} tables, {len(manifest['parts'])} partitions, {len(manifest['synthetic'])} synthetic")


if __name__ == "__main__":
    main()

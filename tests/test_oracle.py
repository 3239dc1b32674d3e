"""CPU: the oracle (oracle/bm_oracle.c) against the golden vectors made by the reference's own
code, against that code itself when it is built here, and against brute force."""
from __future__ import annotations

import ctypes
import random

import numpy as np

from conftest import synth_text


def brute(text: bytes, pat: bytes):
    m = len(pat)
    return np.array([i for i in range(len(text) - m + 1) if text[i:i + m] == pat], dtype=np.int64)


def test_oracle_matches_golden_fixture_cases(oracle, golden):
    n = 0
    for name, pat, pos, _source in golden.cases():
        got = oracle.search(golden.text(name), pat)
        assert np.array_equal(got, pos), (name, pat)
        n += 1
    assert n >= 20


def test_survey_known_answers(oracle, golden):
    """SURVEY.md appendix C rows (count, first, last, sum) for the reference's fixtures."""
    t = golden.text("input5L")
    for pat, count, first, last, total in [(b"is", 3291, 85, 499746, 821909099),
                                           (b"HACKHACK", 471, 176, 498017, 117309212),
                                           (b"position", 1097, 63, 498108, 273594359)]:
        pos = oracle.search(t, pat)
        assert (pos.size, int(pos[0]), int(pos[-1]), int(pos.sum())) == (count, first, last, total)
    t7 = golden.text("input7")
    assert oracle.search(t7, b"there").size == 16 and oracle.search(t7, b"hiHi").size == 19
    for name in ("input2", "input3", "input4", "input5", "input6", "input7"):
        assert oracle.search(golden.text(name), b"is").size == 0


def test_oracle_matches_golden_synthetic(oracle, golden, bmx):
    for spec, pat, pos in golden.synthetic():
        text, p = synth_text(bmx, spec)
        assert p == pat
        assert np.array_equal(oracle.search(text.tobytes(), pat), pos), spec


def test_oracle_tables_match_reference_tables(oracle, golden):
    for pat, bad128, good in golden.tables():
        bad, g = oracle.tables(pat)
        assert np.array_equal(bad[:128], bad128), pat
        assert np.all(bad[128:] == len(pat))
        assert np.array_equal(g[1:], good[1:]), pat


def test_survey_table_examples(oracle):
    ex = {b"HACKHACK": [8, 8, 8, 4, 4, 4, 4], b"there": [2, 5, 5, 5], b"hiHi": [2, 4, 4],
          b"abcbab": [2, 4, 4, 4, 4], b"position": [8] * 7, b"is": [2]}
    for pat, want in ex.items():
        assert list(oracle.tables(pat)[1][1:]) == want, pat


def test_oracle_partitioner_and_partition_counts(oracle, golden):
    for name, pat, nparts, se, ans in golden.parts():
        t = golden.text(name)
        assert np.array_equal(oracle.partition_words(t, nparts), se), name
        assert np.array_equal(oracle.search_partitions(t, pat, se), ans), name
    t = golden.text("input5L")
    assert list(oracle.partition_words(t, 2)) == [0, 250037, 250039, 500006]
    assert list(oracle.search_partitions(t, b"is", [0, 250037, 250039, 500006])) == [1649, 1642]


def test_oracle_vs_reference_code_fuzz(oracle, reflib):
    rnd = random.Random(7)
    for _ in range(20000):
        sigma = rnd.randint(1, 4)
        n, m = rnd.randint(0, 160), rnd.randint(1, 12)
        text = bytes(rnd.randrange(97, 97 + sigma) for _ in range(n))
        if n >= m and rnd.random() < 0.5:
            o = rnd.randint(0, n - m)
            pat = text[o:o + m]
        else:
            pat = bytes(rnd.randrange(97, 97 + sigma) for _ in range(m))
        cap = max(n, 1)
        pos = np.zeros(cap, dtype=np.int64)
        cnt = ctypes.c_uint64()
        rc = reflib.ref_bm_search(ctypes.c_char_p(text), ctypes.c_int64(n), ctypes.c_char_p(pat), ctypes.c_int32(m),
                                  pos.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(cap), ctypes.byref(cnt))
        assert rc == 0
        want = pos[: cnt.value]
        assert np.array_equal(oracle.search(text, pat), want)
        assert np.array_equal(want, brute(text, pat))
        # tables too
        bad128 = np.zeros(128, dtype=np.int32)
        good = np.zeros(m + 1, dtype=np.int32)
        assert reflib.ref_bm_build_tables(ctypes.c_char_p(pat), m, bad128.ctypes.data_as(ctypes.c_void_p),
                                          good.ctypes.data_as(ctypes.c_void_p)) == 0
        obad, ogood = oracle.tables(pat)
        assert np.array_equal(obad[:128], bad128) and np.array_equal(ogood[1:], good[1:m])


def test_oracle_edge_cases(oracle):
    assert oracle.search(b"", b"a").size == 0
    assert oracle.search(b"ab", b"abc").size == 0                       # m > n: loop never entered
    assert list(oracle.search(b"abc", b"abc")) == [0]                   # m == n
    assert list(oracle.search(b"aaaa", b"aaa")) == [0, 1]               # overlapping occurrences
    assert list(oracle.search(b"aaaa", b"a")) == [0, 1, 2, 3]           # m == 1
    assert list(oracle.search(b"\x00a\x00a", b"\x00a")) == [0, 2]       # NUL is an ordinary byte
    hi = bytes([200, 201, 200, 201, 200])
    assert list(oracle.search(hi, bytes([200, 201, 200]))) == [0, 2]    # bytes >= 0x80 (widened)
    assert oracle.lib.oracle_search(b"abc", ctypes.c_int64(3), b"", 0, None, ctypes.c_int64(0),
                                    ctypes.byref(ctypes.c_uint64())) == -1  # empty pattern rejected


def test_oracle_mt_equals_serial(oracle, bmx):
    text = bmx.synth.fill_host(0, 40 << 20, 99, bmx.synth.ALPHABETS["dna"])
    pat = text[12345:12345 + 9].tobytes()
    serial = oracle.search(text.tobytes(), pat)
    mt = oracle.search_np(text, pat, threads=4)
    assert serial.size > 50 and np.array_equal(serial, mt)


def test_synth_twins_agree(oracle, bmx):
    for name, alpha in bmx.synth.ALPHABETS.items():
        for off, ln in [(0, 4096), (13, 1000), (7, 1), (8, 8), (1 << 33, 513)]:
            assert np.array_equal(oracle.synth_fill(off, ln, 42, alpha), bmx.synth.fill_host(off, ln, 42, alpha)), (name, off)

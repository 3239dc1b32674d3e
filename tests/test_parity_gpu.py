"""GPU parity tests: the CUDA path, called through the C ABI, against the golden vectors made by
the reference's own code and against the oracle on the same seeded inputs.  Bit-exact: identical
count and identical ascending position list (integer/byte work, no tolerance)."""
from __future__ import annotations

import ctypes
import os
import random

import numpy as np
import pytest

from conftest import synth_text

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

VARIANTS_FOR = lambda m: ["auto", "window"] + (["qgram"] if m >= 7 else []) + (["shiftand"] if m <= 32 else [])  # noqa: E731


@pytest.fixture(scope="module")
def dev(bmx):
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    assert bmx.device_count() >= 1
    return torch.device("cuda:0")


def to_dev(buf, dev, misalign: int = 0):
    """CUDA uint8 tensor holding buf, optionally starting `misalign` bytes into an allocation."""
    a = np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    t = torch.empty(a.size + misalign + 64, dtype=torch.uint8, device=dev)
    view = t[misalign: misalign + a.size]
    if a.size:
        view.copy_(torch.from_numpy(a.copy()))
    return view


def _find_all(text: bytes, pat: bytes) -> np.ndarray:
    """All (overlapping) occurrences with bytes.find -- for patterns too long for the oracle's O(m^3) tables."""
    out, i = [], text.find(pat)
    while i >= 0:
        out.append(i)
        i = text.find(pat, i + 1)
    return np.array(out, dtype=np.int64)


def gpu_positions(bmx, text_dev, pat, variant="auto", cap=None):
    n = text_dev.numel()
    cap = max(n - len(pat) + 1, 1) if cap is None else cap
    count, pos, stats = bmx.search_device(text_dev, pat, max_positions=cap, variant=variant)
    got = pos.cpu().numpy() if pos is not None else np.zeros(0, dtype=np.int64)
    return count, got, stats


def test_golden_fixtures_device_and_host_paths(bmx, golden, dev):
    """Every known-answer vector of the reference's fixtures, through both entry points."""
    for name, pat, want, _source in golden.cases():
        text = golden.text(name)
        td = to_dev(text, dev)
        for variant in VARIANTS_FOR(len(pat)):
            count, got, _ = gpu_positions(bmx, td, pat, variant)
            assert count == want.size and np.array_equal(got, want), (name, pat, variant)
        count, got = bmx.search(text, pat)                       # host pointers: H2D + scan + D2H
        assert count == want.size and np.array_equal(got, want), (name, pat, "host")


def test_golden_synthetic_cases(bmx, golden, dev):
    for spec, pat, want in golden.synthetic():
        text, _ = synth_text(bmx, spec)
        td = to_dev(text, dev)
        for variant in VARIANTS_FOR(len(pat)):
            count, got, _ = gpu_positions(bmx, td, pat, variant)
            assert count == want.size and np.array_equal(got, want), (spec, variant)


def test_device_generator_matches_host_twin(bmx, dev):
    for name, alpha in bmx.synth.ALPHABETS.items():
        for off, ln, mis in [(0, 4096, 0), (13, 1000, 3), (7, 1, 1), ((1 << 33) + 5, 70001, 9)]:
            t = torch.zeros(ln + mis, dtype=torch.uint8, device=dev)[mis:]
            bmx.synth.fill_device(t, off, 42, alpha)
            assert np.array_equal(t.cpu().numpy(), bmx.synth.fill_host(off, ln, 42, alpha)), (name, off)


def test_fuzz_small_alphabets_all_variants(bmx, oracle, dev):
    """Tiny alphabets make every filter pass constantly and exercise overlap, borders and the
    Boyer-Moore pruning of candidate bits."""
    rnd = random.Random(1234)
    lengths = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 15, 16, 17, 31, 32, 33, 64, 99, 128]
    for it in range(260):
        sigma = rnd.choice([1, 2, 2, 3, 4])
        n = rnd.choice([rnd.randint(1, 300), rnd.randint(300, 70000)])
        m = rnd.choice(lengths)
        text = bytes(rnd.randrange(97, 97 + sigma) for _ in range(n))
        if n >= m and rnd.random() < 0.6:
            o = rnd.randint(0, n - m)
            pat = text[o:o + m]
        else:
            pat = bytes(rnd.randrange(97, 97 + sigma) for _ in range(m))
        want = oracle.search(text, pat)
        td = to_dev(text, dev, misalign=rnd.randint(0, 17))
        for variant in VARIANTS_FOR(m):
            count, got, _ = gpu_positions(bmx, td, pat, variant)
            assert count == want.size and np.array_equal(got, want), (it, sigma, n, m, variant)


def test_short_qgram_both_hash_layouts(bmx, oracle, dev, monkeypatch):
    """7 <= m <= 10: QGRAM hashes either one q-gram length for all four residues or min(8, m - r) bytes
    per residue (chosen from the pattern's alphabet).  Both layouts, forced through the measurement knob,
    must give the serial result on small and large alphabets."""
    rnd = random.Random(77)
    for it in range(120):
        sigma = rnd.choice([1, 2, 4, 6, 26, 200])
        n = rnd.randint(8, 50000)
        m = rnd.choice([7, 8, 9, 10, 11])
        text = bytes(rnd.randrange(40, 40 + sigma) for _ in range(n))
        o = rnd.randint(0, max(n - m, 0))
        pat = text[o:o + m] if n >= m and rnd.random() < 0.7 else bytes(rnd.randrange(40, 40 + sigma) for _ in range(m))
        want = oracle.search(text, pat)
        td = to_dev(text, dev, misalign=rnd.randint(0, 17))
        for knob in ("0", "1", None):
            if knob is None:
                monkeypatch.delenv("BMX_QGRAM_UNIFORM", raising=False)
            else:
                monkeypatch.setenv("BMX_QGRAM_UNIFORM", knob)
            count, got, stats = gpu_positions(bmx, td, pat, "qgram")
            assert stats["variant"] == "qgram"
            assert count == want.size and np.array_equal(got, want), (it, sigma, n, m, knob)


def test_edge_cases(bmx, oracle, dev):
    E = bmx._lib
    # m > n: not an error, count 0 (kernel1.cl:15,19); empty text
    assert bmx.search(b"ab", b"abc")[0] == 0
    assert bmx.search(b"", b"a")[0] == 0
    c, p, _ = bmx.search_device(torch.empty(0, dtype=torch.uint8, device=dev), b"a", max_positions=4)
    assert c == 0 and p.numel() == 0
    # m == n, m == 1, overlapping occurrences, NUL bytes, bytes >= 0x80
    assert list(bmx.search(b"abc", b"abc")[1]) == [0]
    assert list(bmx.search(b"aaaa", b"aaa")[1]) == [0, 1]
    assert list(bmx.search(b"aaaa", b"a")[1]) == [0, 1, 2, 3]
    assert list(bmx.search(b"\x00a\x00a", b"\x00a")[1]) == [0, 2]
    hi = bytes([200, 201, 200, 201, 200])
    assert list(bmx.search(hi, bytes([200, 201, 200]))[1]) == [0, 2]
    # hits at position 0 and n-m, for every variant and odd alignments
    for m in (1, 3, 4, 7, 11, 16, 33, 128):
        pat = bytes((i * 7 + 3) % 251 for i in range(m))
        body = bytes(255 for _ in range(5000))
        text = pat + body + pat
        want = oracle.search(text, pat)
        assert list(want) == [0, len(text) - m]
        for mis in (0, 1, 15):
            td = to_dev(text, dev, misalign=mis)
            for variant in VARIANTS_FOR(m):
                count, got, _ = gpu_positions(bmx, td, pat, variant)
                assert np.array_equal(got, want), (m, mis, variant)
    # empty pattern is rejected (the reference would report n+1 bogus hits)
    with pytest.raises(bmx.BmxError) as e:
        bmx.search(b"abc", b"")
    assert e.value.code == E.BMX_E_BADARG


def test_tile_seam_straddling_hits(bmx, oracle, dev):
    """Occurrences placed across every tile boundary (16 KiB and 32 KiB tiles) and around them."""
    n = 6 * 32768 + 777
    rng = np.random.default_rng(5)
    for m in (2, 4, 5, 7, 8, 11, 16, 31, 64, 128, 300):
        text = rng.integers(0, 256, size=n, dtype=np.uint8)
        pat = bytes(rng.integers(0, 256, size=m, dtype=np.uint8))
        p = np.frombuffer(pat, dtype=np.uint8)
        for seam in range(16384, n - m, 16384):
            for delta in (-m, -m + 1, -(m // 2), -3, -1, 0, 1, 13):
                o = seam + delta
                if 0 <= o <= n - m:
                    text[o:o + m] = p
        want = oracle.search(text.tobytes(), pat)
        for mis in (0, 5):
            td = to_dev(text, dev, misalign=mis)
            for variant in VARIANTS_FOR(m):
                count, got, _ = gpu_positions(bmx, td, pat, variant)
                assert count == want.size and np.array_equal(got, want), (m, mis, variant)


def test_tile_size_and_pipeline_knobs(bmx, oracle, dev):
    """Same answers for both tile sizes, shallow pipelines and one CTA per SM."""
    text = bmx.synth.fill_host(0, 3 << 20, 77, bmx.synth.ALPHABETS["dna"])
    pat = text[1000:1011].tobytes()
    want = oracle.search(text.tobytes(), pat)
    td = to_dev(text, dev)
    old = {k: os.environ.get(k) for k in ("BMX_TILE", "BMX_STAGES", "BMX_CTAS_PER_SM")}
    try:
        for tile, stages, ctas in [(16384, 2, 2), (32768, 3, 2), (32768, 8, 1), (16384, 8, 1)]:
            os.environ.update(BMX_TILE=str(tile), BMX_STAGES=str(stages), BMX_CTAS_PER_SM=str(ctas))
            for variant in ("qgram", "window", "shiftand"):
                count, got, stats = gpu_positions(bmx, td, pat, variant)
                assert stats["tile_bytes"] == tile and stats["stages"] <= stages
                assert count == want.size and np.array_equal(got, want), (tile, stages, ctas, variant)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)


def test_dense_hits_and_capacity(bmx, dev):
    """'aaaa...' / 'aaa': n-2 overlapping hits at 0..n-3; truncated output keeps the smallest."""
    n = (1 << 20) + 123
    td = torch.full((n,), ord("a"), dtype=torch.uint8, device=dev)
    for variant in ("auto", "shiftand"):
        count, pos, _ = bmx.search_device(td, b"aaa", max_positions=n, variant=variant)
        assert count == n - 2
        assert torch.equal(pos, torch.arange(n - 2, device=dev))
    count, pos, _ = bmx.search_device(td, b"aaa", max_positions=1000)
    assert count == n - 2 and torch.equal(pos, torch.arange(1000, device=dev))
    count, pos, _ = bmx.search_device(td, b"aaa")                 # count-only: no positions buffer
    assert count == n - 2 and pos is None
    # periodic pattern longer than the filter: every position verifies and matches
    count, pos, _ = bmx.search_device(td[:200000], b"a" * 64, max_positions=200000)
    assert count == 200000 - 63 and torch.equal(pos, torch.arange(200000 - 63, device=dev))


def test_long_patterns(bmx, oracle, dev):
    """Patterns beyond the shared-memory limits (tables and halo fall back to global memory)."""
    rng = np.random.default_rng(9)
    n = 1 << 20
    text = rng.integers(0, 4, size=n, dtype=np.uint8) + 65

    def find_all(hay: bytes, needle: bytes):
        out, i = [], hay.find(needle)
        while i >= 0:
            out.append(i)
            i = hay.find(needle, i + 1)
        return np.array(out, dtype=np.int64)

    for m in (1024, 1025, 2048, 4096, 5000, 70000):
        o = int(rng.integers(0, n - 2 * m - 7))
        pat = text[o:o + m].tobytes()
        text[o + m + 7: o + 2 * m + 7] = text[o:o + m]          # a second, adjacent copy
        # the oracle keeps the reference's O(m^3) table construction (BoyreMoore.cpp:165-190),
        # which does not finish for m in the thousands: beyond 2048 the checker is bytes.find
        want = oracle.search(text.tobytes(), pat) if m <= 2048 else find_all(text.tobytes(), pat)
        assert want.size >= 2
        count, got, _ = gpu_positions(bmx, to_dev(text, dev, misalign=3), pat)
        assert count == want.size and np.array_equal(got, want), m


def test_pos_base_and_determinism(bmx, oracle, dev):
    text = bmx.synth.fill_host(0, 2 << 20, 5, bmx.synth.ALPHABETS["ascii95"])
    pat = b"the quick brown fox"
    bmx.synth.plant_host(text, pat, bmx.synth.plant_offsets(text.size, len(pat), 300, 5))
    want = oracle.search(text.tobytes(), pat)
    td = to_dev(text, dev)
    base = (1 << 40) + 12345
    count, pos, _ = bmx.search_device(td, pat, max_positions=1000, pos_base=base)
    assert count == want.size and np.array_equal(pos.cpu().numpy(), want + base)
    runs = [bmx.search_device(td, pat, max_positions=1000)[1].cpu().numpy() for _ in range(5)]
    assert all(np.array_equal(r, want) for r in runs)


def test_partition_mirror_reproduces_reference_counts(bmx, oracle, golden):
    """bmx_search_partitions == the kernel entry: the reference's per-process counts (1649/1642)."""
    for name, pat, nparts, se, ans in golden.parts():
        t = golden.text(name)
        se2 = bmx.partition_words(t, nparts)
        assert np.array_equal(se2, se)
        bad, good = bmx.build_tables(pat)
        got = bmx.search_partitions(t, pat, se2, gs=good, bs=bad[:128])   # the reference passes its tables
        assert np.array_equal(got, ans), name
    t = golden.text("input5L")
    assert list(bmx.search_partitions(t, b"is", [0, 250037, 250039, 500006])) == [1649, 1642]
    rnd = random.Random(8)
    for _ in range(20):
        nparts = rnd.randint(1, 10)
        pat = rnd.choice([b"is", b"the", b"HACK", b"e", b"occurrences"])
        se = []
        for _p in range(nparts):
            a = rnd.randint(0, len(t) - 2)
            se += [a, rnd.randint(a - 1, len(t) - 1)]
        assert np.array_equal(bmx.search_partitions(t, pat, se), oracle.search_partitions(t, pat, se)), (pat, se)
    with pytest.raises(bmx.BmxError) as e:
        bmx.search_partitions(t, b"is", [0, 100], gs=[0, 5])
    assert e.value.code == bmx._lib.BMX_E_TABLES


def test_host_path_chunked_copy(bmx, oracle):
    """bmx_search with several H2D chunks: matches straddling chunk seams, pinned and pageable."""
    text = bmx.synth.fill_host(0, (5 << 20) + 999, 21, bmx.synth.ALPHABETS["dna"])
    pat = text[(1 << 20) - 5:(1 << 20) + 9].tobytes()            # straddles the first 1 MiB seam
    want = oracle.search(text.tobytes(), pat)
    old = os.environ.get("BMX_H2D_CHUNK_MB")
    os.environ["BMX_H2D_CHUNK_MB"] = "1"
    try:
        count, got = bmx.search(text, pat)
        assert count == want.size and np.array_equal(got, want)
        pinned = torch.from_numpy(text.copy()).pin_memory()
        count, got, stats = bmx.search(pinned, pat, return_stats=True)
        assert count == want.size and np.array_equal(got, want) and stats["kernel_launches"] >= 10
        count, got = bmx.search(pinned, pat, max_positions=0)    # count-only
        assert count == want.size and got.size == 0
        count, got = bmx.search(pinned, pat, max_positions=3)
        assert count == want.size and np.array_equal(got, want[:3])
    finally:
        os.environ.pop("BMX_H2D_CHUNK_MB", None) if old is None else os.environ.__setitem__("BMX_H2D_CHUNK_MB", old)


def test_scanner_chained_scans_keep_global_order(bmx, oracle, dev):
    text = bmx.synth.fill_host(0, 1 << 20, 31, bmx.synth.ALPHABETS["dna"])
    pat = text[5000:5008].tobytes()
    want = oracle.search(text.tobytes(), pat)
    td = to_dev(text, dev)
    out = torch.empty(want.size + 10, dtype=torch.int64, device=dev)
    s = bmx.Scanner(0)
    stream = torch.cuda.current_stream().cuda_stream
    s.set_pattern(pat, stream=stream)
    s.begin(out, stream=stream)
    m, cut = len(pat), 333_333
    s.scan(td[:cut], 0, stream=stream)                           # starts [0, cut-m]
    s.scan(td[cut - m + 1:], cut - m + 1, stream=stream)         # starts [cut-m+1, n-m]
    count, stats = s.finish(stream=stream)
    s.close()
    assert count == want.size and stats["kernel_launches"] == 4   # 2 x (scan, expand)
    assert np.array_equal(out[:count].cpu().numpy(), want)


def test_scanner_reuse_across_modes_sizes_and_patterns(bmx, oracle, dev):
    """One scanner, a long random sequence of searches: count-only and positions, one scan or several chained
    ones, growing and shrinking texts, changing patterns, searches abandoned without finish().  The scratch the
    kernels expect to find zeroed is cleaned by the kernels themselves (alternating halves), so every order of
    calls must leave it clean for the next one."""
    rnd = random.Random(99)
    s = bmx.Scanner(0)
    stream = torch.cuda.current_stream().cuda_stream
    big = np.random.default_rng(5).integers(97, 101, 9 << 20, dtype=np.uint8)
    td_big = to_dev(big, dev, misalign=3)
    out = torch.empty(big.size, dtype=torch.int64, device=dev)
    for it in range(70):
        n = rnd.choice([rnd.randint(1, 5000), rnd.randint(5000, 400_000), rnd.randint(400_000, big.size)])
        off = rnd.randint(0, big.size - n)
        m = rnd.choice([1, 2, 3, 5, 8, 12, 20])
        o = rnd.randint(0, big.size - m)
        pat = big[o:o + m].tobytes()
        text, td = big[off:off + n], td_big[off:off + n]
        want = oracle.search_np(text, pat, threads=-1) if n >= m else np.zeros(0, dtype=np.int64)
        positions = rnd.random() < 0.6
        s.set_pattern(pat, stream=stream)
        s.begin(out if positions else None, stream=stream)
        pieces = rnd.choice([1, 1, 2, 3])
        cuts = sorted(rnd.randint(0, n) for _ in range(pieces - 1))
        lo = 0
        for cut in cuts + [n]:
            # piece covering the starts [lo, cut - m] needs the bytes [lo, cut); the last piece runs to n
            hi = cut if cut == n else min(n, cut + m - 1)
            if hi - lo >= m:
                s.scan(td[lo:hi], lo, stream=stream)
            lo = max(lo, cut)
        if rnd.random() < 0.15:
            continue                                   # abandoned search: the next begin() starts over
        count, _ = s.finish(stream=stream)
        assert count == want.size, (it, n, m, positions, pieces)
        if positions:
            assert np.array_equal(out[:count].cpu().numpy(), want), (it, n, m, pieces)
    s.close()


def test_find_first_equals_first_position_of_the_serial_result(bmx, oracle, dev, monkeypatch):
    """bmx_find_first / bmx_find_first_device: the smallest start position, -1 without a
    match -- checked against the oracle's list and bytes.find, with matches in the first chunk, in later
    chunks, straddling chunk seams, at 0 and at n-m, and tiny chunks (test knob) that force many rounds."""
    rnd = random.Random(2024)
    for it in range(60):
        sigma = rnd.choice([2, 4, 26, 200])
        n = rnd.choice([rnd.randint(1, 300), rnd.randint(300, 300_000)])
        m = rnd.choice([1, 2, 3, 5, 8, 16, 33])
        text = bytearray(rnd.randrange(30, 30 + sigma) for _ in range(n))
        pat = bytes(rnd.randrange(30, 30 + sigma) for _ in range(m))
        mode = rnd.choice(["random", "late", "end", "start", "none"])
        if n >= m and mode != "none":
            at = {"random": rnd.randint(0, n - m), "late": max(0, n - m - rnd.randint(0, 50)), "end": n - m, "start": 0}[mode]
            text[at:at + m] = pat
        text = bytes(text)
        want_list = oracle.search(text, pat)
        want = int(want_list[0]) if want_list.size else -1
        assert want == text.find(pat)
        for kb in ("1", "7", None):          # host flavour: tiny H2D chunks force many chained find-first scans
            if kb is None:
                monkeypatch.delenv("BMX_H2D_CHUNK_KB", raising=False)
                monkeypatch.delenv("BMX_H2D_CHUNK_MB", raising=False)
            else:
                monkeypatch.setenv("BMX_H2D_CHUNK_KB", kb)
            td = to_dev(text, dev, misalign=rnd.randint(0, 17))
            assert bmx.find_first_device(td, pat) == want, (it, n, m, mode, kb, "device")
            assert bmx.find_first(text, pat) == want, (it, n, m, mode, kb, "host")
    # host flavour across several H2D chunks (1 MiB chunks): match only in the last chunk, then none at all
    monkeypatch.setenv("BMX_H2D_CHUNK_MB", "1")
    big = bmx.synth.fill_host(0, 5 * (1 << 20) + 77, 9, bmx.synth.ALPHABETS["ascii95"])
    pat = b"\x01needle\x02"
    assert bmx.find_first(big, pat) == -1
    for at in (0, (1 << 20) - 4, 3 * (1 << 20) + 5, big.size - len(pat)):
        t = big.copy()
        t[at:at + len(pat)] = np.frombuffer(pat, dtype=np.uint8)
        assert bmx.find_first(t, pat) == at
        assert bmx.find_first_device(to_dev(t, dev), pat) == at
    with pytest.raises(bmx.BmxError):
        bmx.find_first(b"abc", b"")
    # many matches spread over many tiles and CTAs: the device-side stop must still return the SMALLEST start
    monkeypatch.delenv("BMX_H2D_CHUNK_MB", raising=False)
    for seed, alphabet, m in ((3, "dna", 7), (4, "dna", 11), (5, "ascii95", 2)):
        t = bmx.synth.fill_host(0, 40 << 20, seed, bmx.synth.ALPHABETS[alphabet])
        for at in (39 << 20, 17 << 20, 5_000_001, 123_456, 0):
            pat = t[at:at + m].tobytes()
            want = t.tobytes().find(pat)
            td = to_dev(t, dev)
            for _ in range(3):       # repeated searches: the epoch-tagged key word is never cleared in between
                assert bmx.find_first_device(td, pat) == want
            assert bmx.find_first(t, pat) == want


def test_host_path_ring_layout_and_growing_position_buffer(bmx, oracle, monkeypatch):
    """Host text that "does not fit" (BMX_RESIDENT_MAX_MB=0): four ring slots of 64 KiB, the previous chunk's last
    m-1 bytes carried in front of every chunk.  Plus the position buffer that starts small and follows the count."""
    rnd = random.Random(77)
    base = bmx.synth.fill_host(0, (1 << 20) + 12345, 5, bmx.synth.ALPHABETS["dna"])
    for layout in ("ring", "resident"):
        if layout == "ring":
            monkeypatch.setenv("BMX_RESIDENT_MAX_MB", "0")
            monkeypatch.setenv("BMX_H2D_CHUNK_KB", "64")
        else:
            monkeypatch.delenv("BMX_RESIDENT_MAX_MB", raising=False)
            monkeypatch.setenv("BMX_H2D_CHUNK_KB", "256")
        for m in (1, 3, 14, 200, 5000):
            text = base.copy()
            at = (64 << 10) * rnd.randint(1, 10) - rnd.randint(0, m)        # straddles (or touches) a chunk seam
            pat = text[at:at + m].tobytes()
            for seam in range(1, 12):
                text[(seam << 16) - m // 2:(seam << 16) - m // 2 + m] = np.frombuffer(pat, dtype=np.uint8)
            want = oracle.search(text.tobytes(), pat) if m < 2000 else _find_all(text.tobytes(), pat)
            for src in (text, torch.from_numpy(text.copy()).pin_memory()):
                count, got = bmx.search(src, pat)
                assert count == want.size and np.array_equal(got, want), (layout, m)
                count, got = bmx.search(src, pat, max_positions=0)
                assert count == want.size
                count, got = bmx.search(src, pat, max_positions=5)
                assert count == want.size and np.array_equal(got, want[:5])
                assert bmx.find_first(src, pat) == (int(want[0]) if want.size else -1)
        # K patterns: the ring layout ingests the text once per pattern, the resident one once
        pats = [base[100:109].tobytes(), b"ACGTACGTAC", base[70000:70003].tobytes()]
        res = bmx.search_multi(base, pats)
        for (count, got), p in zip(res, pats):
            want = oracle.search(base.tobytes(), p)
            assert count == want.size and np.array_equal(got, want)
        # denser than the first guess of the position buffer (1 Mi entries): the buffer follows the count
        dense = np.full(3_000_000, ord("a"), dtype=np.uint8)
        dense[1_234_567] = ord("b")
        count, got = bmx.search(dense, b"aa")
        want = oracle.search(dense.tobytes(), b"aa")
        assert count == want.size and np.array_equal(got, want)
        count, got = bmx.search(dense, b"aa", max_positions=2_000_000)
        assert count == want.size and np.array_equal(got, want[:2_000_000])
        assert bmx._lib.load().bmx_release_memory(-1) == 0


def test_cached_buffers_threads_and_release(bmx, oracle, dev):
    """Threads come and go (each owns its cached buffers and gives them back when it exits), bmx_release_memory
    empties the caller's cache, and the next call simply allocates again."""
    import threading
    text = bmx.synth.fill_host(0, 3 << 20, 8, bmx.synth.ALPHABETS["ascii95"])
    pat = text[99_000:99_012].tobytes()
    want = oracle.search(text.tobytes(), pat)
    free0 = torch.cuda.mem_get_info()[0]
    errs = []

    def work():
        try:
            for _ in range(3):
                count, got = bmx.search(text, pat)
                assert count == want.size and np.array_equal(got, want)
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    for _ in range(6):
        ts = [threading.Thread(target=work) for _ in range(3)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    assert not errs, errs
    count, got = bmx.search(text, pat)
    assert count == want.size
    assert bmx._lib.load().bmx_release_memory(-1) == 0
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < (64 << 20), "exited threads and bmx_release_memory must give device memory back"
    count, got = bmx.search(text, pat)
    assert count == want.size and np.array_equal(got, want)


def test_pattern_cache_across_streams_and_patterns(bmx, oracle, dev):
    """bmx_search_device skips the pattern upload when the pattern is unchanged; another stream must still see it."""
    text = bmx.synth.fill_host(0, 2 << 20, 12, bmx.synth.ALPHABETS["dna"])
    td = to_dev(text.tobytes(), dev)
    pats = [text[5000:5012].tobytes(), text[77:85].tobytes(), b"ACGT"]
    wants = [oracle.search(text.tobytes(), p) for p in pats]
    streams = [torch.cuda.Stream(), torch.cuda.Stream(), None]
    rnd = random.Random(6)
    for it in range(60):
        k = rnd.randrange(3) if it % 3 else 0           # mostly the same pattern again
        st = streams[rnd.randrange(3)]
        count, got, _ = bmx.search_device(td, pats[k], max_positions=wants[k].size + 3, stream=st)
        (st or torch.cuda.current_stream()).synchronize()
        assert count == wants[k].size and np.array_equal(got.cpu().numpy(), wants[k]), (it, k)


def test_search_multi_shares_one_ingest(bmx, oracle, monkeypatch):
    """bmx_search_multi (first cut): K patterns, one host->device copy; every pattern's result
    equals its own serial search, including patterns longer than the text, duplicates and capped outputs."""
    rnd = random.Random(31337)
    for it, chunk_mb in enumerate([None, "1", None, "1", None]):
        if chunk_mb is None:
            monkeypatch.delenv("BMX_H2D_CHUNK_MB", raising=False)
        else:
            monkeypatch.setenv("BMX_H2D_CHUNK_MB", chunk_mb)
        sigma = rnd.choice([2, 4, 26])
        n = rnd.choice([rnd.randint(1, 2000), rnd.randint(2000, 3_500_000)])
        text = np.random.default_rng(it).integers(97, 97 + sigma, n, dtype=np.uint8)
        pats = []
        for m in [1, 2, 5, 8, 12, 40, n + 3]:
            o = rnd.randint(0, max(n - m, 0))
            pats.append(text[o:o + m].tobytes() if m <= n and rnd.random() < 0.7 else bytes(rnd.randrange(97, 97 + sigma) for _ in range(min(m, 1500))))
        pats.append(pats[2])                                             # a duplicate
        if n < 100_000:
            pats.insert(0, b"z" * (n + 1))                               # the ingest pattern itself longer than the text
        got = bmx.search_multi(text, pats)
        assert len(got) == len(pats)
        for pat, (count, pos) in zip(pats, got):
            want = oracle.search_np(text, pat, threads=-1) if len(pat) <= n else np.zeros(0, dtype=np.int64)
            assert count == want.size and np.array_equal(pos, want), (it, n, len(pat))
        capped = bmx.search_multi(text.tobytes(), pats, max_positions=3)
        for pat, (count, pos) in zip(pats, capped):
            want = oracle.search_np(text, pat, threads=-1) if len(pat) <= n else np.zeros(0, dtype=np.int64)
            assert count == want.size and np.array_equal(pos, want[:3])
    assert bmx.search_multi(b"abc", []) == []
    with pytest.raises(bmx.BmxError):
        bmx.search_multi(b"abc", [b"a", b""])


def test_multi_pattern_single_pass(bmx, oracle, dev):
    """K patterns in ONE pass (shared candidate table, union emission, per-pattern split) against
    K serial oracle searches: eligible sets (7 <= m <= 4096, K <= 64) through the device and the host entry point,
    patterns that are prefixes of each other (two patterns matching at the same start), duplicates, patterns
    sharing q-grams, capped and count-only outputs, a dense text that makes the union list and the outputs grow."""
    rnd = random.Random(4242)
    for it in range(10):
        sigma = rnd.choice([2, 4, 4, 26, 256])
        n = rnd.choice([rnd.randint(50, 3000), rnd.randint(100_000, 4_000_000)])
        text = np.random.default_rng(100 + it).integers(0, sigma, n, dtype=np.uint8) + (65 if sigma < 200 else 0)
        K = rnd.choice([1, 2, 5, 16, 64])
        pats = []
        while len(pats) < K:
            m = rnd.choice([7, 8, 9, 10, 11, 12, 16, 31, 32, 33, 64, 200])
            if m > n:
                continue
            o = rnd.randint(0, n - m)
            kind = rnd.random()
            if kind < 0.6:
                pats.append(text[o:o + m].tobytes())                            # occurs at least once
            elif kind < 0.75 and pats:
                base = rnd.choice(pats)                                          # a prefix of (or equal to) another pattern
                pats.append(base[: max(7, rnd.randint(7, len(base)))])
            elif kind < 0.85 and pats:
                base = bytearray(rnd.choice(pats))                               # shares most q-grams, differs at the end
                base[-1] ^= 1
                pats.append(bytes(base))
            else:
                pats.append(bytes(int(x) for x in np.random.default_rng(it * 977 + len(pats)).integers(0, sigma, m, dtype=np.uint8) + (65 if sigma < 200 else 0)))
        for p_ in pats[: 4]:                                                    # plant a few so that sparse alphabets have hits too
            o = rnd.randint(0, n - len(p_))
            text[o:o + len(p_)] = np.frombuffer(p_, dtype=np.uint8)
        wants = [oracle.search_np(text, p_, threads=-1) for p_ in pats]
        td = to_dev(text, dev, misalign=rnd.randint(0, 17))
        got = bmx.search_multi_device(td, pats, max_positions=n)
        for k, ((count, pos), want) in enumerate(zip(got, wants)):
            assert count == want.size and np.array_equal(pos.cpu().numpy(), want), (it, k, len(pats[k]), "device")
        capped = bmx.search_multi_device(td, pats, max_positions=2)
        counted = bmx.search_multi_device(td, pats, max_positions=0)
        host = bmx.search_multi(text, pats)
        for k, want in enumerate(wants):
            assert capped[k][0] == want.size and np.array_equal(capped[k][1].cpu().numpy(), want[:2])
            assert counted[k][0] == want.size and counted[k][1] is None
            assert host[k][0] == want.size and np.array_equal(host[k][1], want), (it, k, "host")
    # dense: every start matches both patterns; the union list and the per-pattern buffers start small and must grow
    dense = np.full(2_500_000, ord("a"), dtype=np.uint8)
    dense[2_000_000] = ord("b")
    pats = [b"a" * 7, b"a" * 12, b"aaaaaab"]
    wants = [oracle.search_np(dense, p_, threads=-1) for p_ in pats]
    for k, (count, pos) in enumerate(bmx.search_multi(dense, pats)):
        assert count == wants[k].size and np.array_equal(pos, wants[k])
    got = bmx.search_multi_device(to_dev(dense, dev), pats, max_positions=dense.size)
    for k, (count, pos) in enumerate(got):
        assert count == wants[k].size and np.array_equal(pos.cpu().numpy(), wants[k])
    # an ineligible set (a 3-byte pattern) takes the pattern-by-pattern route with the same results
    mixed = [b"aaa", b"a" * 9]
    for k, (count, pos) in enumerate(bmx.search_multi_device(to_dev(dense, dev), mixed, max_positions=10)):
        want = oracle.search_np(dense, mixed[k], threads=-1)
        assert count == want.size and np.array_equal(pos.cpu().numpy(), want[:10])


def test_abi_device_entry_point_raw(bmx, oracle, dev):
    """bmx_search_device exactly as a C caller would use it (ctypes, raw pointers)."""
    lib = bmx._lib.load()
    text = bmx.synth.fill_host(0, 1 << 18, 3, bmx.synth.ALPHABETS["ascii95"])
    pat = text[777:777 + 16].tobytes()
    want = oracle.search(text.tobytes(), pat)
    td = to_dev(text, dev)
    out = torch.empty(64, dtype=torch.int64, device=dev)
    cnt, ms = ctypes.c_uint64(), ctypes.c_float()
    rc = lib.bmx_search_device(ctypes.c_void_p(td.data_ptr()), td.numel(), pat, len(pat), ctypes.c_void_p(out.data_ptr()),
                               64, ctypes.byref(cnt), ctypes.byref(ms), None)
    assert rc == 0 and cnt.value == want.size and ms.value > 0
    assert np.array_equal(out[: cnt.value].cpu().numpy(), want)


def test_refmain_demo_prints_the_reference_console_lines(bmx, golden, tmp_path):
    """bmx_refmain = BoyreMoore.cpp main() on libbmx.so: same files, same lines, same counts."""
    import re
    import subprocess
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import build as b

    exe = bmx.LIB_PATH.parent / "bmx_refmain"
    if not exe.exists():
        b.build()
    (tmp_path / "inputEd.txt").write_bytes(golden.text("input5L"))
    (tmp_path / "input1Search.txt").write_bytes(b"is")          # the reference's own pattern file
    out = subprocess.run([str(exe)], cwd=tmp_path, capture_output=True, timeout=300, check=True).stdout.decode("latin-1")
    assert out.count("The no. of occurrences by process 0 is 1649") == 10     # 10 timed runs
    assert out.count("The no. of occurrences by process 1 is 1642") == 10
    found = [(int(i), int(p)) for i, p in re.findall(r"Found by (\d+) at : (\d+)", out)]
    want = [p for _t, pat, p, _s in [c for c in golden.cases() if c[0] == "input5L" and c[1] == b"is"]][0]
    assert [p for _i, p in found] == list(want)                               # all 3291, ascending
    assert all((i == 0) == (p <= 250037) for i, p in found)                   # printed by the owning process
    assert "Serial result (one range, (m-1)-byte halos): 3291 occurrences" in out
    assert "Average time" in out


def test_exchange_step_with_device_header(bmx, oracle, dev):
    """The multi-GPU step as bench.py runs it (scan -> export_result -> collectives), on a
    one-rank NCCL group: no host sync between the scan and the exchange."""
    import torch.distributed as dist
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd

    text = bmx.synth.fill_host(0, 1 << 20, 61, bmx.synth.ALPHABETS["dna"])
    pat = text[4242:4250].tobytes()
    want = oracle.search(text.tobytes(), pat)
    td = to_dev(text, dev)
    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1, device_id=dev)
    try:
        s = bmx.Scanner(0)
        stream = torch.cuda.current_stream().cuda_stream
        s.set_pattern(pat, stream=stream)
        for cap, fast_cap in [(want.size + 7, 4096), (want.size + 7, 3), (5, 4096)]:
            pos = torch.empty(cap, dtype=torch.int64, device=dev)
            packed = torch.zeros(2 + fast_cap, dtype=torch.int64, device=dev)
            s.begin(pos, stream=stream)
            s.scan(td, 1000, stream=stream)
            s.export_result(packed, stream=stream)
            pending = bd.combine_hits_start(None, pos, device=dev, packed=packed, fast_cap=fast_cap)
            total, counts, gathered = pending.finish()
            assert total == want.size and counts == [want.size]
            assert np.array_equal(gathered.cpu().numpy(), (want + 1000)[:cap])
        s.close()
    finally:
        if created:
            dist.destroy_process_group()


def test_mid_density_and_mixed_blocks(bmx, oracle, dev):
    """Texts whose 2 MiB blocks are dense with PARTIAL hit masks (ticket-ordered CTA expand, staged
    stores), sparse, or a mix of both; capacity truncation inside a dense block."""
    rng = np.random.default_rng(77)
    n = (6 << 20) + 4321
    binary = rng.integers(0, 2, size=n, dtype=np.uint8) + 97           # 'a'/'b': 25 % of starts match 'ab'
    mixed = rng.integers(0, 256, size=n, dtype=np.uint8)
    mixed[: (2 << 20) + 999] = binary[: (2 << 20) + 999]               # dense block, then sparse, then dense again
    mixed[-(1 << 20):] = ord("a")
    periodic = np.frombuffer((b"abc" * (n // 3 + 1))[:n], dtype=np.uint8).copy()
    for text, pats in [(binary, [b"ab", b"aab", b"abab", b"abbabaab"]), (mixed, [b"ab", b"aa", b"aaaaaaaaa"]),
                       (periodic, [b"abc", b"bcab", b"cabcabcabcab"])]:
        td = to_dev(text, dev, misalign=3)
        for pat in pats:
            want = oracle.search_np(text, pat, threads=-1)
            assert want.size > 1000
            count, pos, _ = bmx.search_device(td, pat, max_positions=n)
            assert count == want.size and np.array_equal(pos.cpu().numpy(), want), pat
            cap = want.size // 2 + 1
            count, pos, _ = bmx.search_device(td, pat, max_positions=cap)
            assert count == want.size and np.array_equal(pos.cpu().numpy(), want[:cap]), (pat, "truncated")
            count, _, _ = bmx.search_device(td, pat)
            assert count == want.size


def test_expand_probe_layouts(bmx, oracle, dev, monkeypatch):
    """The expand kernel probes one item flag per lane when one round of loads covers the text and four
    flags per lane on multi-GiB texts.  A tiny expand grid (test knob) forces the four-flag layout and many
    probe rounds on a small text of every density; the result must not change."""
    rnd = random.Random(4242)
    n = 6 * (1 << 20) + 12345
    for sigma, m in [(2, 3), (4, 4), (4, 9), (26, 2), (95, 8)]:
        text = np.random.default_rng(sigma * 100 + m).integers(97, 97 + sigma, n, dtype=np.uint8).tobytes() if sigma < 95 else \
            bmx.synth.fill_host(0, n, 5, bmx.synth.ALPHABETS["ascii95"]).tobytes()
        o = rnd.randint(0, n - m)
        pat = text[o:o + m]
        want = oracle.search_np(np.frombuffer(text, dtype=np.uint8), pat, threads=-1)
        td = to_dev(text, dev, misalign=rnd.randint(0, 17))
        for grid in ("1", "3", None):
            if grid is None:
                monkeypatch.delenv("BMX_EXPAND_GRID", raising=False)
            else:
                monkeypatch.setenv("BMX_EXPAND_GRID", grid)
            count, got, _ = gpu_positions(bmx, td, pat)
            assert count == want.size and np.array_equal(got, want), (sigma, m, grid)


def test_concurrent_host_threads_are_independent(bmx, oracle, dev):
    """The convenience entry points keep per-thread state: searches for different patterns from
    several host threads at once must not disturb each other (the reference is single-threaded;
    the ABI promises re-entrancy per thread)."""
    import threading

    text = bmx.synth.fill_host(0, 4 << 20, 91, bmx.synth.ALPHABETS["dna"])
    td = to_dev(text, dev)
    pats = [text[o:o + m].tobytes() for o, m in [(100, 7), (5000, 9), (77777, 12), (123456, 33), (9, 3), (31, 5)]]
    want = [oracle.search(text.tobytes(), p) for p in pats]
    errors = []

    def worker(i):
        try:
            for rep in range(6):
                if rep % 2:
                    with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                        c, pos, _ = bmx.search_device(td, pats[i], max_positions=want[i].size + 5)
                        got = pos.cpu().numpy()
                else:
                    c, got = bmx.search(text, pats[i])
                if c != want[i].size or not np.array_equal(got, want[i]):
                    errors.append((i, rep, c, want[i].size))
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(pats))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_fuzz_all_densities_and_capacities(bmx, oracle, dev):
    """Random, periodic and mixed texts from 1 byte to a few MiB over alphabets of 1..256 symbols: every
    emission path (solo lane, warp-cooperative, staged, full segments, dense tickets) with full, halved
    and single-entry capacities.  (profiles/stress_fuzz.py is the long-running version of this test.)"""
    rnd = random.Random(2024)
    rng = np.random.default_rng(2024)
    for it in range(90):
        sigma = rnd.choice([1, 2, 2, 3, 4, 8, 26, 256])
        n = rnd.choice([rnd.randint(1, 4000), rnd.randint(4000, 300000), rnd.randint(300000, 3 << 20)])
        off = 60 if sigma < 190 else 0
        kind = rnd.choice(["random", "random", "periodic", "mixed"])
        if kind == "periodic":
            unit = bytes(rnd.randrange(sigma) + off for _ in range(rnd.randint(1, 7)))
            text = np.frombuffer((unit * (n // len(unit) + 1))[:n], dtype=np.uint8).copy()
        else:
            text = (rng.integers(0, sigma, size=n, dtype=np.uint8) + off).astype(np.uint8)
            if kind == "mixed" and n > 1000:
                a, b = sorted(rnd.sample(range(n), 2))
                text[a:b] = text[a]
        m = min(rnd.choice([1, 2, 3, 4, 5, 7, 8, 10, 11, 16, 25, 33, 64, 128, 300]), n)
        if rnd.random() < 0.7:
            o = rnd.randint(0, n - m)
            pat = text[o:o + m].tobytes()
        else:
            pat = bytes(rnd.randrange(sigma) + off for _ in range(m))
        want = oracle.search_np(text, pat, threads=4) if n > 100000 else oracle.search(text.tobytes(), pat)
        td = to_dev(text, dev, misalign=rnd.randint(0, 20))
        for variant in VARIANTS_FOR(m):
            cap = rnd.choice([max(n, 1), max(want.size // 2, 1), 1])
            count, pos, _ = bmx.search_device(td, pat, max_positions=cap, variant=variant)
            assert count == want.size and np.array_equal(pos.cpu().numpy(), want[:cap]), (it, kind, sigma, n, m, variant, cap)


def test_single_process_multi_gpu_entry_point(bmx, oracle, dev):
    """bmx_mg_search on however many GPUs are visible (1 on the test box; the shard logic is the same
    and profiles/mg_check.py exercises it on 2+): seam-straddling plants, capacities, count-only."""
    text = bmx.synth.fill_host(0, (9 << 20) + 77, 55, bmx.synth.ALPHABETS["dna"])
    m = 11
    pat = text[3000:3000 + m].tobytes()
    for ngpus in (1, 0):
        mg = bmx.MultiGpu(ngpus)
        per = -(-text.size // mg.ngpus)
        per = -(-per // 16) * 16
        t2 = text.copy()
        seams = [r * per for r in range(1, mg.ngpus)]
        bmx.synth.plant_host(t2, pat, [s - d for s in seams for d in (m // 2, 1, m - 1, m, 0)])
        want = oracle.search_np(t2, pat, threads=4)
        count, pos, shard = mg.search(t2, pat)
        assert count == want.size == sum(shard) and np.array_equal(pos, want)
        count, pos, _ = mg.search(t2, pat, max_positions=7)
        assert count == want.size and np.array_equal(pos, want[:7])
        count, pos, _ = mg.search(t2, pat, max_positions=0)
        assert count == want.size and pos.size == 0
        assert mg.search(b"ab", b"abc")[0] == 0
        mg.close()


def test_cooperative_verification_edges(bmx, oracle, dev, monkeypatch):
    """The sparse path's flagged chunks are checked by the whole warp (coop_verify16): pattern lengths around its
    word and round boundaries (8 bytes per ballot; 128 bytes per survivor step; the shared-memory pattern limit),
    candidates that share their first 8 bytes with the pattern and differ later, occurrences at both ends of the
    text and across tile seams, overlapping occurrences, and periodic texts where every chunk is flagged.  The
    lane-walk with BM skips (BMX_COOP_VERIFY=0) must give the same list."""
    rng = np.random.default_rng(4242)
    n = (3 << 20) + 77
    base = rng.integers(0, 4, size=n, dtype=np.uint8) + 65                 # ACGT-like: false candidates galore
    for m in (5, 6, 7, 8, 9, 11, 12, 13, 31, 32, 33, 36, 37, 135, 136, 137, 140, 141, 264, 265, 300, 1023, 1024):
        text = base.copy()
        o = int(rng.integers(1000, n // 2))
        pat = text[o:o + m].tobytes()
        near = bytearray(pat)
        near[-1] ^= 1                                                       # differs in the last byte only
        mid = bytearray(pat)
        mid[min(m - 1, 8)] ^= 1                                             # differs right behind the first 8 bytes
        spots = [0, 32768 - m + 3, 32768 - 2, 65536 - 1, (1 << 20) + 5, n - m]    # text ends, tile seams
        for k, s in enumerate(spots):
            text[s:s + m] = np.frombuffer(pat, dtype=np.uint8)
            d = s + 3 * m + 64 + k
            if d + m < n - m - 64:
                text[d:d + m] = np.frombuffer(bytes(near if k % 2 else mid), dtype=np.uint8)
        if m <= 300:                                                        # overlapping copies of a periodic pattern
            text[200000:200000 + 5 * m] = np.frombuffer((pat[: max(1, m // 3)] * (5 * m))[: 5 * m], dtype=np.uint8)
        want = oracle.search_np(text, pat, threads=-1)
        assert want.size >= 4                                               # (neighbouring spots may overwrite each other)
        for misalign in (0, 5):
            td = to_dev(text, dev, misalign=misalign)
            for coop in ("1", "0"):
                monkeypatch.setenv("BMX_COOP_VERIFY", coop)
                count, got, _ = gpu_positions(bmx, td, pat, cap=want.size + 8)
                assert count == want.size and np.array_equal(got, want), (m, misalign, coop)
                count, _, _ = bmx.search_device(td, pat)
                assert count == want.size, (m, misalign, coop, "count-only")
    monkeypatch.delenv("BMX_COOP_VERIFY")
    # every chunk of every segment flagged: 'ab' * k searched for (ab)^4 and (ab)^20, and one byte off
    per = np.frombuffer((b"ab" * (1 << 19)), dtype=np.uint8).copy()
    per[777777] = ord("c")
    td = to_dev(per, dev, misalign=1)
    for pat in (b"abababab", b"ab" * 20, b"babababa" + b"b"):
        want = oracle.search_np(per, pat, threads=-1)
        count, got, _ = gpu_positions(bmx, td, pat, cap=want.size + 1)
        assert count == want.size and np.array_equal(got, want), pat


def test_small_host_text_latency_path(bmx, oracle, monkeypatch):
    """Host texts of at most 4 MiB take the one-stream path with a speculative read-back of the first 8192 positions
    (bmx_host.cu: small_host_search).  Its branches: fewer hits than the read-back, more hits than the read-back, more
    hits than the device buffer (re-scan of the resident copy), count-only, a truncating caller buffer, pinned and
    pageable sources, sizes around the limit -- each against the oracle and against the chunked path."""
    rng = np.random.default_rng(314)
    cases = []
    sparse = rng.integers(0, 95, size=700_001, dtype=np.uint8) + 32
    pat = b"needle in a haystack"
    for o in (0, 5000, 123_456, 700_001 - len(pat)):
        sparse[o:o + len(pat)] = np.frombuffer(pat, dtype=np.uint8)
    cases.append((sparse, pat))                                              # a handful of hits
    cases.append((rng.integers(0, 2, size=300_000, dtype=np.uint8) + 97, b"ab"))     # ~75 000 hits: beyond the read-back
    cases.append((np.full(3_000_000, ord("a"), dtype=np.uint8), b"aa"))              # 3 M hits: beyond the first device buffer
    cases.append((rng.integers(0, 4, size=(4 << 20), dtype=np.uint8) + 65, b"ACGTAC"))   # exactly the size limit
    cases.append((rng.integers(0, 4, size=(4 << 20) + 1, dtype=np.uint8) + 65, b"ACGTAC"))   # one byte more: the chunked path
    cases.append((np.frombuffer(b"abc", dtype=np.uint8).copy(), b"abcd"))            # m > n
    for text, p in cases:
        want = oracle.search_np(text, p, threads=-1) if text.size >= len(p) else np.zeros(0, dtype=np.int64)
        pinned = torch.from_numpy(text.copy()).pin_memory()
        for src in (text, pinned):
            for small in ("1", "0"):
                monkeypatch.setenv("BMX_SMALL_HOST", small)
                count, got = bmx.search(src, p)
                assert count == want.size and np.array_equal(got, want), (text.size, p, small)
                count, got = bmx.search(src, p, max_positions=0)
                assert count == want.size and got.size == 0
                cap = max(1, want.size // 3)
                count, got = bmx.search(src, p, max_positions=cap)
                assert count == want.size and np.array_equal(got, want[:cap]), (text.size, p, small, "truncated")
    monkeypatch.delenv("BMX_SMALL_HOST")

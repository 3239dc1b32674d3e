"""GPU parity at BASELINE.json's full sizes.  The text is generated on the device from the
counter-based stream, copied to the host once, and the oracle (multi-threaded windowed driver
over the restated reference loop) gives the truth for the WHOLE text: identical count and
identical ascending position list, bit for bit.  Size-independent properties (sortedness, every
reported start really matches, count-only == positions, closed form for the periodic case) are
checked on top."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GIB = 1 << 30


@pytest.fixture(scope="module")
def dev(bmx):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def make_text(bmx, dev, n, alphabet, seed):
    t = torch.empty(n, dtype=torch.uint8, device=dev)
    bmx.synth.fill_device(t, 0, seed, bmx.synth.ALPHABETS[alphabet])
    return t


def check_against_oracle(bmx, oracle, text_dev, pat, variant="auto", pos_base=0):
    n = text_dev.numel()
    count, pos, stats = bmx.search_device(text_dev, pat, max_positions=1 << 20, pos_base=pos_base, variant=variant)
    count_only, none, _ = bmx.search_device(text_dev, pat, variant=variant)
    assert none is None and count_only == count
    host = text_dev.cpu().numpy()
    want = oracle.search_np(host, pat, threads=-1)
    assert count == want.size, (count, want.size)
    got = pos.cpu().numpy()
    assert np.array_equal(got, want + pos_base)
    # properties that do not need the oracle
    assert np.all(np.diff(got) > 0)
    idx = (pos - pos_base).unsqueeze(1) + torch.arange(len(pat), device=text_dev.device).unsqueeze(0)
    pt = torch.frombuffer(bytearray(pat), dtype=torch.uint8).to(text_dev.device)
    assert bool((text_dev[idx] == pt).all().item())
    return count, stats


def test_config1_ascii_64MiB_m16(bmx, oracle, dev):
    """BASELINE configs[0]: 64 MiB of 95-symbol ASCII, 16-byte pattern, 1000 plants, seed 42."""
    n, m, seed = 64 << 20, 16, 42
    alpha = bmx.synth.ALPHABETS["ascii95"]
    t = make_text(bmx, dev, n, "ascii95", seed)
    pat = bmx.synth.pattern_from_stream(m, seed, alpha)
    bmx.synth.plant_device(t, pat, bmx.synth.plant_offsets(n, m, 1000, seed))
    count, _ = check_against_oracle(bmx, oracle, t, pat)
    assert 900 <= count <= 1000
    # host-pointer entry point (H2D + scan + D2H) on the same bytes
    host = t.cpu().numpy()
    c2, p2 = bmx.search(host, pat)
    assert c2 == count and np.array_equal(p2, oracle.search_np(host, pat, threads=-1))
    # and against the reference's OWN code (kernel1.cl + BoyreMoore.cpp tables through the shim), which is
    # defined on this input (7-bit bytes, m <= 99, n < 2^31): the serial result, one partition {0, n-1}
    import ctypes
    from pathlib import Path
    so = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "libref_bm.so"
    if so.exists():
        ref = ctypes.CDLL(str(so))
        want = np.zeros(count + 16, dtype=np.int64)
        rc_count = ctypes.c_uint64()
        rc = ref.ref_bm_search(ctypes.c_void_p(host.ctypes.data), ctypes.c_int64(n), ctypes.c_char_p(pat), ctypes.c_int32(m),
                               want.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(want.size), ctypes.byref(rc_count))
        assert rc == 0 and rc_count.value == count
        assert np.array_equal(want[:count], p2)


def test_config2_dna_4GiB_m32(bmx, oracle, dev):
    """BASELINE configs[1]: 4 GiB random DNA, m = 32 cut from the text + 1000 plants, seed 43."""
    n, m, seed = 4 * GIB, 32, 43
    alpha = bmx.synth.ALPHABETS["dna"]
    t = make_text(bmx, dev, n, "dna", seed)
    off = bmx.synth.mix64(seed * 7919) % (n - m)
    pat = bmx.synth.fill_host(off, m, seed, alpha).tobytes()
    bmx.synth.plant_device(t, pat, bmx.synth.plant_offsets(n, m, 1000, seed))
    count, stats = check_against_oracle(bmx, oracle, t, pat)
    assert count >= 1001 and stats["variant"] == "qgram"
    del t
    torch.cuda.empty_cache()


def test_config3_bytes256_4GiB_m4_m16_m128(bmx, oracle, dev):
    """BASELINE configs[2]: 4 GiB of uniform bytes 0..255, m in {4, 16, 128}, 1000 plants each."""
    n = 4 * GIB
    alpha = bmx.synth.ALPHABETS["bytes256"]
    t = make_text(bmx, dev, n, "bytes256", 44)
    for m, seed, variant in [(4, 44, "window"), (16, 45, "qgram"), (128, 46, "qgram")]:
        pat = bmx.synth.pattern_from_stream(m, seed, alpha)
        bmx.synth.plant_device(t, pat, bmx.synth.plant_offsets(n, m, 1000, seed))
        count, stats = check_against_oracle(bmx, oracle, t, pat)
        assert count >= 990 and stats["variant"] == variant
    # 7-bit twin where the verbatim reference code is defined too (sigma = 128, m = 99 = its maximum)
    bmx.synth.fill_device(t[: GIB], 0, 50, bmx.synth.ALPHABETS["ascii128"])
    pat = bmx.synth.pattern_from_stream(99, 52, bmx.synth.ALPHABETS["ascii128"])
    bmx.synth.plant_device(t[: GIB], pat, bmx.synth.plant_offsets(GIB, 99, 500, 52))
    check_against_oracle(bmx, oracle, t[: GIB], pat)
    del t
    torch.cuda.empty_cache()


def test_config4_periodic_1GiB_aaa(bmx, oracle, dev):
    """BASELINE configs[3]: 1 GiB of 'a', pattern 'aaa': n-2 overlapping hits at 0..n-3."""
    n = GIB
    t = torch.full((n,), ord("a"), dtype=torch.uint8, device=dev)
    pos = torch.empty(n, dtype=torch.int64, device=dev)
    count, got, _ = bmx.search_device(t, b"aaa", pos_out=pos)
    assert count == n - 2 == got.numel()
    # closed form, checked without materialising arange: first/last, constant stride, checksum
    assert int(got[0]) == 0 and int(got[-1]) == n - 3
    assert bool((got[1:] - got[:-1] == 1).all().item())
    count_only, _, _ = bmx.search_device(t, b"aaa")
    assert count_only == n - 2
    # the oracle on a 64 MiB window agrees with the same window of the GPU list
    w = 64 << 20
    want = oracle.search_np(t[:w].cpu().numpy(), b"aaa", threads=-1)
    assert np.array_equal(want, got[: w - 2].cpu().numpy())
    # truncated output keeps the smallest positions
    c3, p3, _ = bmx.search_device(t, b"aaa", max_positions=12345)
    assert c3 == n - 2 and torch.equal(p3, got[:12345])
    del t, pos, got
    torch.cuda.empty_cache()


def test_config5_shard_8GiB_ascii_m64(bmx, oracle, dev):
    """One rank's share of BASELINE configs[4]: 8 GiB of 95-symbol ASCII at a shard offset, m = 64,
    plants including ones clipped by the shard edges (as at a seam), global positions reported."""
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd
    world, rank, per = 8, 3, 8 * GIB
    total = per * world
    m, seed = 64, 47
    alpha = bmx.synth.ALPHABETS["ascii95"]
    lo, hi = bd.shard_bounds(total, world, rank)
    lo, end = bd.shard_read_range(total, m, lo, hi)
    assert (lo, hi, end) == (3 * per, 4 * per, 4 * per + m - 1)
    t = torch.empty(end - lo, dtype=torch.uint8, device=dev)
    bmx.synth.fill_device(t, lo, seed, alpha)
    pat = bmx.synth.pattern_from_stream(m, seed, alpha)
    plants = list(bmx.synth.plant_offsets(per, m, 1000, seed, lo=lo))
    plants += [lo - m // 2, lo, hi - m // 2, hi - 1, hi]   # straddling both seams; the last one belongs to rank 4
    bmx.synth.plant_device(t, pat, plants, base=lo)
    count, _ = check_against_oracle(bmx, oracle, t, pat, pos_base=lo)
    assert count >= 1000
    del t
    torch.cuda.empty_cache()

"""Synthetic texts of the BASELINE configurations (tests and bench share these definitions).

The byte stream is defined once (oracle/bm_oracle.c:oracle_synth_fill == bmx_synth_fill_device):
8 bytes per splitmix64 draw, draw j = mix(seed + (j+1)*GOLDEN), byte k of the draw
b = (z >> 8k) & 0xFF mapped to alphabet[(b*sigma) >> 8].  Plants overwrite m bytes at offsets
drawn from the same mixer.  Nothing here computes matches.
"""
from __future__ import annotations

import ctypes

import numpy as np

GOLDEN = 0x9E3779B97F4A7C15
MASK64 = (1 << 64) - 1

ALPHABETS = {
    "dna": b"ACGT",
    "ascii95": bytes(range(0x20, 0x7F)),
    "ascii128": bytes(range(128)),
    "bytes256": bytes(range(256)),
    "a": b"a",
}


def mix64(z: int) -> int:
    z &= MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def fill_host(offset: int, length: int, seed: int, alphabet: bytes) -> np.ndarray:
    """numpy twin of oracle_synth_fill (vectorised; for tests that must not load the oracle)."""
    sigma = len(alphabet)
    first = offset >> 3
    last = (offset + length + 7) >> 3
    j = np.arange(first, last, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & MASK64) + (j + np.uint64(1)) * np.uint64(GOLDEN)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    b = z.view(np.uint8).reshape(-1, 8).astype(np.uint32)  # little-endian host
    table = np.frombuffer(alphabet, dtype=np.uint8)
    out = table[(b * np.uint32(sigma)) >> np.uint32(8)].reshape(-1)
    lo = offset - (first << 3)
    return np.ascontiguousarray(out[lo: lo + length])


def fill_device(text, offset: int, seed: int, alphabet: bytes, stream=None) -> None:
    """Fill a CUDA uint8 tensor in place with stream bytes [offset, offset+len)."""
    import torch

    from . import _lib
    lib = _lib.load()
    s = stream if stream is not None else torch.cuda.current_stream(text.device)
    with torch.cuda.device(text.device):
        _lib.check(lib.bmx_synth_fill_device(ctypes.c_void_p(text.data_ptr()), offset, text.numel(),
                                             ctypes.c_uint64(seed & MASK64), alphabet, len(alphabet),
                                             ctypes.c_void_p(s.cuda_stream)))


def plant_offsets(n: int, m: int, count: int, seed: int, lo: int = 0) -> np.ndarray:
    """`count` plant offsets in [lo, lo + n - m], deterministic in (seed, i)."""
    span = n - m + 1
    if span <= 0 or count <= 0:
        return np.zeros(0, dtype=np.int64)
    return np.array([lo + mix64(seed * 0x100000001B3 + i) % span for i in range(count)], dtype=np.int64)


def pattern_from_stream(m: int, seed: int, alphabet: bytes) -> bytes:
    """A random pattern over the alphabet (its own stream, far from the text's seed)."""
    return fill_host(0, m, seed ^ 0xA5A5A5A5DEADBEEF, alphabet).tobytes()


def plant_host(text: np.ndarray, pattern: bytes, offsets, base: int = 0) -> None:
    p = np.frombuffer(pattern, dtype=np.uint8)
    n = text.size
    for o in offsets:
        o = int(o) - base
        a, b = max(o, 0), min(o + p.size, n)
        if a < b:
            text[a:b] = p[a - o: b - o]


def plant_device(text, pattern: bytes, offsets, base: int = 0) -> None:
    """Overwrite the pattern at each (global) offset; `base` = global offset of text[0].
    Plants straddling the buffer edges are clipped, so neighbouring shards agree on the halo."""
    import torch

    p = torch.frombuffer(bytearray(pattern), dtype=torch.uint8).to(text.device)
    n = text.numel()
    for o in offsets:
        o = int(o) - base
        a, b = max(o, 0), min(o + p.numel(), n)
        if a < b:
            text[a:b] = p[a - o: b - o]

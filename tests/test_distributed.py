"""CPU, world_size 2 and 3 over gloo: the N>1 host logic (shard ownership, halo, count exchange,
position gather to rank 0).  The local scan is injected -- here the oracle plays the device -- so
the exchange code that runs over NCCL on the GPU box is the code exercised here."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

torch = pytest.importorskip("torch")


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_total: int, m: int, seed: int, out_dir: str, fast_cap: int):
    import torch.distributed as dist

    from conftest import load_oracle
    import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = load_oracle()
    alpha = bmx.synth.ALPHABETS["dna"]
    pat = bmx.synth.fill_host(12345, m, seed, alpha).tobytes()

    # every rank materialises only its own bytes [lo, end) from the counter-based generator
    lo, hi = bd.shard_bounds(n_total, world, rank)
    lo, end = bd.shard_read_range(n_total, m, lo, hi)
    shard = bmx.synth.fill_host(lo, end - lo, seed, alpha)
    plants = []
    for r in range(1, world):          # occurrences straddling every shard seam, and touching it
        seam, _ = bd.shard_bounds(n_total, world, r)
        plants += [seam - m // 2, seam - 1, seam - m + 1, seam - m, seam]
    plants += list(bmx.synth.plant_offsets(n_total, m, 40, seed))
    bmx.synth.plant_host(shard, pat, plants, base=lo)

    def oracle_scan(text, pattern, pos_base, cap):
        got = oracle.search(text.numpy().tobytes(), pattern) + pos_base
        return int(got.size), torch.from_numpy(got[:cap].copy()), {"variant": "oracle"}

    if fast_cap < 0:
        # the pipelined form bench.py uses: the scan leaves {count, held, head of the list} packed in one buffer
        # (Scanner.export_result on the GPU), the exchange is enqueued on it and awaited later
        fast_cap = -fast_cap
        count, pos, _ = oracle_scan(torch.from_numpy(shard), pat, lo, 1 << 20)
        packed = torch.zeros(2 + fast_cap, dtype=torch.int64)
        packed[0], packed[1] = count, pos.numel()
        packed[2: 2 + min(pos.numel(), fast_cap)] = pos[:fast_cap]
        pending = bd.combine_hits_start(None, pos, device=torch.device("cpu"), fast_cap=fast_cap, packed=packed)
        total, counts, gathered = pending.finish()
    else:
        total, counts, gathered, _ = bd.sharded_search(torch.from_numpy(shard), lo, pat, max_positions=1 << 20, fast_cap=fast_cap,
                                                       local_scan=oracle_scan)
    np.save(os.path.join(out_dir, f"count_{rank}.npy"), np.array([total] + counts, dtype=np.int64))
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,fast_cap", [(2, 4096), (3, 4096), (2, 5), (2, -4096), (3, -7)])
def test_sharded_search_equals_serial_result(world, fast_cap, tmp_path, bmx, oracle):
    import torch.multiprocessing as mp

    n_total, m, seed = 300_007, 12, 77
    mp.spawn(_worker, args=(world, _free_port(), n_total, m, seed, str(tmp_path), fast_cap), nprocs=world, join=True)

    # serial truth over the whole text, built the same way in one piece
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd
    alpha = bmx.synth.ALPHABETS["dna"]
    pat = bmx.synth.fill_host(12345, m, seed, alpha).tobytes()
    text = bmx.synth.fill_host(0, n_total, seed, alpha)
    plants = []
    for r in range(1, world):
        seam, _ = bd.shard_bounds(n_total, world, r)
        plants += [seam - m // 2, seam - 1, seam - m + 1, seam - m, seam]
    plants += list(bmx.synth.plant_offsets(n_total, m, 40, seed))
    bmx.synth.plant_host(text, pat, plants)
    want = oracle.search(text.tobytes(), pat)

    got = np.load(tmp_path / "gathered.npy")
    assert np.array_equal(got, want)                       # rank-order concatenation is globally ascending
    for r in range(world):
        c = np.load(tmp_path / f"count_{r}.npy")
        assert c[0] == want.size and c[1:].sum() == want.size   # every rank knows the global count
    assert want.size >= 40


def test_shard_bounds_partition_the_text():
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd
    for n in (0, 1, 15, 16, 17, 1000, (1 << 33) + 5):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = bd.shard_bounds(n, world, r)
                assert lo == prev and lo <= hi <= n and lo % bd.SHARD_ALIGN == 0 or lo == n
                prev = hi
                rlo, end = bd.shard_read_range(n, 7, lo, hi)
                assert rlo == lo and end == min(n, hi + 6)
            assert prev == n

"""GPU: the exchange step of the sharded scan inside the library (bmx_exchange_*, bmx_mg_search_device) against
the oracle on the whole text.  The single-process tests run on ONE GPU too (all ranks on device 0, one stream
each); the multi-process test (cudaIpc mappings, what torchrun uses) needs two GPUs."""
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
pytestmark = pytest.mark.gpu


def _make_text(bmx, n, m, seed, alphabet, plants, world, dense_from=None):
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd
    alpha = bmx.synth.ALPHABETS[alphabet]
    text = bmx.synth.fill_host(0, n, seed, alpha)
    pat = bmx.synth.pattern_from_stream(m, seed, alpha)
    offs = list(bmx.synth.plant_offsets(n, m, plants, seed))
    for r in range(1, world):          # occurrences straddling every shard seam, and touching it
        seam, _ = bd.shard_bounds(n, world, r)
        offs += [seam - m // 2, seam - 1, seam - m + 1, seam - m, seam]
    bmx.synth.plant_host(text, pat, [o for o in offs if 0 <= o <= n - m])
    if dense_from is not None:         # a run of one byte: every start inside it matches a one-byte-run pattern
        a, b = dense_from
        text[a:b] = pat[0]
    return text, pat


def _run_world(bmx, text, pat, world, devices, head_cap, tail_cap, local_cap, steps, out_cap):
    """All ranks in this process: returns per-step (total, counts, gathered list on dst)."""
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd
    n, m = text.size, len(pat)
    shards, scanners, xs, streams, pos = [], [], [], [], []
    for r in range(world):
        dev = torch.device("cuda", devices[r])
        lo, hi = bd.shard_bounds(n, world, r)
        lo, end = bd.shard_read_range(n, m, lo, hi)
        with torch.cuda.device(dev):
            shards.append((torch.from_numpy(text[lo:end].copy()).to(dev), lo))
            streams.append(torch.cuda.Stream(device=dev))
            sc = bmx.Scanner(devices[r])
            sc.set_pattern(pat, stream=streams[-1].cuda_stream)
            scanners.append(sc)
            pos.append(torch.empty(max(local_cap, 1), dtype=torch.int64, device=dev) if local_cap else None)
            xs.append(bmx.Exchange(devices[r], r, world, dst=0, head_cap=head_cap, tail_cap=tail_cap, depth=3))
    bmx.Exchange.connect_local(xs)
    out = torch.full((max(out_cap, 1),), -7, dtype=torch.int64, device=torch.device("cuda", devices[0]))
    results = []
    for _ in range(steps):
        seq = 0
        for r in range(world):         # post everywhere first: a collect waits (on the device) for every rank's post
            with torch.cuda.device(devices[r]):
                st = streams[r].cuda_stream
                scanners[r].begin(pos[r], stream=st)
                scanners[r].scan(shards[r][0], shards[r][1], stream=st)
                seq = xs[r].post(scanners[r], st)
        for r in range(world):
            with torch.cuda.device(devices[r]):
                xs[r].collect(out if r == 0 and out_cap else None, streams[r].cuda_stream)
        per_rank = [xs[r].wait(seq) for r in range(world)]
        streams[0].synchronize()
        total, counts, glen = per_rank[0]
        for r in range(1, world):      # every rank learns the same total and per-rank counts
            assert per_rank[r][0] == total and per_rank[r][1] == counts
            assert per_rank[r][2] == 0
        results.append((total, counts, out[:glen].cpu().numpy().copy()))
    for r in range(world):
        with torch.cuda.device(devices[r]):
            torch.cuda.synchronize()
    for x in xs:
        x.close()
    for sc in scanners:
        sc.close()
    return results


def _expected_prefix(oracle, text, pat, world, head_cap, tail_cap, local_cap, out_cap):
    """What the exchange must deliver: exact counts, and the longest prefix of the global list that every
    rank in front could ship completely (own buffer, then head + tail)."""
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd
    want = oracle.search(text.tobytes(), pat)
    counts, take, complete = [], 0, True
    for r in range(world):
        lo, hi = bd.shard_bounds(text.size, world, r)
        c = int(((want >= lo) & (want < hi)).sum())
        counts.append(c)
        sent = min(c, local_cap, head_cap + tail_cap)
        if complete:
            take += sent
        if sent != c:
            complete = False
    return want, counts, want[: min(take, out_cap)]


@pytest.mark.parametrize("world,head_cap,tail_cap,local_cap", [
    (2, 4096, 0, 1 << 16),      # everything rides in the head
    (3, 8, 64, 1 << 12),        # tails in use
    (4, 4, 16, 1 << 12),        # some rank overflows head + tail: the list is a prefix
    (3, 16, 1 << 12, 20),       # a rank's own buffer is too small: prefix ends behind it
    (2, 0, 1 << 12, 1 << 12),   # no head at all
    (3, 64, 0, 0),              # count-only everywhere
])
def test_exchange_single_process(bmx, oracle, world, head_cap, tail_cap, local_cap):
    devices = [r % torch.cuda.device_count() for r in range(world)]
    text, pat = _make_text(bmx, 400_003, 9, 5 + world, "dna", 60, world)
    out_cap = 1 << 14
    want, counts, prefix = _expected_prefix(oracle, text, pat, world, head_cap, tail_cap, local_cap, out_cap)
    assert want.size >= 60
    res = _run_world(bmx, text, pat, world, devices, head_cap, tail_cap, local_cap, steps=8, out_cap=out_cap if local_cap else 0)
    for total, got_counts, got in res:          # 8 steps through a ring of 3: credits and slot reuse
        assert total == want.size and got_counts == counts
        if local_cap:
            assert np.array_equal(got, prefix)


def test_exchange_dense_tail(bmx, oracle):
    """Tens of thousands of hits per rank: the tail blocks of the post kernel and a multi-block collect."""
    world = 3
    devices = [r % torch.cuda.device_count() for r in range(world)]
    text, _ = _make_text(bmx, 300_000, 3, 3, "dna", 0, world)
    text[:] = ord("a")
    text[77_777] = ord("b")
    pat = b"aaa"
    want = oracle.search(text.tobytes(), pat)
    res = _run_world(bmx, text, pat, world, devices, head_cap=128, tail_cap=1 << 17, local_cap=1 << 17, steps=3, out_cap=1 << 19)
    for total, _, got in res:
        assert total == want.size
        assert np.array_equal(got, want)


def test_mg_search_device_equals_serial(bmx, oracle):
    """bmx_mg_search_device on all visible GPUs: sparse text, then a dense one that forces the transport to grow."""
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd
    mg = bmx.MultiGpu(0)
    R = mg.ngpus
    for dense in (False, True):
        text, pat = _make_text(bmx, 1_000_003, 12, 21, "dna", 80, R)
        if dense:
            text[100_000:400_000] = ord("C")
            pat = b"CCCCCCC"
        want = oracle.search(text.tobytes(), pat)
        shards, bases = [], []
        for r in range(R):
            lo, hi = bd.shard_bounds(text.size, R, r)
            lo, end = bd.shard_read_range(text.size, len(pat), lo, hi)
            shards.append(torch.from_numpy(text[lo:end].copy()).to(torch.device("cuda", r)))
            bases.append(lo)
        for cap in (want.size + 10, 50, 0):
            out = torch.empty(cap, dtype=torch.int64, device="cuda:0") if cap else None
            count, got, per = mg.search_device(shards, bases, pat, out)
            assert count == want.size and sum(per) == count
            if cap:
                assert np.array_equal(got.cpu().numpy(), want[:cap])
    mg.close()


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ipc_worker(rank, world, port, n, m, seed, out_dir, shared_gpu):
    import torch.distributed as dist

    import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", 0 if shared_gpu else rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo" if shared_gpu else "nccl", rank=rank, world_size=world)
    alpha = bmx.synth.ALPHABETS["ascii95"]
    pat = bmx.synth.pattern_from_stream(m, seed, alpha)
    lo, hi = bd.shard_bounds(n, world, rank)
    lo, end = bd.shard_read_range(n, m, lo, hi)
    shard = torch.empty(end - lo, dtype=torch.uint8, device=dev)
    bmx.synth.fill_device(shard, lo, seed, alpha)
    plants = list(bmx.synth.plant_offsets(n, m, 200, seed))
    for r in range(1, world):
        seam, _ = bd.shard_bounds(n, world, r)
        plants += [seam - m // 2, seam - 1, seam - m + 1, seam]
    bmx.synth.plant_device(shard, pat, [p for p in plants if p + m > lo and p < end], base=lo)

    if shared_gpu:      # gloo cannot all_gather CUDA tensors: pass the handles through CPU tensors
        x = bmx.Exchange(0, rank, world, 0, 64, 1 << 12, 4)
        mine = torch.frombuffer(bytearray(x.handle()), dtype=torch.uint8)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        x.connect(b"".join(bytes(p.numpy()) for p in parts))
        dist.barrier()
    else:
        x = bd.PeerExchange(dev, head_cap=64, tail_cap=1 << 12)
    stream = torch.cuda.current_stream().cuda_stream
    sc = bmx.Scanner(dev.index)
    sc.set_pattern(pat, stream=stream)
    pos = torch.empty(1 << 12, dtype=torch.int64, device=dev)
    out = torch.empty(1 << 14, dtype=torch.int64, device=dev)
    seq = 0
    for i in range(6):                 # the pipelined order bench.py uses: collect step i-1 behind post i
        sc.begin(pos, stream=stream)
        sc.scan(shard, lo, stream=stream)
        x.post(sc, stream)
        if i:
            seq = x.collect(out if rank == 0 else None, stream)
    seq = x.collect(out if rank == 0 else None, stream)
    total, counts, glen = x.wait(seq)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"count_{rank}.npy"), np.array([total] + counts, dtype=np.int64))
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), out[:glen].cpu().numpy())
    dist.barrier()
    x.close()
    sc.close()
    dist.destroy_process_group()


def _check_ipc(bmx, oracle, tmp_path, world, shared_gpu):
    import torch.multiprocessing as mp
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd

    n, m, seed = 3_000_017, 24, 91
    mp.spawn(_ipc_worker, args=(world, _free_port(), n, m, seed, str(tmp_path), shared_gpu), nprocs=world, join=True)
    alpha = bmx.synth.ALPHABETS["ascii95"]
    pat = bmx.synth.pattern_from_stream(m, seed, alpha)
    text = bmx.synth.fill_host(0, n, seed, alpha)
    plants = list(bmx.synth.plant_offsets(n, m, 200, seed))
    for r in range(1, world):
        seam, _ = bd.shard_bounds(n, world, r)
        plants += [seam - m // 2, seam - 1, seam - m + 1, seam]
    bmx.synth.plant_host(text, pat, plants)
    want = oracle.search(text.tobytes(), pat)
    assert np.array_equal(np.load(tmp_path / "gathered.npy"), want)
    for r in range(world):
        c = np.load(tmp_path / f"count_{r}.npy")
        assert c[0] == want.size and c[1:].sum() == want.size


def test_exchange_processes_over_ipc(bmx, oracle, tmp_path):
    """One process per GPU (what torchrun does): mailboxes mapped with cudaIpc, NVLink peer stores."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs two GPUs")
    _check_ipc(bmx, oracle, tmp_path, world, shared_gpu=False)


@pytest.mark.skipif(not os.environ.get("BMX_TEST_IPC_SHARED_GPU"), reason="opt-in: two processes time-slicing one GPU")
def test_exchange_processes_sharing_one_gpu(bmx, oracle, tmp_path):
    _check_ipc(bmx, oracle, tmp_path, 2, shared_gpu=True)

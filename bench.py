#!/usr/bin/env python
"""bench.py -- text GB/s scanned by the Boyer-Moore path on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic text:
  N = 1   workload dna_m32_4GiB      (BASELINE.json configs[1]: 4 GiB random DNA, m = 32)
  N > 1   workload ascii95_m64_shard (configs[4]: 8 GiB per GPU of 95-symbol ASCII, m = 64,
                                      sharded with (m-1) halo, counts and position
                                      lists gathered to rank 0 over NCCL) -- weak scaling.
`value`  : device-resident text, CUDA-event timed, K back-to-back scans (positions written).
`e2e`    : the same scan through the host-pointer C-ABI call bmx_search_ex: pinned host text,
           host->device copy, scan and position read-back all inside the timed region.
`roofline`: algorithmic bytes (n + 8*hits) per scan-kernel launch / its average duration vs the
           measured HBM copy bandwidth of MEASURED_PEAKS.json.
`cpu_baseline`: the reference's own serial code (oracle/_ref/libref_bm.so, 1 core) on a bounded
           sample of the same text (rank 0, N = 1 only).
--impl reference: the reference's own CPU code on all host threads (windowed), same metric.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GIB = 1 << 30
_REAL_STDOUT = None


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)
WORKLOADS = {
    # name: (alphabet, bytes per GPU, m, seed, plants per GPU, pattern source)
    "dna_m32_4GiB": dict(alphabet="dna", n=4 * GIB, m=32, seed=43, plants=1000, pattern="from_text"),
    "ascii95_m64_shard": dict(alphabet="ascii95", n=8 * GIB, m=64, seed=47, plants=1000, pattern="random"),
    "ascii95_m16_64MiB": dict(alphabet="ascii95", n=64 << 20, m=16, seed=42, plants=1000, pattern="random"),
    "bytes256_m4_4GiB": dict(alphabet="bytes256", n=4 * GIB, m=4, seed=44, plants=1000, pattern="random"),
    "bytes256_m16_4GiB": dict(alphabet="bytes256", n=4 * GIB, m=16, seed=45, plants=1000, pattern="random"),
    "bytes256_m128_4GiB": dict(alphabet="bytes256", n=4 * GIB, m=128, seed=46, plants=1000, pattern="random"),
    "aaa_1GiB": dict(alphabet="a", n=1 * GIB, m=3, seed=1, plants=0, pattern="aaa"),
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        if self.nv:
            try:   # the first queries of a process are slow (tens of ms): take them before the timed region
                self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                get = getattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or self.nv.nvmlDeviceGetCurrentClocksThrottleReasons
                get(self.h)
            except Exception:
                pass

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = get(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.001)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def make_pattern(bmx, w, total_n):
    alpha = bmx.synth.ALPHABETS[w["alphabet"]]
    if w["pattern"] == "from_text":       # cut from the text itself (so there is at least one hit)
        off = bmx.synth.mix64(w["seed"] * 7919) % (total_n - w["m"])
        return bmx.synth.fill_host(off, w["m"], w["seed"], alpha).tobytes()
    if w["pattern"] == "random":
        return bmx.synth.pattern_from_stream(w["m"], w["seed"], alpha)
    return w["pattern"].encode()


def plant_list(bmx, w, total_n, world):
    """Global plant offsets: `plants` per GPU anywhere, plus plants straddling every shard seam."""
    from parallel_implementation_of_string_matching_algorithms_opencl_b200.distributed import shard_bounds
    offs = list(bmx.synth.plant_offsets(total_n, w["m"], w["plants"] * world, w["seed"]))
    if w["plants"]:
        for r in range(1, world):
            seam, _ = shard_bounds(total_n, world, r)
            for d in (w["m"] // 2, 1, w["m"] - 1):
                offs.append(seam - d)       # starts left of the seam, ends right of it
            offs.append(seam)
    return np.array(sorted(set(int(o) for o in offs if 0 <= o <= total_n - w["m"])), dtype=np.int64)


def verify_hits(torch, text, lo, pat, pos, count, cap):
    """Size-independent checks: ascending, every reported start really matches."""
    if count == 0 or pos is None:
        return True
    k = min(count, cap)
    p = pos[:k]
    ok = bool((p[1:] > p[:-1]).all().item()) if k > 1 else True
    sample = p if k <= 100000 else p[torch.randint(0, k, (100000,), device=p.device)]
    pt = torch.frombuffer(bytearray(pat), dtype=torch.uint8).to(text.device)
    idx = (sample - lo).unsqueeze(1) + torch.arange(len(pat), device=text.device).unsqueeze(0)
    ok = ok and bool((text[idx] == pt.unsqueeze(0)).all().item())
    return ok


def run_ours(args):
    import torch
    import torch.distributed as dist

    import parallel_implementation_of_string_matching_algorithms_opencl_b200 as bmx
    from parallel_implementation_of_string_matching_algorithms_opencl_b200 import distributed as bd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "VERSION") == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout (one JSON line only)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload or ("dna_m32_4GiB" if world == 1 else "ascii95_m64_shard")
    w = dict(WORKLOADS[name])
    if args.bytes_per_gpu:
        w["n"] = args.bytes_per_gpu
    alpha = bmx.synth.ALPHABETS[w["alphabet"]]
    m = w["m"]
    total_n = w["n"] * world
    lo, hi = bd.shard_bounds(total_n, world, rank)
    lo, end = bd.shard_read_range(total_n, m, lo, hi)
    pat = make_pattern(bmx, w, total_n)
    plants = plant_list(bmx, w, total_n, world)

    text = torch.empty(end - lo, dtype=torch.uint8, device=dev)
    bmx.synth.fill_device(text, lo, w["seed"], alpha)
    mine = plants[(plants + m > lo) & (plants < end)]
    bmx.synth.plant_device(text, pat, mine, base=lo)
    torch.cuda.synchronize()

    dense = w["alphabet"] == "a"
    cap = (end - lo) if dense else max(4 * len(plants) + 1024, 1 << 16)
    pos = torch.empty(cap, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    scanner = bmx.Scanner(local)
    scanner.set_pattern(pat, variant=args.variant, stream=stream)

    # N > 1: the exchange of step i (one all-gather of counts + position lists, NCCL) is
    # enqueued behind scan i and awaited only after the next scans have been queued, so the GPUs
    # always have a scan to run while the tiny collectives and their host sync complete.
    depth = 3 if world > 1 else 1
    ring = [(pos if i == 0 else torch.empty(cap, dtype=torch.int64, device=dev),
             torch.zeros(2 + bd.FAST_GATHER_CAP, dtype=torch.int64, device=dev)) for i in range(depth)]
    inflight = []
    state = {"i": 0, "last": None}

    def step():
        if world == 1:
            scanner.begin(pos, stream=stream)
            scanner.scan(text, lo, stream=stream)
            return
        t0 = time.perf_counter()
        if len(inflight) == depth:
            state["last"] = inflight.pop(0).finish()
        t1 = time.perf_counter()
        p_i, packed_i = ring[state["i"] % depth]
        state["i"] += 1
        scanner.begin(p_i, stream=stream)
        scanner.scan(text, lo, stream=stream)
        scanner.export_result(packed_i, stream=stream)
        t2 = time.perf_counter()
        inflight.append(bd.combine_hits_start(None, p_i, group=None, device=dev, packed=packed_i))
        t3 = time.perf_counter()
        state["host"] = [a + b for a, b in zip(state.get("host", [0.0, 0.0, 0.0]), (t1 - t0, t2 - t1, t3 - t2))]

    def drain():
        while inflight:
            state["last"] = inflight.pop(0).finish()

    for _ in range(args.warmup):
        step()
    drain()
    count, stats = scanner.finish(stream=stream)
    ok = verify_hits(torch, text, lo, pat, pos, count, cap)

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scanner.set_timing(0)          # no event records between the kernels of the headline loop (they cost ~10 us/step)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            step()
        drain()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    count, stats = scanner.finish(stream=stream)
    ms_total = e0.elapsed_time(e1)
    total_hits = count
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        if state["last"] is not None:
            total_hits = state["last"][0]
            ok = ok and state["last"][0] == sum(state["last"][1])
    ms_step = ms_total / args.steps
    value = total_n / (ms_step * 1e-3) / 1e9

    # instrumented repeat of the same steps (local scan only): the library's CUDA events around the scan
    # kernel alone and around memset + scan + expand, on the launching stream, averaged over the repeats
    scanner.set_timing(2)
    torch.cuda.synchronize()
    kreps = max(3, min(args.steps, 20))
    scan_ms, whole_ms = [], []
    for _ in range(kreps):
        scanner.begin(pos, stream=stream)
        scanner.scan(text, lo, stream=stream)
        _c, st_i = scanner.finish(stream=stream)
        scan_ms.append(st_i["scan_kernel_ms"])
        whole_ms.append(st_i["device_ms"])
    stats = st_i
    stats_scan_ms = float(np.mean(scan_ms))
    step_gpu_ms = float(np.mean(whole_ms))

    # roofline of the dominant kernel (scan_kernel; expand_kernel when the text is dense): CUDA events
    # recorded by the library on the launching stream around that kernel, inside the timed region
    peak, peak_src = peaks()
    dense_text = dense
    kernel_ms = (step_gpu_ms - stats_scan_ms) if dense_text else stats_scan_ms
    alg_bytes = 8 * min(count, cap) if dense_text else (end - lo)   # expand writes the positions; scan reads the text
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    step_bytes = (end - lo) + 8 * min(count, cap)
    traffic = None
    tf = ROOT / "profiles" / "r01_traffic.json"
    if tf.exists():
        rec = json.loads(tf.read_text()).get(name)
        if rec and not args.bytes_per_gpu:
            traffic = rec["dram_read_bytes"] + rec["dram_write_bytes"]

    # ---- e2e: host-pointer C-ABI call, pinned host text, H2D + scan + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        host_text = torch.empty(end - lo, dtype=torch.uint8, pin_memory=True)
        host_text.copy_(text)
        torch.cuda.synchronize()
        e2e_cap = min(cap, 1 << 24)
        e2e_steps = max(2, min(args.steps, args.e2e_steps))
        bmx.search(host_text, pat, max_positions=e2e_cap, device=local)     # warm-up (allocations, pool)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            c2, p2 = bmx.search(host_text, pat, max_positions=e2e_cap, device=local)
            if world > 1:
                bd.combine_hits(c2, torch.from_numpy(p2 + lo).to(dev), group=None, device=dev)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        ok = ok and c2 == count
        e2e = {"value": total_n / (dt / e2e_steps) / 1e9, "unit": "GB/s", "steps": e2e_steps,
               "h2d_bytes_per_step": int(end - lo) + len(pat) + 1024 + 4 * len(pat),
               "d2h_bytes_per_step": 8 * int(min(c2, e2e_cap)) + 8,
               "api": "bmx_search_ex (host pointers, pinned text)"}
        host_sample = host_text
    else:
        host_sample = None

    # ---- cpu_baseline: the reference's own serial code on a bounded sample (rank 0, N = 1)
    cpu = None
    if world == 1 and not args.no_cpu and rank == 0:
        cpu = cpu_baseline(host_sample if host_sample is not None else text.cpu(), pat, threads=1,
                           sample_bytes=args.cpu_sample_bytes)

    if os.environ.get("BENCH_DEBUG") and "host" in state:
        n_steps = max(state["i"], 1)
        print(f"[rank {rank}] host ms/step: finish(oldest)={state['host'][0] / n_steps * 1e3:.3f} "
              f"scan-enqueue={state['host'][1] / n_steps * 1e3:.3f} collectives-enqueue={state['host'][2] / n_steps * 1e3:.3f}",
              file=sys.stderr, flush=True)
    if rank == 0:
        line = {
            "metric": "text GB/s scanned (device-timed)", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": name, "bytes_per_gpu": int(w["n"]), "total_bytes": int(total_n), "pattern_len": m,
                       "alphabet": w["alphabet"], "plants": int(len(plants)), "hits": int(total_hits),
                       "variant": stats["variant"], "tile_bytes": stats["tile_bytes"], "stages": stats["stages"],
                       "grid": stats["grid"], "l2": "inputs larger than L2 (no flush needed)" if w["n"] > (256 << 20) else "input fits L2: flushless, see DESIGN.md",
                       "verified": bool(ok), "parallelism": f"shard{world}" if world > 1 else "single"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_ms,
                         "algorithmic_bytes": int(alg_bytes), "kernel": "bmx::expand_kernel" if dense_text else "bmx::scan_kernel",
                         "kernel_timing": f"CUDA events recorded by libbmx around the kernel, mean of {kreps} instrumented repeats right after the timed loop",
                         "step_gpu_ms": step_gpu_ms, "step_algorithmic_bytes": int(step_bytes),
                         "step_achieved": step_bytes / (step_gpu_ms * 1e-3) / 1e9,
                         "step_frac": step_bytes / (step_gpu_ms * 1e-3) / 1e9 / peak},
            "e2e": e2e, "gpu_launches": int(args.steps) * 2, "clocks": clk.summary(), "cpu_baseline": cpu,
        }
        emit(line)
    scanner.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(host_text, pat, threads: int, sample_bytes: int):
    """Times oracle/_ref/libref_bm.so (the reference's own code) -- or the oracle port when the
    reference build is absent -- on a bounded prefix of the text."""
    arr = host_text.numpy() if hasattr(host_text, "numpy") else host_text
    n = int(min(arr.size, sample_bytes))
    cnt = ctypes.c_uint64()
    ref_so = ROOT / "oracle" / "_ref" / "libref_bm.so"
    legal7 = len(pat) <= 99 and max(pat) < 0x80 and int(arr[: min(n, 1 << 24)].max()) < 0x80
    if ref_so.exists() and legal7:
        lib, kind = ctypes.CDLL(str(ref_so)), "reference"
        fn = lib.ref_bm_search_windowed
    else:
        so = ROOT / "oracle" / "liboracle.so"
        if not so.exists():
            import subprocess
            subprocess.run(["bash", str(ROOT / "oracle" / "build_oracle.sh")], check=True, capture_output=True)
        lib, kind = ctypes.CDLL(str(so)), "port"
        fn = lib.oracle_search_mt
    t0 = time.perf_counter()
    rc = fn(ctypes.c_void_p(arr.ctypes.data), ctypes.c_int64(n), ctypes.c_char_p(pat), ctypes.c_int32(len(pat)),
            None, ctypes.c_int64(0), ctypes.byref(cnt), ctypes.c_int32(threads))
    dt = time.perf_counter() - t0
    assert rc == 0, rc
    return {"value": n / dt / 1e9, "unit": "GB/s", "cores": threads if threads > 0 else os.cpu_count(), "kind": kind,
            "sample": f"first {n} bytes of the workload text, {cnt.value} hits, {dt:.2f} s", "host_cores_total": os.cpu_count()}


def run_reference(args):
    """The reference's own CPU implementation of the path on all host threads (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import parallel_implementation_of_string_matching_algorithms_opencl_b200.synth as synth

    name = args.workload or ("dna_m32_4GiB" if world == 1 else "ascii95_m64_shard")
    w = dict(WORKLOADS[name])
    alpha = synth.ALPHABETS[w["alphabet"]]
    total_n = w["n"] * world
    n = int(min(w["n"], args.cpu_sample_bytes))          # bounded sample: a prefix of the text
    so = ROOT / "oracle" / "liboracle.so"
    if not so.exists():
        import subprocess
        subprocess.run(["bash", str(ROOT / "oracle" / "build_oracle.sh")], check=True, capture_output=True)
    orc = ctypes.CDLL(str(so))
    text = np.empty(n, dtype=np.uint8)
    chunk = 64 << 20

    def fill(lo):
        ln = min(chunk, n - lo)
        orc.oracle_synth_fill(ctypes.c_void_p(text.ctypes.data + lo), ctypes.c_int64(lo), ctypes.c_int64(ln),
                              ctypes.c_uint64(w["seed"]), ctypes.c_char_p(alpha), len(alpha))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(os.cpu_count() or 1) as ex:
        list(ex.map(fill, range(0, n, chunk)))

    class _B:  # minimal stand-in so make_pattern/plant_list can be shared
        pass
    b = _B()
    b.synth = synth
    pat = make_pattern(b, w, total_n)
    plants = plant_list(b, w, total_n, world)
    synth.plant_host(text, pat, plants[plants < n])

    threads = os.cpu_count() or 1
    res = cpu_baseline(text, pat, threads=-1, sample_bytes=n)       # warm-up, also sizes the sample
    rate = res["value"] * 1e9
    budget_s = 120.0                                                  # the whole K-step run stays within minutes
    if n / rate * args.steps > budget_s:
        n = max(64 << 20, int(rate * budget_s / args.steps) // (64 << 20) * (64 << 20))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = cpu_baseline(text, pat, threads=-1, sample_bytes=n)
    dt = (time.perf_counter() - t0) / args.steps
    value = n / dt / 1e9
    res["value"] = value
    res["cores"] = threads
    emit({
        "impl": "reference", "metric": "text GB/s scanned (device-timed)", "value": value, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": name, "bytes_per_gpu": int(w["n"]), "total_bytes": int(total_n), "pattern_len": w["m"],
                   "alphabet": w["alphabet"], "sample_bytes": n,
                   "note": "reference serial BM (its own kernel1.cl + BoyreMoore.cpp tables) run window-parallel on all host threads"},
        "cpu_baseline": res, "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--variant", default="auto")
    ap.add_argument("--bytes-per-gpu", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-sample-bytes", type=int, default=4 * GIB)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL's version banner)
    # are sent to stderr for the duration of the run
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

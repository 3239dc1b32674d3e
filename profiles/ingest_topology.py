#!/usr/bin/env python
"""Where does the host->device ingest rate go when several GPUs ingest at once?  (55 GB/s per GPU at
N = 1 and 2, 29 at N = 4, 23 at N = 8.)  One process, one pinned 2 GiB buffer and one stream per GPU; the buffers
are allocated either wherever the allocating thread happens to run ("unbound") or with the thread pinned to the
CPUs of the GPU's NUMA node ("bound": first touch puts the pages next to the GPU's PCIe root).  Copies are timed
with CUDA events, per GPU, for single GPUs and for groups running concurrently.
    python profiles/ingest_topology.py"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

N = torch.cuda.device_count()
SIZE = 2 << 30
REPS = 3
ALL_CPUS = os.sched_getaffinity(0)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"({e})"


def gpu_node(i):
    p = torch.cuda.get_device_properties(i)
    bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except Exception:  # noqa: BLE001
        node = None
    return bus, node


def node_cpus(node):
    cpus = set()
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus & ALL_CPUS


print("== host ==")
print(sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)'"))
print("affinity of this process:", len(ALL_CPUS), "cpus")
print(sh("cat /sys/devices/system/node/node*/meminfo 2>/dev/null | egrep 'MemTotal|MemFree' | head -8"))
print("== nvidia-smi topo -m ==")
print(sh("nvidia-smi topo -m"))
print("== GPUs ==")
nodes = []
for i in range(N):
    bus, node = gpu_node(i)
    nodes.append(node)
    print(f"GPU{i}: {bus} numa_node={node}  link gen/width: {sh(f'nvidia-smi -i {i} --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv,noheader')}")

groups = [[i] for i in range(N)]
for g in ([0, 1], [0, 1, 2, 3], [4, 5, 6, 7], [0, 4], [0, 1, 4, 5], list(range(N))):
    if all(x < N for x in g) and len(g) > 1 and g not in groups:
        groups.append(g)

dev = [torch.empty(SIZE, dtype=torch.uint8, device=f"cuda:{i}") for i in range(N)]
streams = [torch.cuda.Stream(device=i) for i in range(N)]
for mode in ("unbound", "bound"):
    bufs = []
    for i in range(N):
        if mode == "bound" and nodes[i] is not None and nodes[i] >= 0:
            os.sched_setaffinity(0, node_cpus(nodes[i]) or ALL_CPUS)
        else:
            os.sched_setaffinity(0, ALL_CPUS)
        b = torch.empty(SIZE, dtype=torch.uint8, pin_memory=True)
        b.fill_(i + 1)
        bufs.append(b)
    os.sched_setaffinity(0, ALL_CPUS)
    print(f"== pinned buffers allocated {mode} ==")
    for g in groups:
        ev = {}
        for i in g:                      # warm-up
            with torch.cuda.device(i), torch.cuda.stream(streams[i]):
                dev[i].copy_(bufs[i], non_blocking=True)
        for i in g:
            streams[i].synchronize()
        for i in g:
            with torch.cuda.device(i), torch.cuda.stream(streams[i]):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(REPS):
                    dev[i].copy_(bufs[i], non_blocking=True)
                e1.record()
                ev[i] = (e0, e1)
        rates = []
        for i in g:
            streams[i].synchronize()
            rates.append(SIZE * REPS / (ev[i][0].elapsed_time(ev[i][1]) * 1e-3) / 1e9)
        print(f"GPUs {g}: per GPU " + " ".join(f"{r:5.1f}" for r in rates) + f"  | sum {sum(rates):6.1f} GB/s", flush=True)
    del bufs
